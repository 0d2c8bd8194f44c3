//  ovl_ctx.h -- device-side data structures of one ovl context and the kernel launchers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ovlb200.h"
#include "ovl_common.cuh"

//  NVTX ranges around the C-ABI stages (tracing): visible in Nsight Systems / Nsight Compute timelines;
//  header-only NVTX v3, a no-op when no profiler is attached.
#include <nvtx3/nvToolsExt.h>
struct NvtxRange { explicit NvtxRange(const char *name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

//  A set of reads resident in HBM (one for the hash block, one for the ref batch).
struct DevReads {
  uint32_t  n = 0;
  uint32_t  first_id = 0;
  uint64_t  total_bases = 0;
  uint64_t  n_words = 0;          // dp4 words (incl. padding)
  uint64_t  n_pos = 0;            // 32-aligned position index space (sum of round_up(len,32))
  uint64_t  n_windows = 0;        // sum of max(0, len-K+1): k-mer windows, one orientation
  uint32_t  max_len = 0;
  uint64_t *fwd = nullptr;        // dp4, forward
  uint64_t *rc  = nullptr;        // dp4, reverse complement
  uint64_t *woff = nullptr;       // [n] first word of read i
  uint32_t *len = nullptr;        // [n]
  uint64_t *pbase = nullptr;      // [n+1] first position index of read i (multiple of 32)
  uint32_t *flags = nullptr;      // hash set: [n] bit0 lfrag_end_screened, bit1 rfrag_end_screened
                                  // ref  set: [2n] per (read,dir): bit0 left_end_screened, bit1 right_end_screened
  uint32_t *grp_read = nullptr;   // [n_pos/32] read index of each 32-position group
  size_t    cap_words = 0, cap_reads = 0, cap_groups = 0;
};

//  The k-mer index over the hash block: one 32-byte slot (= one DRAM sector) per DISTINCT k-mer, the slots stored in
//  PATH ORDER (by the position of the k-mer's first occurrence in the hash block) so that the consecutive windows of
//  a read mostly fall into consecutive slots, plus an open-addressed hash table k-mer -> slot index (16-byte entries,
//  load <= 0.5) for the places where the path breaks.  A slot points at the k-mer's occurrences in `occ`:
//  contiguous, and ordered by the class of the base that PRECEDES the occurrence in its hash read (0 = none: read
//  start or N; 1..4 = A,C,G,T): class c is occ[c ? end[c-1] : start, end[c]).  Bit 63 of `key` marks a skip k-mer
//  (the reference's Empty flag).
#define OVL_SKIP_BIT (1ull << 63)
struct __align__(32) IndexSlot { uint64_t key; uint32_t start; uint32_t end[5]; };
struct __align__(8) HashEntry { uint32_t idx; uint32_t fp; };      // one 64-bit word: fingerprint << 32 | slot index; buckets of four

struct DevIndex {
  uint32_t   n_slots = 0;         // slots in use: distinct k-mers + skip k-mers no hash read holds
  IndexSlot *slots = nullptr;     // [n_slots] path order
  HashEntry *htab = nullptr;      // [hcap]
  uint64_t   hcap = 0;
  uint32_t  *occ = nullptr;       // [n_occ] hash position index of each occurrence, sorted by (k-mer, class)
  uint64_t   n_occ = 0, n_distinct = 0;
  bool       built = false;
  bool       bucketed = false;    // the bucketed build produced this index (else: the sorted build)
  uint64_t  *tkey = nullptr, *tkey2 = nullptr;   // build scratch: (k-mer << 3 | class) before / after the sort
  uint32_t  *tval = nullptr;                     // build scratch: position index before the sort
  IndexSlot *tmp_slots = nullptr;                // build scratch: slots in discovery order
  uint32_t  *gk = nullptr, *gv = nullptr, *gk2 = nullptr, *gv2 = nullptr;   // build scratch: (first position, slot) before / after the path sort
  size_t     slots_cap = 0, htab_cap = 0, occ_cap = 0, tkey_cap = 0, tkey2_cap = 0, tval_cap = 0, tmp_cap = 0, gk_cap = 0, gv_cap = 0, gk2_cap = 0, gv2_cap = 0;
};

struct DevCounters {              // mirrors ovlb_counters; device-resident, atomically updated
  unsigned long long v[24];
};
enum {
  CT_HITS_WITHOUT = 0, CT_HITS_WITH, CT_HITS_SKIPPED, CT_MULTI, CT_TOTAL, CT_CONTAINED, CT_DOVETAIL,
  CT_EXT_CALLS, CT_DP_CELLS, CT_HASH_KMERS, CT_REF_KMERS, CT_SEED_HITS, CT_SEED_RUNS, CT_PAIRS,
  CT_ERR_FLAGS /* bit0: run buffer overflow, bit1: record overflow, bit2: arena overflow */,
  CT_EXT_BUSY /* ns the extension warps spent between their first and their last pair, summed over warps */,
  CT_EXT_CAPACITY /* host only: launched warps x kernel duration, ns */,
  CT_N
};

struct DevParams {                // kernel-visible job parameters
  int       K;
  int       partial, unique, min_olap_len, use_hopeless;
  unsigned long long filter_by_kmer_count;
  double    erate, bmv, min_tail_slope, minkmers_factor;
  const int32_t *eml;             // Edit_Match_Limit
  uint32_t  n_eml;
  int       ext_prefetch;         // tuning: 0 none, 1 L1, 2 L2 prefetch of the read lines ahead of the DP wavefront
};

//  Per-warp scratch of the extension kernel.
struct ExtScratch {
  int       n_warps = 0;
  int       emax = 0;             // max Error_Limit this scratch supports
  uint64_t  arena_cap = 0;        // uint2 entries per warp
  uint32_t  gring_cap = 0;        // ints per global ring (power of two)
  uint2    *arena = nullptr;      // [n_warps][arena_cap]   from-code bit planes
  int2     *row_meta = nullptr;   // [n_warps][emax+2]  (leftmost diagonal of the row, first arena word of the row)
  int32_t  *gring = nullptr;      // [n_warps][2][gring_cap]
  uint8_t  *path = nullptr;       // [n_warps][emax+2]
  int32_t  *ival = nullptr;       // [n_warps][emax+2]
  uint32_t *ikc = nullptr;        // [n_warps][emax+2]
  int32_t  *ldelta = nullptr;     // [n_warps][emax+2]
  int32_t  *rdelta = nullptr;     // [n_warps][emax+2]
};

struct PairRec {                  // one oriented candidate pair after chaining
  uint32_t ref_idx, hash_idx;
  int32_t  dir, consistent;
  int32_t  diag_ct, diag_bgn, diag_end;
  int32_t  n_seeds;               // 0 when the pair was dropped (hopeless / --minkmers)
  int64_t  seed_begin;
};

struct ovlb_ctx {
  int          device = 0;
  cudaStream_t stream = nullptr;
  ovlb_params  P;
  DevParams    dp;
  int32_t     *d_eml = nullptr;
  DevReads     hash, ref;
  DevIndex     index;
  DevCounters *d_counters = nullptr;
  ExtScratch   ext;
  uint64_t     mem_budget = 0;
  uint32_t     ht_fpmask = 0xFFFFFFFFu;   // source of the async copy to the device symbol (must outlive the call)
  int          sm_count = 148;
  int          bucket_tb = 11;            // table bits k_bucket_group2 starts with (12 once a block needed the large table)
  bool         bucket_attr_set = false;   // k_bucket_group's dynamic shared-memory size was raised on this context's device

  //  staging
  //  one set per side (0 = hash, 1 = ref): the ref batch is uploaded on `copy_stream` while the index is being built
  struct Staging {
    uint8_t  *d_packed = nullptr;   size_t packed_cap = 0;
    uint64_t *d_boff = nullptr;     size_t boff_cap = 0;
    uint32_t *d_nread = nullptr, *d_npos = nullptr; size_t nn_cap = 0;
    uint32_t *d_srclen = nullptr;   size_t srclen_cap = 0;   // [2n] stored length | clear-range begin of blobs uploaded as stored
    std::vector<uint64_t> h_woff, h_pbase;          // host copies that must outlive the asynchronous upload
  } stg[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t  ref_ready = nullptr, ref_up0 = nullptr, ref_up1 = nullptr;   // ref upload + encode done; brackets for its timing
  bool         ref_pending = false;                   // a ref upload is in flight on copy_stream
  //  second ref slot (ovlb_stage_next_ref_batch): batch i+1 is uploaded and encoded on the copy stream while batch i runs;
  //  ovlb_advance_staged swaps the slots
  DevReads     ref_next;
  Staging      stg_next;
  cudaEvent_t  next_ready = nullptr, next_up0 = nullptr, next_up1 = nullptr;
  bool         next_pending = false, staged_next = false;
  uint8_t  *h_pinned = nullptr;   size_t pinned_cap = 0;
  std::vector<uint64_t> skip_keys;

  //  lookup products for the current ref batch
  uint32_t *ref_valid = nullptr;  size_t ref_valid_cap = 0;   // [2 * ref.n_pos / 32] bit set iff the window hits a non-skip k-mer
  uint4    *item_small = nullptr, *item_large = nullptr;      // [run_cap] occurrence ranges holding run heads
  uint64_t *run_key = nullptr, *run_val = nullptr, *run_key2 = nullptr, *run_val2 = nullptr;
  uint64_t  run_cap = 0;
  OvlRun   *runs_extra = nullptr; size_t runs_extra_cap = 0;
  uint64_t  n_runs = 0;
  uint32_t *pair_flag = nullptr, *pair_idx = nullptr;
  const uint32_t *pair_order = nullptr;   // extension order of the pairs (heaviest first) or null = index order; aliases chain scratch
  void     *cub_temp = nullptr;   size_t cub_temp_cap = 0;
  PairRec  *pairs = nullptr;      uint64_t pair_cap = 0, n_pairs = 0;
  int32_t  *seed_start = nullptr, *seed_off = nullptr, *seed_len = nullptr;   // [n_runs] list order per pair
  int32_t  *sim_nxt = nullptr, *sim_hits = nullptr, *sim_act = nullptr, *sim_order = nullptr;
  uint8_t  *seed_alive = nullptr;
  uint64_t  seed_cap = 0;
  ovlb_record *d_records = nullptr;  uint64_t rec_cap = 0;  uint64_t n_records = 0;
  unsigned long long *d_work = nullptr;   // [8] device-side cursors and counts
  unsigned long long host_counters[24] = {0};   // counters known on the host (added to the device ones on readout)

  ovlb_timings timings;
  cudaEvent_t  ev_start = nullptr, ev_stop = nullptr;     // ovlb_timer_start / ovlb_timer_stop
  uint64_t     launches = 0;
  uint64_t     ext_warps_launched = 0;             // warps of the last k_extend_pairs launch (0: none)
  bool         staged = false;
};

//  launchers (each returns cudaError_t of the launch; all on ctx->stream)
void ovl_set_error(const std::string &msg);
