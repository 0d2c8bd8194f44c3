/*
 *  ovlb200.h -- C ABI of the B200-native `overlapInCore` (ovl) hot path.
 *
 *  This is the drop-in boundary: plain C, opaque handle, plain
 *  pointers and sizes, no CUDA / torch types.  The C++ host driver
 *  (canu_b200/host/, the `overlapInCore` replacement executable), the Python
 *  ctypes binding (canu_b200/api.py) and a maintainer's binding inside Canu
 *  itself all sit on these entry points.
 *
 *  Every call returns 0 on success and a negative code on failure;
 *  ovlb_last_error() returns a human-readable message for the calling thread's
 *  last failure.  No exceptions cross the boundary.  The caller owns all host
 *  buffers.  A context is bound to one CUDA device and must be used by one host
 *  thread at a time.  There is NO CPU fallback: without a usable CUDA device
 *  ovlb_create() fails.
 *
 *  Reference interfaces replaced (paths under /root/reference/src/overlapInCore):
 *
 *    ovlb_params / ovlb_create        oicParameters G + prefixEditDistance ctor:
 *                                     overlapInCore.H:367-473, overlapInCore.C:293-459,
 *                                     liboverlap/prefixEditDistance.C:23-107
 *    ovlb_load_hash_reads +           Build_Hash_Index():
 *      ovlb_mark_skip_kmers +           overlapInCore-Build_Hash_Index.C:415-631 (build),
 *      ovlb_build_index                 :186-257 (Mark_Skip_Kmers)
 *    ovlb_overlap_ref_batch           Process_Overlaps() -> Find_Overlaps() ->
 *                                     Process_String_Olaps() -> Process_Matches() ->
 *                                     prefixEditDistance::Extend_Alignment() ->
 *                                     Output_Overlap()/Output_Partial_Overlap():
 *                                     overlapInCore-Process_Overlaps.C:25-122,
 *                                     overlapInCore-Find_Overlaps.C:235-336,
 *                                     overlapInCore-Process_String_Overlaps.C:355-690,
 *                                     liboverlap/prefixEditDistance-extend.C:36-183,
 *                                     overlapInCore-Output.C:27-264
 *    ovlb_record                      ovOverlap (a_iid, b_iid, ovOverlapDAT):
 *                                     ../stores/ovOverlap.H:49-78,279-291
 *    ovlb_counters                    the -s statistics: overlapInCore.C:550-558
 */
#ifndef OVLB200_H
#define OVLB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVLB_MAX_READLEN_BITS 21                       /* AS_MAX_READLEN_BITS, ../stores/sqStore.H:57 */
#define OVLB_MAX_READLEN      ((1u << OVLB_MAX_READLEN_BITS) - 1)

/* error codes */
#define OVLB_OK              0
#define OVLB_ERR_ARG        -1     /* bad argument */
#define OVLB_ERR_CUDA       -2     /* CUDA runtime / kernel failure, or no device */
#define OVLB_ERR_CAPACITY   -3     /* a device buffer overflowed; split the batch and retry */
#define OVLB_ERR_STATE      -4     /* call sequence violated */

typedef struct ovlb_ctx ovlb_ctx;

/*  Parameters of one overlap job.  All tables are computed on the HOST exactly
 *  as the reference computes them (they inherit float-parsed error rates and
 *  libm results) and uploaded at create time.  */
typedef struct {
  uint32_t       kmer_len;              /* G.Kmer_Len, 2..30                                   */
  int32_t        partial;               /* G.Doing_Partial_Overlaps (-partial)                 */
  int32_t        unique_per_pair;       /* G.Unique_Olap_Per_Pair (-u default 1, -m 0)         */
  int32_t        min_olap_len;          /* G.Min_Olap_Len (--minlength)                        */
  int32_t        use_hopeless_check;    /* G.Use_Hopeless_Check after the erate>0.06 / -z fix-up */
  uint64_t       filter_by_kmer_count;  /* G.Filter_By_Kmer_Count after overlapInCore.C:410-411; 0 = off */
  double         max_erate;             /* G.maxErate as a double (already rounded via strtof) */
  double         branch_match_value;    /* maxErate / (1 + maxErate)                           */
  double         min_branch_tail_slope; /* 1.0 if maxErate > 0.06 else 0.20                    */
  double         minkmers_exp_factor;   /* exp(-K * maxErate): computeExpected()'s constant    */
  const int32_t *edit_match_limit;      /* Edit_Match_Limit[0 .. n_edit_match_limit)           */
  uint32_t       n_edit_match_limit;    /* MAX_ERRORS                                          */
  uint32_t       max_read_len;          /* longest read that will ever be loaded (sizes scratch) */
  uint64_t       device_mem_budget;     /* bytes of HBM this context may use; 0 = 80% of free  */
} ovlb_params;

/*  A set of reads handed to the device: 2-bit packed bases exactly as sqStore
 *  stores them (4 bases per byte, first base in the two MOST significant bits;
 *  A=0 C=1 G=2 T=3; utility/src/sequence/sequence-v1.C:268-351), each read
 *  starting on a byte boundary at byte_offset[i].  Non-ACGT bases ('N') are
 *  packed as A and listed separately as (read index, position) pairs.  A read
 *  with len 0 is absent (deleted / too short / filtered by library) but keeps
 *  its slot so that read ID = first_read_id + index.  */
typedef struct {
  const uint8_t  *packed;               /* packed bases of all reads                           */
  uint64_t        packed_bytes;
  const uint64_t *byte_offset;          /* [n_reads]                                           */
  const uint32_t *len;                  /* [n_reads] bases                                     */
  uint32_t        n_reads;
  uint32_t        first_read_id;        /* ID of reads[0]                                      */
  const uint32_t *n_read;               /* [n_n] read index of each N                          */
  const uint32_t *n_pos;                /* [n_n] position of each N                            */
  uint64_t        n_n;
  /*  Optional on-device read preparation: sqStore blobs uploaded AS STORED, homopolymer compression
   *  (utility/src/sequence/sequence-v1.C:203-261) and clear-range trimming (stores/sqStore.H:397-413) done by the
   *  device.  For a read with src_len[i] > 0 the packed bytes at byte_offset[i] hold src_len[i] bases as the store
   *  keeps them; the device first collapses every run of equal bases to one base when homopoly_compress != 0, then
   *  keeps len[i] bases starting at position clear_bgn[i] of the (compressed) sequence.  src_len == NULL, or
   *  src_len[i] == 0: the packed bytes already are the final read (and only such reads may appear in the N list).  */
  const uint32_t *src_len;              /* [n_reads] or NULL                                   */
  const uint32_t *clear_bgn;            /* [n_reads] or NULL (= 0)                             */
  uint32_t        homopoly_compress;
} ovlb_reads;

/*  One overlap, bit-identical to the reference's in-memory ovOverlap
 *  (24 bytes: a_iid, b_iid, dat[0], dat[1]).  */
typedef struct {
  uint32_t a_iid, b_iid;
  uint64_t dat0;   /* ahg5:21 | ahg3:21 | evalue:16 | flipped | forOBT | forDUP | forUTG | 2 spare (LSB first) */
  uint64_t dat1;   /* bhg5:21 | bhg3:21 | span:21 | 1 spare                                                  */
} ovlb_record;

/*  Statistics; the first seven are the reference's counters, the rest are the
 *  measurement counters of SURVEY.md 8d.  All are cumulative over the context's
 *  life until ovlb_reset_counters().  */
typedef struct {
  uint64_t kmer_hits_without_olap;
  uint64_t kmer_hits_with_olap;
  uint64_t kmer_hits_skipped;
  uint64_t multi_overlap;
  uint64_t total_overlaps;
  uint64_t contained;
  uint64_t dovetail;
  uint64_t extend_calls;      /* forward/reverse extensions that ran a DP                     */
  uint64_t dp_cells;          /* DP cells, counted as the reference evaluates them            */
  uint64_t hash_kmers;        /* k-mers inserted into the index                               */
  uint64_t ref_kmers;         /* ref windows looked up (both orientations)                    */
  uint64_t seed_hits;         /* exact k-mer hits (Add_Ref calls in the reference)            */
  uint64_t seed_runs;         /* maximal diagonal runs those hits collapse into               */
  uint64_t pairs;             /* oriented candidate read pairs                                */
  uint64_t ext_busy_ns;       /* extension kernel: time its warps spent working, summed over warps (globaltimer)   */
  uint64_t ext_capacity_ns;   /* extension kernel: launched warps x kernel duration; busy/capacity = 1 - tail loss  */
} ovlb_counters;

/*  Per-stage device times of the last ovlb_build_index / ovlb_overlap_ref_batch
 *  call, in milliseconds, measured with CUDA events on the context's stream.  */
typedef struct {
  float upload_ms, encode_ms;
  float index_tuples_ms, index_sort_ms, index_table_ms, index_skip_ms;
  float probe_ms, expand_ms, sort_ms, chain_ms, extend_ms, download_ms;
  float total_ms;
} ovlb_timings;

const char *ovlb_last_error(void);
int  ovlb_device_count(void);
/*  Free and total HBM of a device, bytes (cudaMemGetInfo); lets a host that runs several contexts on one device
 *  split the memory between them through ovlb_params.device_mem_budget.  */
int  ovlb_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);

/*  Total HBM of a device from its properties: does not create a context on it (cheap on a multi-GPU node).  */
int  ovlb_device_total_memory(int device, uint64_t *total_bytes);

int  ovlb_create(int device, const ovlb_params *params, ovlb_ctx **out);
void ovlb_destroy(ovlb_ctx *ctx);

/*  Hash side.  load -> (mark skip k-mers)* -> build.  Loading replaces any
 *  previous block.  skip keys: k-mer with base j in bits [2j, 2j+1] (A0 C1 G2 T3),
 *  BOTH orientations must be supplied by the caller (Build_Hash_Index.C:226-243).  */
int  ovlb_load_hash_reads(ovlb_ctx *ctx, const ovlb_reads *reads);
int  ovlb_mark_skip_kmers(ovlb_ctx *ctx, const uint64_t *keys, uint64_t n);
int  ovlb_build_index(ovlb_ctx *ctx);

/*  Ref side: overlap every read of the batch, in both orientations, against the
 *  current hash block (only pairs with refID < hashID are computed,
 *  Find_Overlaps.C:279,320).  Records are written to out[0..*n_out); if more than
 *  out_cap are produced the call fails with OVLB_ERR_CAPACITY and *n_out holds
 *  the required capacity.  Record order is unspecified (as in the reference).  */
int  ovlb_overlap_ref_batch(ovlb_ctx *ctx, const ovlb_reads *reads,
                            ovlb_record *out, uint64_t out_cap, uint64_t *n_out);

/*  The same in three steps.  ovlb_stage_ref_batch uploads and encodes the batch on the context's COPY stream and
 *  returns without waiting when the caller's buffers are page-locked (ovlb_host_register): it may be called before
 *  ovlb_build_index, so that the upload of the ref batch overlaps the index build, and the buffers it was given must
 *  stay valid and unchanged until ovlb_run_staged returns.  ovlb_run_staged waits for the upload, runs the device
 *  pipeline (it can be repeated on the staged batch: benchmarking with inputs resident in HBM) and
 *  ovlb_fetch_records copies the records out.  */
int  ovlb_stage_ref_batch(ovlb_ctx *ctx, const ovlb_reads *reads);
int  ovlb_run_staged(ovlb_ctx *ctx, uint64_t *n_records);
/*  Double buffering for a host that streams ref batches (Process_Overlaps' read loop, overlapInCore-Process_Overlaps.C:
 *  25-122): ovlb_stage_next_ref_batch uploads and encodes batch i+1 into a SECOND slot on the copy stream -- call it
 *  before ovlb_run_staged of batch i and the two overlap; ovlb_advance_staged then makes it the current batch (the
 *  previous one is dropped).  The caller's buffers must stay valid until the batch has been run.  */
int  ovlb_stage_next_ref_batch(ovlb_ctx *ctx, const ovlb_reads *reads);
int  ovlb_advance_staged(ovlb_ctx *ctx);
int  ovlb_fetch_records(ovlb_ctx *ctx, ovlb_record *out, uint64_t out_cap, uint64_t *n_out);   /* out: host OR device memory (UVA) */

int  ovlb_get_counters(ovlb_ctx *ctx, ovlb_counters *out);
int  ovlb_reset_counters(ovlb_ctx *ctx);
int  ovlb_get_timings(ovlb_ctx *ctx, ovlb_timings *out);
uint64_t ovlb_kernel_launches(ovlb_ctx *ctx);   /* kernels launched by this context so far */

/*  Device-time bracket for benchmarks: CUDA events recorded on the context's own stream (the stream
 *  every kernel of this context is launched on).  ovlb_timer_stop() synchronises and returns the
 *  elapsed milliseconds between the two events.  */
int  ovlb_timer_start(ovlb_ctx *ctx);
int  ovlb_timer_stop(ovlb_ctx *ctx, float *ms);

/*  Page-lock (pin) a caller-owned host buffer so that the copies in ovlb_load_hash_reads /
 *  ovlb_stage_ref_batch / ovlb_fetch_records run at full PCIe/NVLink-C2C speed and asynchronously.
 *  Optional: every call also accepts pageable memory.  */
int  ovlb_host_register(const void *ptr, uint64_t bytes);
int  ovlb_host_unregister(const void *ptr);

/*  Kernel-granularity debug taps used by the parity tests (tests/ only).  After
 *  ovlb_run_staged()/ovlb_overlap_ref_batch() the candidate pairs and their
 *  ordered seed lists (the reference's String_Olap_t / Match_Node_t lists just
 *  before Process_Matches) can be read back.  */
typedef struct {
  uint32_t ref_id, hash_id;
  int32_t  dir, consistent, diag_ct, diag_bgn, diag_end, n_seeds;
  int64_t  seed_begin;
} ovlb_pair_info;
typedef struct { int32_t start, offset, len; } ovlb_seed;
int  ovlb_debug_pairs(ovlb_ctx *ctx, ovlb_pair_info *pairs, uint64_t pair_cap, uint64_t *n_pairs,
                      ovlb_seed *seeds, uint64_t seed_cap, uint64_t *n_seeds);

/*  Shape of the index built last: out[0] distinct k-mers, out[1] occurrences, out[2] slots (distinct + skip k-mers no
 *  hash read holds), out[3] 1 if the bucketed build produced it, 0 if the sorted build did (tiny block, a bucket that
 *  did not fit shared memory, or OVLB_BUCKETED=0).  */
int  ovlb_debug_index_info(ovlb_ctx *ctx, uint64_t out[4]);

/*  Extend one seed between two reads already on the device (ref batch read
 *  `ref_index` in orientation `dir`, hash read `hash_index`) exactly as
 *  Extend_Alignment would; out[0..7) = s_lo, s_hi, t_lo, t_hi, errors, kind, delta_ct.
 *  deltas (optional) receives Left_Delta.  */
int  ovlb_debug_extend(ovlb_ctx *ctx, uint32_t n, const uint32_t *ref_index, const int32_t *dir,
                       const uint32_t *hash_index, const int32_t *seed_start, const int32_t *seed_offset,
                       const int32_t *seed_len, int32_t *out7, int32_t *deltas, uint32_t delta_stride);

/* ---------------------------------------------------------------------------------------------
 *  Host-side helpers (plain C++, no CUDA): the parts of overlapInCore's main() and of the
 *  prefixEditDistance constructor that turn command-line values into the tables above.
 * ------------------------------------------------------------------------------------------- */

/*  Fill *p from command-line level values, computing every derived field and the
 *  Edit_Match_Limit table exactly as the reference does:
 *    max_erate / align_noise  must already be rounded through float (the reference parses
 *                             them with strtof, overlapInCore.C:380-382); ovlb_parse_erate() does it.
 *    use_hopeless_check       = !no_hopeless && !(max_erate > 0.06)          (overlapInCore.C:400-401)
 *    filter_by_kmer_count     = min_kmers ? floor(exp(-K*erate)*(minlen-K+1)) : 0   (:410-411)
 *    MAX_ERRORS, Branch_Match_Value, MIN_BRANCH_TAIL_SLOPE, Edit_Match_Limit
 *                             (liboverlap/prefixEditDistance.C:23-107, Binomial_Bound.C:104-188)
 *  The table is heap-allocated; release it with ovlb_params_free().  */
int    ovlb_params_init(ovlb_params *p, uint32_t kmer_len, double max_erate, double align_noise,
                        int partial, int unique_per_pair, int min_olap_len, int no_hopeless, int min_kmers,
                        uint32_t max_read_len);
void   ovlb_params_free(ovlb_params *p);
double ovlb_parse_erate(const char *text);        /* (double)strtof(text) */

/*  Pack n ASCII reads (any case; ACGT + N; anything else is an error) into the ovlb_reads wire
 *  format.  bases/offsets/lens describe the ASCII input.  The returned object owns its buffers;
 *  release with ovlb_reads_free().  Reads shorter than min_len get len 0 (slot kept).  */
typedef struct ovlb_reads_owner ovlb_reads_owner;
int    ovlb_pack_reads(const char *bases, const uint64_t *offsets, const uint32_t *lens, uint32_t n_reads,
                       uint32_t first_read_id, uint32_t min_len, ovlb_reads_owner **out);
const ovlb_reads *ovlb_reads_view(const ovlb_reads_owner *o);
void   ovlb_reads_free(ovlb_reads_owner *o);

/*  k-mer text -> key (base j in bits 2j..2j+1); returns 0 on success, fills fwd and reverse-complement keys. */
int    ovlb_kmer_keys(const char *kmer, uint32_t kmer_len, uint64_t *fwd_key, uint64_t *rc_key);

/* ---------------------------------------------------------------------------------------------
 *  Tile planning for multi-GPU runs (host only).  Replaces overlapInCorePartition's partitionLength()
 *  (overlapInCorePartition.C:127-257): the read range is cut into hash blocks x ref blocks; every tile is
 *  one `overlapInCore -h hash_bgn-hash_end -r ref_bgn-ref_end --hashdatalen hash_bases` job and tiles share
 *  nothing, so they are spread over GPUs without any collective.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t hash_bgn, hash_end, ref_bgn, ref_end;   /* inclusive read ID ranges                        */
  uint64_t hash_bases;                             /* the --hashdatalen value: bases + reads hashed   */
  uint64_t ref_bases;
  double   cost;                                   /* relative cost estimate used for load balancing  */
  int32_t  has_hash_reads;                         /* reference omits --hashdatalen when 0            */
} ovlb_tile;

/*  read_len[id] for id in 1..n_reads (read_len[0] unused).  With out == NULL only counts.
 *  strict_reference = 1: reproduce the reference's loop bounds (the last read never starts a block);
 *  0: cover every read.  */
int  ovlb_plan_tiles(const uint32_t *read_len, uint32_t n_reads, uint32_t min_olap_len,
                     uint64_t hash_block_len, uint64_t ref_block_len,
                     uint32_t hash_min, uint32_t hash_max, uint32_t ref_min, uint32_t ref_max,
                     int strict_reference, ovlb_tile *out, uint64_t out_cap, uint64_t *n_out);
/*  Bases of hash reads one context can index: the re-blocking counterpart of Canu's ovlHashBlockLength
 *  (pipelines/canu/Configure.pm:585-602 picks it from the host's memory per job; here it comes from the device's).
 *  budget_bytes = the memory the context may use.  Taken off first: the seed-run buffers (1/12 of the budget), the
 *  extension scratch (from the longest read and the error rate, at most 1/4), two ref slots of ref_batch_bases, records
 *  and slack; 90 % of the rest at 150 B per hash base (1 B of dp4 reads, ~40 B of tuple scratch, ~100 B per distinct
 *  k-mer -- a large job's blocks are mostly distinct k-mers).  Clamped to [1 Mbase, 1.5 Gbases].  */
uint64_t ovlb_hash_block_bases(uint64_t budget_bytes, uint32_t max_read_len, double max_erate, uint64_t ref_batch_bases);

/*  Cost-balanced cut of ONE hash block's ref range into n_parts contiguous tiles (at most; fewer when the range holds
 *  fewer reads).  Only refID < hashID pairs are computed (overlapInCore-Find_Overlaps.C:279,320), so the work of ref read r
 *  is ~ len_r x (hash bases with ID > r); equal-base blocks (overlapInCorePartition.C:204-226) are unequal work.  Used
 *  when a job has fewer hash blocks than GPUs: every GPU indexes the hash block and takes one part.
 *  lookup_weight = cost of looking one ref base up (probe + seeding, which does not depend on the hash IDs) relative to
 *  extending it against the WHOLE hash block: measured ~0.6 on HiFi-like reads at --maxerate 0.01 (C2: lookup 17.6 ms,
 *  extension 15.5 ms for half the pairs), ~0.002 at --maxerate 0.06 where the extension is > 99 % of the time.
 *  With out == NULL only counts.  */
int  ovlb_plan_balanced(const uint32_t *read_len, uint32_t n_reads, uint32_t min_olap_len,
                        uint32_t hash_bgn, uint32_t hash_end, uint32_t ref_bgn, uint32_t ref_end,
                        uint32_t n_parts, double lookup_weight, ovlb_tile *out, uint64_t out_cap, uint64_t *n_out);
/*  owner[i] in [0, n_workers): longest-processing-time-first on tiles[i].cost, deterministic.  */
int  ovlb_assign_tiles(const ovlb_tile *tiles, uint64_t n_tiles, uint32_t n_workers, uint32_t *owner);

/*  The store-ingest step that follows the overlapper: what ovStoreBuild / ovStoreBucketizer +
 *  ovStoreSorter do to every record before writing it.  For each input record the mirrored twin is made
 *  (ovOverlap::swapIDs, stores/ovOverlap.C:215-246), records whose evalue exceeds max_evalue lose their
 *  forUTG/forOBT/forDUP flags and are dropped with their twin (ovStoreFilter::filterOverlap,
 *  stores/ovStoreFilter.C:71-150), and the survivors are sorted by (a_iid, b_iid, dat0, dat1)
 *  (ovOverlap::operator<, stores/ovOverlap.H:265-279).  An ID outside 1..max_id fails the call with OVLB_ERR_ARG
 *  (the reference exits).  out must hold up to 2 * n records; host buffers.  */
int  ovlb_ingest_records(ovlb_ctx *ctx, const ovlb_record *in, uint64_t n, uint32_t max_evalue, uint32_t max_id,
                         ovlb_record *out, uint64_t out_cap, uint64_t *n_out);

/*  The step BEFORE the overlapper: the frequent k-mers Canu passes as `-k <file>`.  Counts the canonical
 *  k-mers (k = the context's kmer_len) of the reads loaded with ovlb_load_hash_reads and returns those that
 *  `meryl count | meryl greater-than 1 | meryl print at-least distinct=D at-least threshold=T` would print
 *  (src/pipelines/canu/Meryl.pm:529-533,603-607,663-671): count >= max(T, the smallest count v such that the k-mers of
 *  count 2..v are at least the fraction D of all distinct k-mers of count >= 2: src/meryl/src/meryl/merylOp-nextMer.C:103-115).
 *  distinct_fraction < 0 switches that filter off; min_count 0 likewise.  Each returned key is one member of the
 *  (k-mer, reverse complement) pair -- the smaller in the A0 C1 G2 T3 encoding -- with base j in bits [2j, 2j+1]; the
 *  order is unspecified.  The k-mer space is processed in 2^slice_bits passes to bound the scratch memory.
 *  stats (optional): [0] distinct k-mers of count >= 2, [1] their occurrences, [2] k-mers of count 1, [3] threshold used.
 *  More than `cap` results: OVLB_ERR_CAPACITY with *n_out = the number found so far.  */
int  ovlb_kmer_census(ovlb_ctx *ctx, uint32_t slice_bits, double distinct_fraction, uint64_t min_count,
                      uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n_out, uint64_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif
