//  ovl_capi.cu -- the C ABI of include/ovlb200.h over the CUDA pipeline.
#include "ovl_ctx.h"

#include <cstring>
#include <cstdlib>
#include <cmath>
#include <utility>

static thread_local std::string g_last_error;
void ovl_set_error(const std::string &msg) { g_last_error = msg; }

int ovl_upload_reads(ovlb_ctx *c, const ovlb_reads *in, DevReads &dst, int slot, float *upload_ms, float *encode_ms);
int ovl_build_index(ovlb_ctx *c);
int ovl_seed_ref_batch(ovlb_ctx *c);
int ovl_extend_pairs(ovlb_ctx *c);
int ovl_prepare_ext_scratch(ovlb_ctx *c);
int ovl_kmer_census(ovlb_ctx *c, uint32_t slice_bits, double distinct_fraction, uint64_t min_count,
                    uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n_out, uint64_t stats[4]);
int ovl_ingest_records(ovlb_ctx *c, const ovlb_record *in, uint64_t n, uint32_t max_evalue, uint32_t max_id,
                       ovlb_record *out, uint64_t out_cap, uint64_t *n_out);
int ovl_debug_extend(ovlb_ctx *c, uint32_t n, const uint32_t *ref_index, const int32_t *dir, const uint32_t *hash_index,
                     const int32_t *seed_start, const int32_t *seed_offset, const int32_t *seed_len,
                     int32_t *out7, int32_t *deltas, uint32_t delta_stride);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ovl_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return OVLB_ERR_CUDA; } } while (0)

struct EvT {
  cudaEvent_t a, b; cudaStream_t s;
  EvT(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
  float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b); return ms; }
};

extern "C" {

const char *ovlb_last_error(void) { return g_last_error.c_str(); }

int ovlb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ovlb_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes) {
  size_t f = 0, t = 0;
  if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) {
    cudaGetLastError(); ovl_set_error("ovlb_device_memory: bad device or no CUDA"); return OVLB_ERR_CUDA;
  }
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return OVLB_OK;
}

int ovlb_device_total_memory(int device, uint64_t *total_bytes) {
  cudaDeviceProp prop;
  if (!total_bytes || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); ovl_set_error("ovlb_device_total_memory: bad device or no CUDA"); return OVLB_ERR_CUDA; }
  *total_bytes = prop.totalGlobalMem;
  return OVLB_OK;
}

int ovlb_create(int device, const ovlb_params *p, ovlb_ctx **out) {
  if (!p || !out) { ovl_set_error("ovlb_create: null argument"); return OVLB_ERR_ARG; }
  *out = nullptr;
  if (p->kmer_len < 2 || p->kmer_len > 30) { ovl_set_error("ovlb_create: kmer_len must be in 2..30 (k-mer + 3 class bits + 1 sentinel bit must fit 64 bits)"); return OVLB_ERR_ARG; }
  if (!p->edit_match_limit || p->n_edit_match_limit < 2) { ovl_set_error("ovlb_create: edit_match_limit table missing"); return OVLB_ERR_ARG; }
  if (!(p->max_erate > 0.0) || p->max_erate >= 1.0) { ovl_set_error("ovlb_create: max_erate must be in (0,1)"); return OVLB_ERR_ARG; }
  if (p->max_read_len == 0 || p->max_read_len > OVLB_MAX_READLEN) { ovl_set_error("ovlb_create: max_read_len must be in 1..AS_MAX_READLEN"); return OVLB_ERR_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    ovl_set_error("ovlb_create: no CUDA device available (this library has no CPU fallback)");
    return OVLB_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { ovl_set_error("ovlb_create: bad device index"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(device));
  ovlb_ctx *c = new ovlb_ctx();
  c->device = device;
  c->P = *p;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  c->mem_budget = p->device_mem_budget ? p->device_mem_budget : (uint64_t)(free_b * 0.8);
  //  the index is probed one 32-byte sector at a time at random: do not let L2 fetch 64/128 B around each miss
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); cudaGetLastError();
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ref_ready, cudaEventDisableTiming));
  CK(cudaEventCreate(&c->ref_up0)); CK(cudaEventCreate(&c->ref_up1));
  CK(cudaEventCreateWithFlags(&c->next_ready, cudaEventDisableTiming));
  CK(cudaEventCreate(&c->next_up0)); CK(cudaEventCreate(&c->next_up1));
  CK(cudaMalloc((void **)&c->d_eml, (size_t)p->n_edit_match_limit * 4));
  CK(cudaMemcpyAsync(c->d_eml, p->edit_match_limit, (size_t)p->n_edit_match_limit * 4, cudaMemcpyHostToDevice, c->stream));
  c->P.edit_match_limit = nullptr;                     // caller's buffer is not retained
  CK(cudaMalloc((void **)&c->d_counters, sizeof(DevCounters)));
  CK(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), c->stream));
  CK(cudaMalloc((void **)&c->d_work, 64));
  CK(cudaMemsetAsync(c->d_work, 0, 64, c->stream));
  CK(cudaStreamSynchronize(c->stream));                 // all setup is ordered on the context's own (non-blocking) stream
  c->dp.K = (int)p->kmer_len;
  c->dp.partial = p->partial; c->dp.unique = p->unique_per_pair; c->dp.min_olap_len = p->min_olap_len;
  c->dp.use_hopeless = p->use_hopeless_check;
  c->dp.filter_by_kmer_count = p->filter_by_kmer_count;
  c->dp.erate = p->max_erate; c->dp.bmv = p->branch_match_value; c->dp.min_tail_slope = p->min_branch_tail_slope;
  c->dp.minkmers_factor = p->minkmers_exp_factor;
  c->dp.eml = c->d_eml; c->dp.n_eml = p->n_edit_match_limit;
  c->dp.ext_prefetch = 1;
  if (const char *ev = getenv("OVLB_EXT_PREFETCH")) c->dp.ext_prefetch = atoi(ev);
  memset(&c->timings, 0, sizeof(c->timings));
  *out = c;
  return OVLB_OK;
}

static void free_reads(DevReads &d) {
  void *ptrs[] = { d.fwd, d.rc, d.woff, d.len, d.pbase, d.flags, d.grp_read };
  for (void *p : ptrs) if (p) cudaFree(p);
  d = DevReads();
}

void ovlb_destroy(ovlb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  free_reads(c->hash); free_reads(c->ref); free_reads(c->ref_next);
  { void *np[] = { c->stg_next.d_packed, c->stg_next.d_boff, c->stg_next.d_nread, c->stg_next.d_npos, c->stg_next.d_srclen };
    for (void *p : np) if (p) cudaFree(p); }
  if (c->next_ready) { cudaEventDestroy(c->next_ready); cudaEventDestroy(c->next_up0); cudaEventDestroy(c->next_up1); }
  void *ptrs[] = { c->d_eml, c->d_counters, c->d_work, c->index.slots, c->index.htab, c->index.tmp_slots, c->index.gk, c->index.gv, c->index.gk2, c->index.gv2,
                   c->index.occ, c->index.tkey, c->index.tkey2, c->index.tval,
                   c->ext.arena, c->ext.row_meta, c->ext.gring, c->ext.path, c->ext.ival, c->ext.ikc, c->ext.ldelta, c->ext.rdelta,
                   c->stg[0].d_packed, c->stg[0].d_boff, c->stg[0].d_nread, c->stg[0].d_npos, c->stg[0].d_srclen, c->stg[1].d_srclen,
                   c->stg[1].d_packed, c->stg[1].d_boff, c->stg[1].d_nread, c->stg[1].d_npos, c->ref_valid, c->item_small, c->item_large,
                   c->run_key, c->run_val, c->run_key2, c->run_val2, c->runs_extra, c->pair_flag, c->pair_idx, c->cub_temp, c->pairs,
                   c->seed_start, c->seed_off, c->seed_len, c->sim_nxt, c->sim_hits, c->sim_act, c->sim_order, c->seed_alive, c->d_records };
  for (void *p : ptrs) if (p) cudaFree(p);
  if (c->ev_start) { cudaEventDestroy(c->ev_start); cudaEventDestroy(c->ev_stop); }
  cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ref_ready) { cudaEventDestroy(c->ref_ready); cudaEventDestroy(c->ref_up0); cudaEventDestroy(c->ref_up1); }
  delete c;
}

int ovlb_load_hash_reads(ovlb_ctx *c, const ovlb_reads *reads) {
  NvtxRange nvtx_("ovlb_load_hash_reads");
  if (!c || !reads) { ovl_set_error("ovlb_load_hash_reads: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (reads->n_reads >= (1u << OVL_RUNKEY_HASH_BITS)) { ovl_set_error("hash block has too many reads (max 16777215); use a smaller hash block"); return OVLB_ERR_CAPACITY; }
  c->index.built = false;
  c->skip_keys.clear();
  return ovl_upload_reads(c, reads, c->hash, 0, &c->timings.upload_ms, &c->timings.encode_ms);
}

int ovlb_mark_skip_kmers(ovlb_ctx *c, const uint64_t *keys, uint64_t n) {
  if (!c || (n && !keys)) { ovl_set_error("ovlb_mark_skip_kmers: null argument"); return OVLB_ERR_ARG; }
  if (c->index.built) { ovl_set_error("ovlb_mark_skip_kmers: must be called before ovlb_build_index"); return OVLB_ERR_STATE; }
  const uint64_t lim = 1ull << (2 * c->P.kmer_len);
  for (uint64_t i = 0; i < n; i++) {
    if (keys[i] >= lim) { ovl_set_error("ovlb_mark_skip_kmers: key wider than 2*kmer_len bits"); return OVLB_ERR_ARG; }
    c->skip_keys.push_back(keys[i]);
  }
  return OVLB_OK;
}

int ovlb_build_index(ovlb_ctx *c) {
  NvtxRange nvtx_("ovlb_build_index");
  if (!c) { ovl_set_error("ovlb_build_index: null context"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (c->hash.fwd == nullptr && c->hash.n == 0 && c->hash.cap_reads == 0) { ovl_set_error("ovlb_build_index: no hash reads loaded"); return OVLB_ERR_STATE; }
  return ovl_build_index(c);
}

int ovlb_stage_ref_batch(ovlb_ctx *c, const ovlb_reads *reads) {
  NvtxRange nvtx_("ovlb_stage_ref_batch");
  if (!c || !reads) { ovl_set_error("ovlb_stage_ref_batch: null argument"); return OVLB_ERR_ARG; }
  if (reads->n_reads >= (1u << OVL_RUNKEY_REF_BITS)) { ovl_set_error("ref batch has too many reads (max 262143); split it"); return OVLB_ERR_CAPACITY; }
  CK(cudaSetDevice(c->device));
  c->staged = false;
  int rc = ovl_upload_reads(c, reads, c->ref, 1, &c->timings.upload_ms, &c->timings.encode_ms);
  if (rc) return rc;
  c->staged = true;
  return OVLB_OK;
}

int ovlb_stage_next_ref_batch(ovlb_ctx *c, const ovlb_reads *reads) {
  NvtxRange nvtx_("ovlb_stage_next_ref_batch");
  if (!c || !reads) { ovl_set_error("ovlb_stage_next_ref_batch: null argument"); return OVLB_ERR_ARG; }
  if (reads->n_reads >= (1u << OVL_RUNKEY_REF_BITS)) { ovl_set_error("ref batch has too many reads (max 262143); split it"); return OVLB_ERR_CAPACITY; }
  if (c->staged_next) { ovl_set_error("ovlb_stage_next_ref_batch: a next batch is already staged; call ovlb_advance_staged first"); return OVLB_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  float up = 0, en = 0;
  int rc = ovl_upload_reads(c, reads, c->ref_next, 2, &up, &en);
  if (rc) return rc;
  c->staged_next = true;
  return OVLB_OK;
}

int ovlb_advance_staged(ovlb_ctx *c) {
  if (!c) { ovl_set_error("ovlb_advance_staged: null context"); return OVLB_ERR_ARG; }
  std::swap(c->ref, c->ref_next);
  std::swap(c->stg[1], c->stg_next);
  std::swap(c->ref_ready, c->next_ready); std::swap(c->ref_up0, c->next_up0); std::swap(c->ref_up1, c->next_up1);
  std::swap(c->ref_pending, c->next_pending);
  c->staged = c->staged_next;
  c->staged_next = false;
  return OVLB_OK;
}

int ovlb_run_staged(ovlb_ctx *c, uint64_t *n_records) {
  NvtxRange nvtx_("ovlb_run_staged");
  if (!c) { ovl_set_error("ovlb_run_staged: null context"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (!c->staged) { ovl_set_error("ovlb_run_staged: no staged ref batch"); return OVLB_ERR_STATE; }
  if (!c->index.built) { ovl_set_error("ovlb_run_staged: build the index first"); return OVLB_ERR_STATE; }
  if (c->ref_pending) {                                  // the upload ran beside whatever the caller did since staging
    CK(cudaStreamWaitEvent(c->stream, c->ref_ready, 0));
    CK(cudaEventSynchronize(c->ref_up1));
    float ms = 0; cudaEventElapsedTime(&ms, c->ref_up0, c->ref_up1);
    c->timings.upload_ms = ms; c->timings.encode_ms = 0;
    c->ref_pending = false;
  }
  //  a failed run (buffer overflow -> the caller splits the batch and retries) must not leave its partial counts behind
  DevCounters snap; unsigned long long hsnap[24];
  CK(cudaMemcpyAsync(&snap, c->d_counters, sizeof(snap), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(hsnap, c->host_counters, sizeof(hsnap));
  auto rollback = [&]() { snap.v[CT_ERR_FLAGS] = 0; cudaMemcpyAsync(c->d_counters, &snap, sizeof(snap), cudaMemcpyHostToDevice, c->stream);
                          cudaStreamSynchronize(c->stream); memcpy(c->host_counters, hsnap, sizeof(hsnap)); };
  EvT tt(c->stream);
  c->n_records = 0;
  int rc;
  { NvtxRange r1("seed: probe + expand + sort + chain"); rc = ovl_seed_ref_batch(c); }
  if (rc) { tt.stop(); rollback(); return rc; }
  NvtxRange r2("extend: k_extend_pairs");
  if (c->n_pairs) { rc = ovl_prepare_ext_scratch(c); if (rc) { tt.stop(); rollback(); return rc; } }   // allocation stays outside the kernel's bracket
  EvT te(c->stream);
  c->ext_warps_launched = 0;
  rc = ovl_extend_pairs(c);
  if (rc) { te.stop(); tt.stop(); rollback(); return rc; }
  unsigned long long w[4] = {0, 0, 0, 0}, flags = 0;
  CK(cudaMemcpyAsync(w, c->d_work, 32, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&flags, &c->d_counters->v[CT_ERR_FLAGS], 8, cudaMemcpyDeviceToHost, c->stream));
  cudaError_t e = cudaStreamSynchronize(c->stream);
  c->timings.extend_ms = te.stop();
  c->host_counters[CT_EXT_CAPACITY] += (unsigned long long)((double)c->ext_warps_launched * (double)c->timings.extend_ms * 1e6);
  c->timings.total_ms = tt.stop();
  if (e != cudaSuccess) { ovl_set_error(std::string("extension kernel failed: ") + cudaGetErrorString(e)); return OVLB_ERR_CUDA; }
  if (flags) {
    rollback();
    ovl_set_error("device buffer overflow in the extension kernel (flags " + std::to_string(flags) + ")");
    return OVLB_ERR_CAPACITY;
  }
  c->n_records = c->n_pairs ? w[2] : 0;
  if (n_records) *n_records = c->n_records;
  return OVLB_OK;
}

int ovlb_fetch_records(ovlb_ctx *c, ovlb_record *out, uint64_t out_cap, uint64_t *n_out) {
  NvtxRange nvtx_("ovlb_fetch_records");
  if (!c || !n_out) { ovl_set_error("ovlb_fetch_records: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  *n_out = c->n_records;
  if (c->n_records > out_cap) { ovl_set_error("ovlb_fetch_records: output buffer too small"); return OVLB_ERR_CAPACITY; }
  if (c->n_records && !out) { ovl_set_error("ovlb_fetch_records: null output"); return OVLB_ERR_ARG; }
  EvT t(c->stream);
  if (c->n_records) CK(cudaMemcpyAsync(out, c->d_records, c->n_records * sizeof(ovlb_record), cudaMemcpyDefault, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->timings.download_ms = t.stop();
  return OVLB_OK;
}

int ovlb_overlap_ref_batch(ovlb_ctx *c, const ovlb_reads *reads, ovlb_record *out, uint64_t out_cap, uint64_t *n_out) {
  int rc = ovlb_stage_ref_batch(c, reads);
  if (rc) return rc;
  uint64_t n = 0;
  rc = ovlb_run_staged(c, &n);
  if (rc) return rc;
  return ovlb_fetch_records(c, out, out_cap, n_out);
}

int ovlb_get_counters(ovlb_ctx *c, ovlb_counters *out) {
  if (!c || !out) { ovl_set_error("ovlb_get_counters: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  DevCounters h;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(&h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
  out->kmer_hits_without_olap = h.v[CT_HITS_WITHOUT]; out->kmer_hits_with_olap = h.v[CT_HITS_WITH];
  out->kmer_hits_skipped = h.v[CT_HITS_SKIPPED]; out->multi_overlap = h.v[CT_MULTI];
  out->total_overlaps = h.v[CT_TOTAL]; out->contained = h.v[CT_CONTAINED]; out->dovetail = h.v[CT_DOVETAIL];
  out->extend_calls = h.v[CT_EXT_CALLS]; out->dp_cells = h.v[CT_DP_CELLS]; out->hash_kmers = h.v[CT_HASH_KMERS];
  out->ref_kmers = h.v[CT_REF_KMERS]; out->seed_hits = h.v[CT_SEED_HITS]; out->seed_runs = h.v[CT_SEED_RUNS];
  out->pairs = h.v[CT_PAIRS];
  out->ext_busy_ns = h.v[CT_EXT_BUSY]; out->ext_capacity_ns = c->host_counters[CT_EXT_CAPACITY];
  out->hash_kmers += c->host_counters[CT_HASH_KMERS];
  out->ref_kmers  += c->host_counters[CT_REF_KMERS];
  out->seed_runs  += c->host_counters[CT_SEED_RUNS];
  return OVLB_OK;
}

int ovlb_reset_counters(ovlb_ctx *c) {
  if (!c) { ovl_set_error("ovlb_reset_counters: null context"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memset(c->host_counters, 0, sizeof(c->host_counters));
  return OVLB_OK;
}

int ovlb_get_timings(ovlb_ctx *c, ovlb_timings *out) {
  if (!c || !out) { ovl_set_error("ovlb_get_timings: null argument"); return OVLB_ERR_ARG; }
  *out = c->timings;
  return OVLB_OK;
}

uint64_t ovlb_kernel_launches(ovlb_ctx *c) { return c ? c->launches : 0; }

int ovlb_host_register(const void *ptr, uint64_t bytes) {
  if (!ptr || bytes == 0) return OVLB_OK;
  cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), (size_t)bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return OVLB_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); ovl_set_error(std::string("cudaHostRegister: ") + cudaGetErrorString(e)); return OVLB_ERR_CUDA; }
  return OVLB_OK;
}

int ovlb_host_unregister(const void *ptr) {
  if (!ptr) return OVLB_OK;
  cudaError_t e = cudaHostUnregister(const_cast<void *>(ptr));
  if (e != cudaSuccess) { cudaGetLastError(); if (e != cudaErrorHostMemoryNotRegistered) { ovl_set_error(std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); return OVLB_ERR_CUDA; } }
  return OVLB_OK;
}

int ovlb_timer_start(ovlb_ctx *c) {
  if (!c) { ovl_set_error("ovlb_timer_start: null context"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (!c->ev_start) { CK(cudaEventCreate(&c->ev_start)); CK(cudaEventCreate(&c->ev_stop)); }
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaEventRecord(c->ev_start, c->stream));
  return OVLB_OK;
}

int ovlb_timer_stop(ovlb_ctx *c, float *ms) {
  if (!c || !ms || !c->ev_start) { ovl_set_error("ovlb_timer_stop: no timer running"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev_stop, c->stream));
  CK(cudaEventSynchronize(c->ev_stop));
  CK(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
  return OVLB_OK;
}

int ovlb_ingest_records(ovlb_ctx *c, const ovlb_record *in, uint64_t n, uint32_t max_evalue, uint32_t max_id,
                        ovlb_record *out, uint64_t out_cap, uint64_t *n_out) {
  if (!c || !n_out || (n && (!in || !out))) { ovl_set_error("ovlb_ingest_records: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  return ovl_ingest_records(c, in, n, max_evalue, max_id, out, out_cap, n_out);
}

int ovlb_kmer_census(ovlb_ctx *c, uint32_t slice_bits, double distinct_fraction, uint64_t min_count,
                     uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n_out, uint64_t stats[4]) {
  if (!c || !n_out || (cap && (!kmers || !counts))) { ovl_set_error("ovlb_kmer_census: null argument"); return OVLB_ERR_ARG; }
  if (distinct_fraction > 1.0) { ovl_set_error("ovlb_kmer_census: distinct_fraction must be <= 1 (negative = off)"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (c->hash.cap_reads == 0) { ovl_set_error("ovlb_kmer_census: load the reads with ovlb_load_hash_reads first"); return OVLB_ERR_STATE; }
  return ovl_kmer_census(c, slice_bits, distinct_fraction, min_count, kmers, counts, cap, n_out, stats);
}

int ovlb_debug_index_info(ovlb_ctx *c, uint64_t out[4]) {
  if (!c || !out) { ovl_set_error("ovlb_debug_index_info: null argument"); return OVLB_ERR_ARG; }
  if (!c->index.built) { ovl_set_error("ovlb_debug_index_info: no index built"); return OVLB_ERR_STATE; }
  out[0] = c->index.n_distinct; out[1] = c->index.n_occ; out[2] = c->index.n_slots; out[3] = c->index.bucketed ? 1 : 0;
  return OVLB_OK;
}

int ovlb_debug_pairs(ovlb_ctx *c, ovlb_pair_info *pairs, uint64_t pair_cap, uint64_t *n_pairs,
                     ovlb_seed *seeds, uint64_t seed_cap, uint64_t *n_seeds) {
  if (!c || !n_pairs || !n_seeds) { ovl_set_error("ovlb_debug_pairs: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  *n_pairs = c->n_pairs; *n_seeds = c->n_runs;
  if (c->n_pairs > pair_cap || c->n_runs > seed_cap) { ovl_set_error("ovlb_debug_pairs: buffers too small"); return OVLB_ERR_CAPACITY; }
  if (c->n_pairs == 0) return OVLB_OK;
  std::vector<PairRec> hp(c->n_pairs);
  std::vector<int32_t> s0(c->n_runs), s1(c->n_runs), s2(c->n_runs);
  CK(cudaMemcpy(hp.data(), c->pairs, c->n_pairs * sizeof(PairRec), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(s0.data(), c->seed_start, c->n_runs * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(s1.data(), c->seed_off, c->n_runs * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(s2.data(), c->seed_len, c->n_runs * 4, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < c->n_pairs; i++) {
    const PairRec &p = hp[i];
    int64_t end = (i + 1 < c->n_pairs) ? hp[i + 1].seed_begin : (int64_t)c->n_runs;
    pairs[i].ref_id = c->ref.first_id + p.ref_idx; pairs[i].hash_id = c->hash.first_id + p.hash_idx;
    pairs[i].dir = p.dir; pairs[i].consistent = p.consistent;
    pairs[i].diag_ct = p.diag_ct; pairs[i].diag_bgn = p.diag_bgn; pairs[i].diag_end = p.diag_end;
    pairs[i].n_seeds = (int32_t)(end - p.seed_begin);     // all seeds, even if the pair was dropped before extension
    pairs[i].seed_begin = p.seed_begin;
    if (p.n_seeds == 0) pairs[i].consistent |= 0x100;     // marker: dropped (hopeless / --minkmers)
  }
  for (uint64_t i = 0; i < c->n_runs; i++) { seeds[i].start = s0[i]; seeds[i].offset = s1[i]; seeds[i].len = s2[i]; }
  return OVLB_OK;
}

int ovlb_debug_extend(ovlb_ctx *c, uint32_t n, const uint32_t *ref_index, const int32_t *dir, const uint32_t *hash_index,
                      const int32_t *seed_start, const int32_t *seed_offset, const int32_t *seed_len,
                      int32_t *out7, int32_t *deltas, uint32_t delta_stride) {
  if (!c || !ref_index || !dir || !hash_index || !seed_start || !seed_offset || !seed_len || !out7) { ovl_set_error("ovlb_debug_extend: null argument"); return OVLB_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  if (!c->staged) { ovl_set_error("ovlb_debug_extend: stage a ref batch first"); return OVLB_ERR_STATE; }
  if (n == 0) return OVLB_OK;
  return ovl_debug_extend(c, n, ref_index, dir, hash_index, seed_start, seed_offset, seed_len, out7, deltas, delta_stride);
}

}  // extern "C"
