//  ovl_ingest.cu -- the step right after the overlapper: what Canu's overlap-store build does to
//  every record before it is written -- mirror, filter, sort -- on the GPU, so that an overlapper running on the device
//  can hand the store pre-mirrored, pre-sorted slices instead of feeding a disk-bound bucket sort.
//
//  Replaces, bit-exactly (paths under /root/reference/src/stores):
//    ovStoreFilter::filterOverlap        ovStoreFilter.C:71-150   ID range check, mirrored twin, error-rate filter
//    ovOverlap::swapIDs                  ovOverlap.C:215-246      hang swap (and 5'/3' reversal for flipped overlaps)
//    std::sort + ovOverlap::operator<    ovStoreBuild.C:252, ovStoreSorter.C:223, ovOverlap.H:265-279
//
//  Byte/integer work, HBM-bound: one pass writes both twins, then a least-significant-key-first sequence of three stable
//  64-bit radix sorts of (key, index) -- dat1, dat0, (a_iid, b_iid) -- gives the reference's full 192-bit order whatever
//  the number of records that share an (a, b) pair, and one gather writes the records out.
#include "ovl_ctx.h"

#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ovl_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return OVLB_ERR_CUDA; } } while (0)

static inline unsigned div_up64(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

#define ING_FLAGS  OVL_ING_FLAGS
#define ING_DROP   0xFFFFFFFFFFFFFFFFull           // key of a record that no longer carries a flag (a_iid is never 2^32-1 ... and b)

//  one thread per input record: both twins -> tmp[2i], tmp[2i+1] with their sort keys; out[0] += kept, out[1] |= bad IDs
__global__ void __launch_bounds__(256)
k_ingest_mirror(const ovlb_record *__restrict__ in, uint64_t n, uint32_t max_evalue, uint32_t max_id,
                ovlb_record *__restrict__ tmp, uint32_t *__restrict__ idx, unsigned long long *out) {
  __shared__ unsigned int blk_kept;
  if (threadIdx.x == 0) blk_kept = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int kept = 0;
  if (i < n) {
    const ovlb_record f = in[i];
    if (f.a_iid == 0 || f.b_iid == 0 || f.a_iid > max_id || f.b_iid > max_id) atomicOr(&out[1], 1ull);
    ovlb_record ff = f, r;
    r.a_iid = f.b_iid; r.b_iid = f.a_iid;
    const bool keep = ovl_ingest_twin(f.dat0, f.dat1, max_evalue, &ff.dat0, &r.dat0, &r.dat1) != 0;   // the twin carries the same flags
    tmp[2 * i] = ff; tmp[2 * i + 1] = r;
    idx[2 * i] = (uint32_t)(2 * i); idx[2 * i + 1] = (uint32_t)(2 * i + 1);
    kept = keep ? 2u : 0u;
  }
  if (kept) atomicAdd(&blk_kept, kept);
  __syncthreads();
  if (threadIdx.x == 0 && blk_kept) atomicAdd(&out[0], (unsigned long long)blk_kept);
}

__global__ void __launch_bounds__(256)
k_ingest_gather(const ovlb_record *__restrict__ tmp, const uint32_t *__restrict__ order, uint64_t n, ovlb_record *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tmp[order[i]];
}

//  sort key `which` (0: dat1, 1: dat0, 2: a_iid << 32 | b_iid, all-ones for a record without flags) of record perm[i]
__global__ void __launch_bounds__(256)
k_ingest_key(const ovlb_record *__restrict__ tmp, const uint32_t *__restrict__ perm, uint64_t n, int which, uint64_t *__restrict__ key) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ovlb_record r = tmp[perm[i]];
  key[i] = which == 0 ? r.dat1 : which == 1 ? r.dat0 : ((r.dat0 & ING_FLAGS) ? (((uint64_t)r.a_iid << 32) | r.b_iid) : ING_DROP);
}

int ovl_ingest_records(ovlb_ctx *c, const ovlb_record *in, uint64_t n, uint32_t max_evalue, uint32_t max_id,
                       ovlb_record *out, uint64_t out_cap, uint64_t *n_out) {
  *n_out = 0;
  if (n == 0) return OVLB_OK;
  if (2 * n >= 0xFFFFFFF0ull) { ovl_set_error("ovlb_ingest_records: more than 2^31 records in one call; split the input"); return OVLB_ERR_CAPACITY; }
  const uint64_t m = 2 * n;
  ovlb_record *d_in = nullptr, *d_tmp = nullptr, *d_out = nullptr;
  uint64_t *d_key = nullptr, *d_key2 = nullptr; uint32_t *d_idx = nullptr, *d_idx2 = nullptr; void *d_cub = nullptr;
  auto cleanup = [&]() { void *p[] = { d_in, d_tmp, d_out, d_key, d_key2, d_idx, d_idx2, d_cub }; for (void *q : p) if (q) cudaFree(q); };
#define CKF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ovl_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); cleanup(); return OVLB_ERR_CUDA; } } while (0)
  CKF(cudaMalloc((void **)&d_in, n * sizeof(ovlb_record)));
  CKF(cudaMalloc((void **)&d_tmp, m * sizeof(ovlb_record)));
  CKF(cudaMalloc((void **)&d_key, m * 8)); CKF(cudaMalloc((void **)&d_key2, m * 8));
  CKF(cudaMalloc((void **)&d_idx, m * 4)); CKF(cudaMalloc((void **)&d_idx2, m * 4));
  CKF(cudaMemcpyAsync(d_in, in, n * sizeof(ovlb_record), cudaMemcpyHostToDevice, c->stream));
  CKF(cudaMemsetAsync(&c->d_work[5], 0, 16, c->stream));
  k_ingest_mirror<<<div_up64(n, 256), 256, 0, c->stream>>>(d_in, n, max_evalue, max_id, d_tmp, d_idx, &c->d_work[5]);
  c->launches++;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key, d_key2, d_idx, d_idx2, (int64_t)m, 0, 64, c->stream);
  CKF(cudaMalloc(&d_cub, tb + 256));
  for (int which = 0; which < 3; which++) {                             // least significant key first; every sort is stable
    k_ingest_key<<<div_up64(m, 256), 256, 0, c->stream>>>(d_tmp, d_idx, m, which, d_key);
    size_t tb2 = tb;
    CKF(cub::DeviceRadixSort::SortPairs(d_cub, tb2, d_key, d_key2, d_idx, d_idx2, (int64_t)m, 0, 64, c->stream));
    std::swap(d_idx, d_idx2);
    c->launches += 11;
  }
  unsigned long long h[2] = {0, 0};
  CKF(cudaMemcpyAsync(h, &c->d_work[5], 16, cudaMemcpyDeviceToHost, c->stream));
  CKF(cudaStreamSynchronize(c->stream));
  if (h[1]) { ovl_set_error("ovlb_ingest_records: Overlap has IDs out of range (maxID " + std::to_string(max_id) + "), possibly corrupt input data."); cleanup(); return OVLB_ERR_ARG; }
  const uint64_t kept = h[0];
  *n_out = kept;
  if (kept > out_cap) { ovl_set_error("ovlb_ingest_records: output buffer too small"); cleanup(); return OVLB_ERR_CAPACITY; }
  if (kept) {
    CKF(cudaMalloc((void **)&d_out, kept * sizeof(ovlb_record)));
    k_ingest_gather<<<div_up64(kept, 256), 256, 0, c->stream>>>(d_tmp, d_idx, kept, d_out);
    c->launches++;
    CKF(cudaMemcpyAsync(out, d_out, kept * sizeof(ovlb_record), cudaMemcpyDeviceToHost, c->stream));
    CKF(cudaStreamSynchronize(c->stream));
  }
  CKF(cudaGetLastError());
  cleanup();
#undef CKF
  return OVLB_OK;
}
