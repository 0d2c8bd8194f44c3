//  Microbenchmark: random 32-byte (one sector) gathers from a table much larger than L2, to find
//  what the B200 memory system gives a k-mer probe, and whether cudaLimitMaxL2FetchGranularity matters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(32) Slot { uint64_t a, b, c, d; };
template <int MODE>
__global__ void gather(const Slot *__restrict__ t, uint64_t n_slots, uint64_t n_loads, uint64_t *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (; i < n_loads; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t h = __umul64hi(i * 0x9E3779B97F4A7C15ull + 0x1234567, n_slots);
    uint64_t a, b, c, d;
    if (MODE == 0)      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h));
    else if (MODE == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h));
    else if (MODE == 2) asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h));
    else if (MODE == 3) { asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(t + h)); b = c = d = 0; }
    else if (MODE == 4) { asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h)); }
    else if (MODE == 5) { asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h)); }
    else if (MODE == 6) { asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(t + h)); c = d = 0; }
    else if (MODE == 7) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(t + h)); c = d = 0; }
    else if (MODE == 8) { asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(t + h)); c = d = 0; }
    else                { asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + h)); }
    acc += a ^ b ^ c ^ d;
  }
  if (acc == 0x123456789) out[0] = acc;
}
template <int MODE> float run(const Slot *t, uint64_t n_slots, uint64_t n_loads, uint64_t *out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<MODE><<<148 * 16, 256>>>(t, n_slots, n_loads / 8, out);
  cudaEventRecord(e0);
  gather<MODE><<<148 * 16, 256>>>(t, n_slots, n_loads, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  const uint64_t n_slots = (4ull << 30) / 32, n_loads = 500000000ull;
  Slot *t; uint64_t *out;
  cudaMalloc(&t, n_slots * 32); cudaMalloc(&out, 8); cudaMemset(t, 1, n_slots * 32);
  const char *names[] = {"ld.global.nc.v4.u64", "ld.global.v4.u64", "ld.global.L1::no_allocate.v4.u64", "ld.global.nc.u64 (8 B)", "ld.global.cs.v4.u64", "ld.global.cg.v4.u64", "ld.volatile.global.v2.u64", "ld.relaxed.gpu.global.v2.u64", "ld.global.cv.v2.u64", "ld.global.L1::no_allocate.L2::evict_first.v4"};
  size_t lims[] = {0, 32};
  for (size_t lim : lims) {
    if (lim) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, lim); printf("set limit %zu -> %s\n", lim, cudaGetErrorString(e)); }
    size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now %zu\n", cur);
    float ms[10] = { run<0>(t, n_slots, n_loads, out), run<1>(t, n_slots, n_loads, out), run<2>(t, n_slots, n_loads, out), run<3>(t, n_slots, n_loads, out), run<4>(t, n_slots, n_loads, out), run<5>(t, n_slots, n_loads, out), run<6>(t, n_slots, n_loads, out), run<7>(t, n_slots, n_loads, out), run<8>(t, n_slots, n_loads, out), run<9>(t, n_slots, n_loads, out) };
    for (int m = 0; m < 10; m++) printf("  %-36s %8.3f ms  %7.1f G loads/s  %7.1f GB/s of 32 B sectors\n", names[m], ms[m], n_loads / ms[m] / 1e6, n_loads * 32.0 / ms[m] / 1e6);
  }
  return 0;
}
