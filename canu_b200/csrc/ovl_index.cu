//  ovl_index.cu -- read encoding, k-mer index build, lookup and seed-run emission, chaining.
//
//  Kernels (SURVEY.md 2.1 numbering):
//    K0  encode_fwd / apply_n / encode_rc       sqStore 2-bit bytes -> dp4 words, both orientations
//    K1  hash_tuples_compact -> radix partition -> bucket_group -> path sort -> path_slots
//                                               Build_Hash_Index + Put_String_In_Hash + Hash_Insert + chain coalescing
//                                               (overlapInCore-Build_Hash_Index.C:267-404,415-631): one 32-byte slot per
//                                               distinct k-mer in PATH order, class-partitioned occurrence lists, a
//                                               hash table k-mer -> slot.  Fallback (a bucket that does not fit shared
//                                               memory): hash_tuples -> full radix sort -> count_distinct -> group_heads
//    K1b index_skip                             Mark_Skip_Kmers / Hash_Mark_Empty / Mark_Screened_Ends (:98-257)
//    K2a ref_probe                              Find_Overlaps window loop + Hash_Find (Find_Overlaps.C:177-336):
//                                               speculative coalesced slot loads along the path, hash table at breaks
//    K2b expand_small / expand_large            chain walk + Add_Ref, collapsed to maximal diagonal runs
//    K3  pair_heads / pair_scatter / chain_pairs / pair_cost
//                                               Add_Match replay, hopeless check, --minkmers filter
//                                               (Find_Overlaps.C:26-163, Process_String_Overlaps.C:384-415,581-637),
//                                               heaviest-first order of the pairs for the extension kernel
//
//  All integer / byte work: coalesced streaming of dp4 words and tuples, shared-memory grouping, 32-byte slot loads,
//  warp- and block-aggregated appends.  No tensor cores.
#include "ovl_ctx.h"

#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>

#define WARPS_PER_BLOCK 8
#define THREADS (WARPS_PER_BLOCK * 32)

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
//  K0: encoding
// ------------------------------------------------------------------------------------------------

//  one warp per read: packed 2-bit (4 bases/byte, first base in the top bits) -> dp4 forward words
__global__ void __launch_bounds__(THREADS)
k_encode_fwd(const uint8_t *__restrict__ packed, const uint64_t *__restrict__ boff, const uint32_t *__restrict__ len,
             const uint64_t *__restrict__ woff, uint64_t *__restrict__ fwd, uint32_t n, const uint32_t *__restrict__ src_len) {
  uint32_t r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  if (src_len && src_len[r]) return;                    // a stored blob: k_encode_raw prepares it
  const int lane = threadIdx.x & 31;
  const uint32_t L = len[r];
  const uint32_t nw = (L + 15) / 16 + 2;                  // two zero pad words after every read
  const uint8_t *src = packed + boff[r];
  const uint32_t nbytes = (L + 3) / 4;
  uint64_t *dst = fwd + woff[r];
  for (uint32_t w = lane; w < nw; w += 32) {
    uint64_t out = 0;
    uint32_t b0 = 4 * w;
    if (b0 < nbytes) {
      uint32_t bytes = 0;                                  // bytes[k] = src[b0+k]
      #pragma unroll
      for (int k = 0; k < 4; k++)
        if (b0 + k < nbytes) bytes |= (uint32_t)src[b0 + k] << (8 * k);
      #pragma unroll
      for (int k = 0; k < 16; k++) {
        uint32_t pos = 16 * w + k;
        if (pos < L) {
          uint32_t code = (bytes >> (8 * (k >> 2) + 6 - 2 * (k & 3))) & 3u;
          out |= (uint64_t)(1u << code) << (4 * k);
        }
      }
    }
    dst[w] = out;
  }
}

//  One warp per read uploaded AS STORED (ovlb_reads.src_len > 0): homopolymer compression and clear-range trimming on the
//  device (what sqStore does on the host when it loads a read: decode, homopolyCompress -- utility/src/sequence/
//  sequence-v1.C:203-261 -- then the clear range, stores/sqStore.H:397-413), straight into dp4 words.  A lane takes 16
//  consecutive stored bases (one 32-bit load), keeps the first base of every run (the previous lane's last base arrives
//  by shuffle), a warp scan gives every kept base its position in the compressed read, the lane compacts its kept bases
//  into one-hot nibbles and ORs them into at most two output words.  `fwd` must be zero where the read goes.
__global__ void __launch_bounds__(THREADS)
k_encode_raw(const uint8_t *__restrict__ packed, const uint64_t *__restrict__ boff, const uint32_t *__restrict__ len,
             const uint64_t *__restrict__ woff, uint64_t *fwd, uint32_t n,
             const uint32_t *__restrict__ src_len, const uint32_t *__restrict__ clear_bgn, int hpc) {
  const uint32_t r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  const uint32_t S = src_len[r];
  if (S == 0) return;
  const int lane = threadIdx.x & 31;
  const int L = (int)len[r];
  const int cb = clear_bgn ? (int)clear_bgn[r] : 0;
  const uint8_t *src = packed + boff[r];
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(fwd + woff[r]);
  int c_base = 0;                                       // compressed position of the first base of this round
  uint32_t prev_code = 4;                               // last base of the previous round (4 = none)
  for (uint32_t i0 = 0; i0 < S; i0 += 512) {
    const uint32_t b0 = i0 + 16 * lane;
    const int nv = b0 < S ? (int)min(16u, S - b0) : 0;
    uint32_t codes = 0;                                 // base j of the lane in bits [2j, 2j+1]
    if (nv > 0) {
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t byte = (b0 >> 2) + k;
        const uint32_t v = (byte < ((S + 3) >> 2)) ? src[byte] : 0u;
        codes |= ((v >> 6) & 3u) << (8 * k) | ((v >> 4) & 3u) << (8 * k + 2) | ((v >> 2) & 3u) << (8 * k + 4) | (v & 3u) << (8 * k + 6);
      }
    }
    const uint32_t my_last = nv > 0 ? (codes >> (2 * (nv - 1))) & 3u : 4u;
    uint32_t before = __shfl_up_sync(0xffffffffu, my_last, 1);
    if (lane == 0) before = prev_code;
    //  keep mask: every base when not compressing, else the bases that differ from their predecessor
    uint32_t keep = 0;
    {
      uint32_t p = before;
      #pragma unroll
      for (int j = 0; j < 16; j++) {
        const uint32_t cj = (codes >> (2 * j)) & 3u;
        if (j < nv && (!hpc || cj != p)) keep |= 1u << j;
        p = cj;
      }
    }
    const int cnt = __popc(keep);
    int inc = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    int o0 = c_base + inc - cnt - cb;                  // output position of my first kept base
    //  compact the kept bases into one-hot nibbles
    unsigned long long str = 0; int k2 = 0;
    #pragma unroll
    for (int j = 0; j < 16; j++)
      if ((keep >> j) & 1u) {
        const int o = o0 + k2;
        if (o >= 0 && o < L) str |= (unsigned long long)(1u << ((codes >> (2 * j)) & 3u)) << (4 * k2);
        k2++;
      }
    if (str) {
      //  nibble k2 of `str` belongs at output position o0 + k2 (positions outside [0, L) hold zero nibbles)
      if (o0 < 0) { str >>= 4 * (-o0); o0 = 0; }
      const int w = o0 >> 4, sh = (o0 & 15) << 2;
      if (str << sh) atomicOr(&dst[w], str << sh);
      if (sh && (str >> (64 - sh))) atomicOr(&dst[w + 1], str >> (64 - sh));
    }
    c_base += total;
    //  last base of the round: the last lane that had bases
    const unsigned has = __ballot_sync(0xffffffffu, nv > 0);
    prev_code = __shfl_sync(0xffffffffu, my_last, 31 - __clz(has));
  }
}

__global__ void k_apply_n(const uint32_t *__restrict__ nread, const uint32_t *__restrict__ npos, uint64_t nn,
                          const uint64_t *__restrict__ woff, const uint32_t *__restrict__ len, uint64_t *fwd) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  uint32_t r = nread[i], p = npos[i];
  if (p >= len[r]) return;
  atomicOr((unsigned long long *)&fwd[woff[r] + (p >> 4)], 0xFull << ((p & 15) << 2));
}

//  one warp per read: reverse complement = 64-bit bit reversal of the mirrored window
__global__ void __launch_bounds__(THREADS)
k_encode_rc(const uint64_t *__restrict__ fwd, const uint32_t *__restrict__ len, const uint64_t *__restrict__ woff,
            uint64_t *__restrict__ rc, uint32_t n) {
  uint32_t r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const int L = (int)len[r];
  const uint32_t nw = (L + 15) / 16 + 2;
  const uint64_t *src = fwd + woff[r];
  uint64_t *dst = rc + woff[r];
  for (uint32_t w = lane; w < nw; w += 32) {
    uint64_t v = 0;
    int x = L - 16 - 16 * (int)w;
    if (16 * (int)w < L) {
      if (x >= 0) v = ovl_fetch16(src, x);
      else        v = src[0] << (4 * (-x));
    }
    dst[w] = __brevll(v);
  }
}

//  one warp per read: read index of every 32-position group
__global__ void __launch_bounds__(THREADS)
k_fill_groups(const uint64_t *__restrict__ pbase, uint32_t *__restrict__ grp_read, uint32_t n) {
  uint32_t r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  uint64_t g0 = pbase[r] >> 5, g1 = pbase[r + 1] >> 5;
  for (uint64_t g = g0 + lane; g < g1; g += 32) grp_read[g] = r;
}

// ------------------------------------------------------------------------------------------------
//  k-mers of the 32 windows of one position group, computed by the whole warp.
//
//  Window p0+lane of the read whose dp4 words start at w (p0 a multiple of 32).  Lanes 0..4 load the five
//  words that hold bases p0-16 .. p0+63 and convert them to 2-bit codes once; every lane then assembles its
//  own 64-bit k-mer from three of them with two funnel shifts.  Returns false if the window runs past the
//  end of the read or holds a base that is not ACGT.  `cls` is the class of the PRECEDING base:
//  0 = none (start of read, or N), 1..4 = A,C,G,T.
// ------------------------------------------------------------------------------------------------
//  the word lane `lane` (< 5) contributes to the k-mers of group p0: split from the arithmetic so that a caller can have
//  the loads of many groups in flight at once
__device__ __forceinline__ uint64_t warp_kmers_load(const uint64_t *__restrict__ w, int p0, int lane, bool &have) {
  const int idx = (p0 >> 4) - 1 + lane;
  have = lane < 5 && idx >= 0;
  return have ? w[idx] : 0ull;
}
__device__ __forceinline__ bool warp_kmers_from(uint64_t word, bool have, int p0, int L, int K, int lane, uint64_t &key, int &cls);

__device__ __forceinline__ bool warp_kmers(const uint64_t *__restrict__ w, int p0, int L, int K, int lane,
                                           uint64_t &key, int &cls) {
  bool have;
  const uint64_t word = warp_kmers_load(w, p0, lane, have);
  return warp_kmers_from(word, have, p0, L, K, lane, key, cls);
}

__device__ __forceinline__ bool warp_kmers_from(uint64_t word, bool have, int p0, int L, int K, int lane, uint64_t &key, int &cls) {
  uint32_t codes = 0, inv = 0xFFFFu;
  if (have) codes = ovl_codes16(word, &inv);
  const int j0 = 1 + (lane >> 4);                      // first of my three words (word 0 = the one before p0)
  const uint32_t cp = __shfl_sync(0xffffffffu, codes, j0 - 1), ip = __shfl_sync(0xffffffffu, inv, j0 - 1);
  const uint32_t c0 = __shfl_sync(0xffffffffu, codes, j0),     i0 = __shfl_sync(0xffffffffu, inv, j0);
  const uint32_t c1 = __shfl_sync(0xffffffffu, codes, j0 + 1), i1 = __shfl_sync(0xffffffffu, inv, j0 + 1);
  const uint32_t c2 = __shfl_sync(0xffffffffu, codes, j0 + 2), i2 = __shfl_sync(0xffffffffu, inv, j0 + 2);
  const int s = lane & 15;
  const uint32_t lo = __funnelshift_r(c0, c1, 2 * s);
  const uint32_t hi = __funnelshift_r(c1, c2, 2 * s);
  key = (((uint64_t)hi << 32) | lo) & ((1ull << (2 * K)) - 1);
  const uint64_t iv = ((uint64_t)i0 | ((uint64_t)i1 << 16) | ((uint64_t)i2 << 32)) >> s;
  const int p = p0 + lane;
  uint32_t pc, pi;
  if (s > 0) { pc = (c0 >> (2 * (s - 1))) & 3u; pi = (i0 >> (s - 1)) & 1u; }
  else       { pc = cp >> 30;                   pi = ip >> 15; }
  cls = (p > 0 && pi == 0) ? (int)(1 + pc) : 0;
  return (p + K <= L) && (((uint32_t)iv & ((1u << K) - 1)) == 0);
}

//  The same for a warp that walks consecutive groups of one read (the probe kernel): the 2-bit codes of 32 consecutive
//  dp4 words (512 bases) are converted once and kept one word per lane; a group needs five of them, so a window serves
//  14 groups before it is reloaded -- one coalesced 256-byte load and one conversion per 14 groups instead of a
//  40-byte load and a conversion per group.  `W` caches (read pointer, first word index, codes, invalid mask).
struct KmerWindow { const uint64_t *w; int w0; uint32_t codes, inv; };

__device__ __forceinline__ bool warp_kmers_win(KmerWindow &W, const uint64_t *__restrict__ w, int p0, int L, int K, int lane,
                                               uint64_t &key, int &cls) {
  const int need = (p0 >> 4) - 1;                       // first of the five words of this group (may be -1)
  if (W.w != w || need < W.w0 || need + 4 > W.w0 + 31) {
    W.w = w; W.w0 = need;
    const int idx = need + lane;
    W.codes = 0; W.inv = 0xFFFFu;
    if (idx >= 0) W.codes = ovl_codes16(w[idx], &W.inv);
  }
  const int sb = need - W.w0;
  const int j0 = sb + 1 + (lane >> 4);                 // first of my three words
  const uint32_t cp = __shfl_sync(0xffffffffu, W.codes, j0 - 1), ip = __shfl_sync(0xffffffffu, W.inv, j0 - 1);
  const uint32_t c0 = __shfl_sync(0xffffffffu, W.codes, j0),     i0 = __shfl_sync(0xffffffffu, W.inv, j0);
  const uint32_t c1 = __shfl_sync(0xffffffffu, W.codes, j0 + 1), i1 = __shfl_sync(0xffffffffu, W.inv, j0 + 1);
  const uint32_t c2 = __shfl_sync(0xffffffffu, W.codes, j0 + 2), i2 = __shfl_sync(0xffffffffu, W.inv, j0 + 2);
  const int s = lane & 15;
  const uint32_t lo = __funnelshift_r(c0, c1, 2 * s);
  const uint32_t hi = __funnelshift_r(c1, c2, 2 * s);
  key = (((uint64_t)hi << 32) | lo) & ((1ull << (2 * K)) - 1);
  const uint64_t iv = ((uint64_t)i0 | ((uint64_t)i1 << 16) | ((uint64_t)i2 << 32)) >> s;
  const int p = p0 + lane;
  uint32_t pc, pi;
  if (s > 0) { pc = (c0 >> (2 * (s - 1))) & 3u; pi = (i0 >> (s - 1)) & 1u; }
  else       { pc = cp >> 30;                   pi = ip >> 15; }
  cls = (p > 0 && pi == 0) ? (int)(1 + pc) : 0;
  return (p + K <= L) && (((uint32_t)iv & ((1u << K) - 1)) == 0);
}

// ------------------------------------------------------------------------------------------------
//  K1: index build = tuple generation -> radix sort -> one slot per distinct k-mer
//
//  Replaces Put_String_In_Hash / Hash_Insert / chain coalescing (Build_Hash_Index.C:267-404,613-628).
//  Sorting (k-mer, class of the preceding base) groups the occurrences of a k-mer contiguously AND splits
//  each group by the base that precedes the occurrence in its read; the lookup side needs exactly that to
//  find the heads of diagonal runs without touching the occurrences that merely continue a run.
// ------------------------------------------------------------------------------------------------

//  one warp per 32-position group of the hash block
__global__ void __launch_bounds__(THREADS)
k_hash_tuples(const uint64_t *__restrict__ fwd, const uint64_t *__restrict__ woff, const uint32_t *__restrict__ len,
              const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read, uint64_t n_groups,
              int K, uint64_t *__restrict__ tkey, uint32_t *__restrict__ tval) {
  uint64_t g = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= n_groups) return;
  const int lane = threadIdx.x & 31;
  const uint32_t r = grp_read[g];
  const int L = (int)len[r];
  const int p0 = (int)(g * 32 - pbase[r]);
  uint64_t key; int cls;
  const bool ok = warp_kmers(fwd + woff[r], p0, L, K, lane, key, cls);
  tkey[g * 32 + lane] = ok ? ((key << 3) | (uint64_t)cls) : (1ull << (2 * K + 3));     // invalid windows sort last
  tval[g * 32 + lane] = (uint32_t)(g * 32 + lane);
}

// ------------------------------------------------------------------------------------------------
//  Bucketed index build.  The tuples (mixed k-mer << 3 | class, position) are cut into buckets of ~BK_TARGET by the top
//  B bits of the key (two partition kernels below); one CTA then finishes a bucket entirely in shared memory: a hash
//  table gives every distinct k-mer of the bucket a slot and counts its occurrences per class of the preceding base, a
//  prefix sum turns the counts into offsets, and the positions are scattered to their (k-mer, class) segment.  That is
//  everything the index needs -- the order of the k-mers inside a bucket and of the positions inside a class is
//  immaterial -- so no sort, no distinct count and no class-boundary searches (k_count_distinct, k_group_heads of the
//  sorted build) are run at all.  Two CTAs of 512 threads share an SM (109 KB of shared memory each): one CTA's
//  barrier-separated phases overlap the other's.
// ------------------------------------------------------------------------------------------------
#define BK_THREADS 512
#define BK_CAP     3072                 // tuples of a bucket that fit shared memory
#define BK_PER     (BK_CAP / BK_THREADS)     // tuples per thread, staged in registers
#define BK_TARGET  2048                 // mean bucket size the bucket count is chosen for
//  Hash-table entries per bucket = 2^TB, a template parameter of the bucket kernels: TB = 11 (2048 entries, <= 1792
//  distinct k-mers, 109 KB of shared memory, two CTAs per SM) serves blocks that cover their genome many times; a block
//  of a LARGE job covers it a few times at most and nearly every tuple of a bucket is a distinct k-mer: TB = 12 (4096
//  entries >= BK_CAP, 170 KB, one CTA per SM) can never run out of entries.
#define BK_TABLE_OF(TB)   (1 << (TB))
#define BK_MAXDIST_OF(TB) ((1 << (TB)) / 8 * 7)

//  counts distinct k-mers and finds the number of valid tuples in the sorted array
#define CNT_PER_THREAD 8
__global__ void __launch_bounds__(256)
k_count_distinct(const uint64_t *__restrict__ skey, uint64_t n, uint64_t sentinel, unsigned long long *out /* [0] distinct, [1] n_occ */) {
  __shared__ unsigned int blk;
  if (threadIdx.x == 0) blk = 0;
  __syncthreads();
  const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * CNT_PER_THREAD;
  unsigned int c = 0;
  uint64_t prev = (base > 0 && base <= n) ? skey[base - 1] : ~0ull;
  #pragma unroll
  for (int j = 0; j < CNT_PER_THREAD; j++) {
    const uint64_t i = base + j;
    if (i < n) {
      const uint64_t k = skey[i];
      if (k < sentinel) {
        if (i == 0 || (k >> 3) != (prev >> 3)) c++;
        if (i + 1 == n) out[1] = i + 1;
      } else if (i == 0) {
        out[1] = 0;
      } else if (prev < sentinel) {
        out[1] = i;
      }
      prev = k;
    }
  }
  if (c) atomicAdd(&blk, c);
  __syncthreads();
  if (threadIdx.x == 0 && blk) atomicAdd(&out[0], (unsigned long long)blk);
}

//  first index in [lo, n) whose key is >= target, galloping from lo (all keys before lo are < target)
__device__ __forceinline__ uint32_t gallop_lower_bound(const uint64_t *__restrict__ k, uint32_t lo, uint32_t n, uint64_t target) {
  uint32_t hi = lo, step = 1;
  while (hi < n && k[hi] < target) {
    lo = hi + 1;
    hi = (n - hi > step) ? hi + step : n;
    step <<= 1;
  }
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (k[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

//  ------------------------------------------------------------------------------------------------
//  The index proper: PATH-ORDERED slots + a small hash table that finds a k-mer's slot.
//
//  One 32-byte IndexSlot per distinct k-mer, stored in the order in which the k-mers first occur in the hash block
//  (position of the first occurrence, ascending).  Consecutive windows of a read therefore sit in consecutive slots
//  wherever that read -- or any read covering the same stretch without an error -- was the first to bring those
//  k-mers in, so the lookup of 32 consecutive ref windows is normally ONE 1 KB coalesced load (8 lines) instead of
//  32 random probes that cost a 128-byte line each (tools/micro/rand_sector.cu: 36.9 G random lines/s is all B200
//  gives, whatever the load width).  `htab` (below) is only consulted where the path breaks: a read start, an error, the
//  border between two first-coverage stretches -- and for every window whose k-mer is not in the block at all.
//  ------------------------------------------------------------------------------------------------
//  reverse complement of a k-mer key (base j in bits 2j..2j+1, A0 C1 G2 T3): complement = ~code, order reversed
__device__ __forceinline__ uint64_t kmer_rc(uint64_t key, int K) {
  uint64_t x = __brevll(~key) >> (64 - 2 * K);                            // pairs reversed, bits inside a pair swapped
  return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

#define HT_EMPTY   0xFFFFFFFFFFFFFFFFull
#define HT_NOTFOUND 0xFFFFFFFFu
#define HT_MULT    0x9E3779B97F4A7C15ull

//  `htab`: buckets of four 8-byte entries = one 32-byte sector, the unit DRAM and L2 move anyway.  An entry is
//  (fingerprint << 32 | slot index); the k-mer itself is only in the slot, which a lookup reads anyway and which
//  confirms the match.  hcap = 4 x (distinct k-mers): a quarter of the entries are used, 1.9 % of the buckets are full,
//  so a k-mer that is NOT in the table -- most windows of a ref read from a part of the genome the hash block does not
//  cover, and nearly all reverse-strand windows at low coverage -- costs ONE 32-byte load and no divergent probe loop
//  (the 16-byte-entry table with linear probing it replaces: 2.5 dependent probes per miss, 124 DRAM bytes per window,
//  16.9 of 32 lanes active; profiles/r2l_ncu_probe_sparse.txt).  Same 32 bytes per distinct k-mer as before.
//  A key lives in the first bucket of its probe sequence (home, home + 1, ...) that had a free entry when it was
//  inserted; entries are never removed, so a lookup may stop at the first bucket that still has a free entry.
//  The bucket and the fingerprint come from the CANONICAL k-mer (the smaller of the k-mer and its reverse complement):
//  a k-mer and its reverse complement are different keys with different slots, but they share a probe sequence, so the
//  one bucket load that looks a forward window up also says whether the reverse-strand window over the same bases can
//  be in the table (slot_lookup_fr below) -- half the random loads of a job whose windows mostly miss.
__device__ __forceinline__ uint64_t ht_hash_of(uint64_t key, uint64_t rkey) { return (key < rkey ? key : rkey) * HT_MULT; }
__device__ __forceinline__ uint64_t ht_bucket_of(uint64_t h, uint64_t hcap) { return __umul64hi(h, hcap >> 2); }
//  fingerprint = low bits of the multiplicative hash (the bucket comes from its high bits).  The mask is all ones except
//  in tests, which narrow it (OVLB_HT_FPMASK) so that different k-mers of a bucket share fingerprints and the
//  confirm-by-slot-and-go-on path is exercised.
__device__ uint32_t g_ht_fpmask = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t ht_fp_of(uint64_t h) { return (uint32_t)h & g_ht_fpmask; }
//  ... and the lowest fingerprint bit says which of the two this entry is (0 = the canonical form), so that a lookup
//  does not have to read the slot of the other strand's entry to tell them apart
__device__ __forceinline__ uint32_t ht_fp2(uint64_t h, uint64_t key, uint64_t rkey) { return (ht_fp_of(h) & ~1u) | (key > rkey ? 1u : 0u); }

template <bool NC>
__device__ __forceinline__ void ht_load_bucket(const HashEntry *ht, uint64_t b, uint64_t (&e)[4]) {
  if (NC) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(ht + 4 * b));
  else    asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(ht + 4 * b) : "memory");
}

//  slot index of `key`, or HT_NOTFOUND.  NC = the table and the slots are read-only for the duration of the kernel.
template <bool NC>
__device__ __forceinline__ uint32_t ht_find(const HashEntry *ht, uint64_t hcap, const IndexSlot *slots, uint64_t key, int K) {
  const uint64_t rkey = kmer_rc(key, K);
  const uint64_t h = ht_hash_of(key, rkey), nb = hcap >> 2;
  const uint32_t fp = ht_fp2(h, key, rkey);
  for (uint64_t b = ht_bucket_of(h, hcap);;) {
    uint64_t e[4];
    ht_load_bucket<NC>(ht, b, e);
    bool open = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (e[i] == HT_EMPTY) { open = true; continue; }
      if ((uint32_t)(e[i] >> 32) != fp) continue;
      const uint64_t *kp = &slots[(uint32_t)e[i]].key;
      const uint64_t k = NC ? __ldg(kp) : *(const volatile uint64_t *)kp;
      if ((k & ~OVL_SKIP_BIT) == key) return (uint32_t)e[i];
    }
    if (open) return HT_NOTFOUND;
    if (++b == nb) b = 0;
  }
}

//  claims an entry for `key` (which must not be in the table yet, or be inserted by nobody else concurrently)
__device__ __forceinline__ void ht_insert(HashEntry *ht, uint64_t hcap, uint64_t key, uint32_t idx, int K) {
  const uint64_t rkey = kmer_rc(key, K);
  const uint64_t h = ht_hash_of(key, rkey), nb = hcap >> 2;
  const unsigned long long val = ((unsigned long long)ht_fp2(h, key, rkey) << 32) | idx;
  for (uint64_t b = ht_bucket_of(h, hcap);;) {
    unsigned long long *p = reinterpret_cast<unsigned long long *>(ht + 4 * b);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (*(volatile unsigned long long *)(p + i) == HT_EMPTY && atomicCAS(p + i, (unsigned long long)HT_EMPTY, val) == HT_EMPTY) return;
    if (++b == nb) b = 0;
  }
}

//  one thread per sorted tuple; the first tuple of every k-mer emits the k-mer's slot (unordered, compacted) together
//  with the position of the k-mer's first occurrence, the key of the path order
__global__ void __launch_bounds__(256)
k_group_heads(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ occ, uint32_t n_occ,
              IndexSlot *__restrict__ tmp, uint32_t *__restrict__ gk, uint32_t *__restrict__ gv, unsigned long long *counter) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool head = false;
  uint64_t k = 0;
  if (i < n_occ) {
    k = skey[i];
    head = (i == 0) || ((skey[i - 1] >> 3) != (k >> 3));
  }
  uint32_t e[5]; uint32_t minpos = 0xFFFFFFFFu;
  if (head) {
    const uint64_t kmer = k >> 3;
    uint32_t lo = (uint32_t)i;
    #pragma unroll
    for (int c = 0; c < 5; c++) {
      const uint32_t nx = gallop_lower_bound(skey, lo, n_occ, (kmer << 3) + (uint64_t)(c + 1));
      if (nx > lo) { const uint32_t p = occ[lo]; if (p < minpos) minpos = p; }      // occurrences of a class are in position order
      lo = nx; e[c] = nx;
    }
  }
  //  compact index: ONE atomic per block (a warp-level atomic per 32 tuples was 7.8 M same-address atomics per C2 tile,
  //  which alone cost more than the searches)
  __shared__ unsigned int warp_cnt[8];
  __shared__ unsigned long long blk_base;
  const unsigned m = __ballot_sync(0xffffffffu, head);
  if (lane == 0) warp_cnt[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot = 0;
    #pragma unroll
    for (int w = 0; w < 8; w++) { const unsigned int t = warp_cnt[w]; warp_cnt[w] = tot; tot += t; }
    blk_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (head) {
    const uint32_t c = (uint32_t)blk_base + warp_cnt[threadIdx.x >> 5] + __popc(m & ((1u << lane) - 1));
    uint4 *sp = reinterpret_cast<uint4 *>(&tmp[c]);
    const uint64_t kmer = k >> 3;
    sp[0] = make_uint4((uint32_t)kmer, (uint32_t)(kmer >> 32), (uint32_t)i, e[0]);
    sp[1] = make_uint4(e[1], e[2], e[3], e[4]);
    gk[c] = minpos; gv[c] = c;
  }
}

//  one thread per path position: gather the slot into place and publish it in the hash table
__global__ void __launch_bounds__(256)
k_path_slots(const IndexSlot *__restrict__ tmp, const uint32_t *__restrict__ order, uint32_t n, IndexSlot *__restrict__ slots,
             HashEntry *ht, uint64_t hcap, int K) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint4 *src = reinterpret_cast<const uint4 *>(&tmp[order[j]]);
  const uint4 a = src[0], b = src[1];
  uint4 *dst = reinterpret_cast<uint4 *>(&slots[j]);
  dst[0] = a; dst[1] = b;
  ht_insert(ht, hcap, (uint64_t)a.x | ((uint64_t)a.y << 32), j, K);
}

// ------------------------------------------------------------------------------------------------
//  Bucketed build, round 2: no library sort on the way.
//
//    k_part1        k-mers of a 4096-position chunk -> tuples, scattered straight into 2^B1 coarse partitions of the top B1
//                   bits of the mixed key (fixed-stride regions, one global reservation per (chunk, partition)); the
//                   12-byte tuple is written once, already partitioned -- no separate tuple array, no first sort pass
//    k_part2        a 4096-tuple chunk of one coarse partition -> 2^B2 buckets; the tuple shrinks to 8 bytes on the way:
//                   the top B bits are the bucket number, what is left of the key shares a word with the position
//    k_bucket_group2  one bucket per CTA iteration; the bucket's tuples arrive in shared memory by ONE bulk asynchronous
//                   copy (cp.async.bulk + mbarrier: the TMA unit moves them, no thread issues loads), and the copy of
//                   the NEXT bucket is in flight while this one is grouped
//    path order     = rank of the k-mer's first position among all first positions: a bitmap of first positions
//                   (31 MB per 250 M positions, L2 resident), a prefix popcount, and every distinct k-mer computes its
//                   own slot index -- no sort of (position, slot) pairs
//
//  Bytes per hash k-mer (K <= 24): 0.5 read + 12 written (part1), 12 read + 8 written (part2), 8 read + 4 written
//  (bucket_group2) = 44.5, against 72.5 for tuples + two library radix passes + bucket read in round 1.
// ------------------------------------------------------------------------------------------------
#define PART_CH      4096               // tuples per CTA chunk of k_part2
#define PART_THREADS 256
#define PART_MAXD    1024               // digits per level (B1, B2 <= 10)
#define CNT1_STRIDE  32                 // u32 words between two coarse-partition cursors: one 128-byte line each
#define PART1_GROUPS 256                // 32-position groups per CTA of k_part1: one group per THREAD

//  k_part1: one THREAD per 32-position group.  The thread loads the five dp4 words that hold the group's windows itself,
//  converts them to 2-bit codes once and then walks the 32 windows in registers (shift, mask, multiply): no shuffles,
//  no staging in shared memory.  (A first version computed the k-mers warp-wide with shuffles and staged tuples and
//  ranks in shared memory: 19 shared-memory / shuffle / scattered-store operations per 32 tuples against 3 in k_part2,
//  and ncu showed it MIO-bound -- `mio` and `short_scoreboard` stalls, 14.7 % issue active, 9.5 ms.)  Two passes over
//  the windows: count per partition (shared-memory histogram), reserve the CTA's share of every partition with one
//  global atomic per partition, then recompute, take a slot and write the 16-byte tuple record in place.
__device__ __forceinline__ bool group_window(uint64_t lo, uint64_t hi, uint64_t iv, uint32_t prev_code0, uint32_t prev_inv0,
                                             int w, int p0, int L, int K, uint64_t kmask, uint64_t &key, int &cls) {
  const int sh = 2 * w;
  key = (w ? ((lo >> sh) | (hi << (64 - sh))) : lo) & kmask;
  const uint32_t bad = (uint32_t)(iv >> w) & ((1u << K) - 1u);          // K <= 30
  uint32_t pc, pi;
  if (w) { const int b = w - 1; pc = (uint32_t)(lo >> (2 * b)) & 3u; pi = (uint32_t)(iv >> b) & 1u; }
  else   { pc = prev_code0; pi = prev_inv0; }
  const int p = p0 + w;
  cls = (p > 0 && pi == 0) ? (int)(1 + pc) : 0;
  return (p + K <= L) && bad == 0;
}

//  Census only: tuples that do not fit their partition or bucket -- k-mers with thousands to millions of copies, the
//  very ones the census is after -- are counted in a global open-addressed table instead (key = mixed canonical k-mer),
//  and the whole bucket of such a k-mer is diverted there (`bucket_spill`, or its count exceeding the capacity), so
//  that every k-mer is counted in exactly one place.  key == nullptr: not a census (the index build falls back instead).
struct Spill { unsigned long long *key; unsigned int *cnt; uint64_t mask; unsigned int *bucket_spill; int sh2; unsigned long long *flag; };
__device__ __forceinline__ void spill_add(const Spill &S, uint64_t km) {
  uint64_t h = ovl_mix64(km) & S.mask;
  for (int probe = 0; probe < 4096; probe++) {
    const unsigned long long cur = S.key[h];
    if (cur == km) { atomicAdd(&S.cnt[h], 1u); return; }
    if (cur == ~0ull) {
      const unsigned long long old = atomicCAS(&S.key[h], ~0ull, (unsigned long long)km);
      if (old == ~0ull || old == km) { atomicAdd(&S.cnt[h], 1u); return; }
    }
    h = (h + 1) & S.mask;
  }
  atomicOr(S.flag, 4ull);                                                 // table full
}

//  CENSUS = k-mer counting for the skip list (ovl_kmer_census): the tuple key is the CANONICAL k-mer (smaller of the
//  k-mer and its reverse complement), no class, and only the k-mers of slice `slice` of 2^slice_bits (by a hash of the
//  canonical k-mer) are emitted, so that a store of any size is counted in passes over k-mer space.
template <bool CENSUS>
__device__ __forceinline__ bool part1_tuple(uint64_t key, int cls, int K, uint64_t kmask, uint64_t mixc, int slice_bits, uint32_t slice, uint64_t &t) {
  if (CENSUS) {
    const uint64_t rc = kmer_rc(key, K);
    const uint64_t canon = key < rc ? key : rc;
    if (slice_bits && (uint32_t)(ovl_mix64(canon) >> 40 & ((1u << slice_bits) - 1u)) != slice) return false;
    t = ((canon * mixc) & kmask) << 3;
  } else {
    t = (((key * mixc) & kmask) << 3) | (uint64_t)cls;
  }
  return true;
}

template <bool CENSUS>
__global__ void __launch_bounds__(PART_THREADS)
k_part1(const uint64_t *__restrict__ fwd, const uint64_t *__restrict__ woff, const uint32_t *__restrict__ len,
        const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read, uint64_t n_groups,
        int K, uint64_t mixc, int sh1, uint32_t nd1, uint64_t cap1,
        uint4 *__restrict__ trec, unsigned int *cnt1, unsigned long long *flag, int slice_bits, uint32_t slice, Spill S) {
  __shared__ uint32_t hist[PART_MAXD], cursor[PART_MAXD];
  const int tid = threadIdx.x;
  const uint64_t kmask = (1ull << (2 * K)) - 1;
  for (uint32_t d = tid; d < nd1; d += PART_THREADS) hist[d] = 0;
  const uint64_t g = (uint64_t)blockIdx.x * PART1_GROUPS + tid;
  uint64_t lo = 0, hi = 0, iv = ~0ull; uint32_t pc0 = 0, pi0 = 1; int p0 = 0, L = 0;
  if (g < n_groups) {
    const uint32_t r = grp_read[g];
    const uint64_t *w = fwd + woff[r];
    L = (int)len[r];
    p0 = (int)(g * 32 - pbase[r]);
    const int i0 = (p0 >> 4) - 1;                                        // word of the 16 bases before p0
    uint64_t wd[5];
    #pragma unroll
    for (int j = 0; j < 5; j++) wd[j] = (i0 + j >= 0) ? w[i0 + j] : 0ull;   // reads end in two zero pad words: in bounds
    uint32_t c[5], in[5];
    #pragma unroll
    for (int j = 0; j < 5; j++) c[j] = ovl_codes16(wd[j], &in[j]);
    lo = (uint64_t)c[1] | ((uint64_t)c[2] << 32); hi = (uint64_t)c[3] | ((uint64_t)c[4] << 32);
    iv = (uint64_t)in[1] | ((uint64_t)in[2] << 16) | ((uint64_t)in[3] << 32) | ((uint64_t)in[4] << 48);
    pc0 = c[0] >> 30; pi0 = (i0 >= 0) ? (in[0] >> 15) : 1u;
  }
  __syncthreads();
  for (int w = 0; w < 32; w++) {
    uint64_t key; int cls;
    uint64_t t;
    if (group_window(lo, hi, iv, pc0, pi0, w, p0, L, K, kmask, key, cls) && part1_tuple<CENSUS>(key, cls, K, kmask, mixc, slice_bits, slice, t))
      atomicAdd(&hist[(uint32_t)(t >> sh1)], 1u);
  }
  __syncthreads();
  for (uint32_t d = tid; d < nd1; d += PART_THREADS) { const uint32_t h = hist[d]; cursor[d] = h ? atomicAdd(&cnt1[d * CNT1_STRIDE], h) : 0u; }
  __syncthreads();
  bool over = false;
  for (int w = 0; w < 32; w++) {
    uint64_t key; int cls;
    uint64_t t;
    if (group_window(lo, hi, iv, pc0, pi0, w, p0, L, K, kmask, key, cls) && part1_tuple<CENSUS>(key, cls, K, kmask, mixc, slice_bits, slice, t)) {
      const uint32_t d = (uint32_t)(t >> sh1);
      const uint64_t o = atomicAdd(&cursor[d], 1u);
      if (o < cap1) trec[d * cap1 + o] = make_uint4((uint32_t)t, (uint32_t)(t >> 32), CENSUS ? 0u : (uint32_t)(g * 32 + w), 0u);
      else if (CENSUS && S.key) { spill_add(S, t >> 3); S.bucket_spill[(uint32_t)(t >> S.sh2)] = 1u; }
      else over = true;
    }
  }
  if (over) atomicOr(flag, 1ull);
}

#define PART2_THREADS 512
#define PART2_PER     (PART_CH / PART2_THREADS)
__global__ void __launch_bounds__(PART2_THREADS)
k_part2(const uint4 *__restrict__ trec, const unsigned int *__restrict__ cnt1, uint64_t cap1,
        int sh2, int B2, uint64_t lowmask, int posbits, uint32_t bk_cap,
        uint64_t *__restrict__ btup, unsigned int *cnt2, unsigned long long *flag, Spill S) {
  __shared__ uint32_t hist[PART_MAXD], gbase[PART_MAXD];
  const int tid = threadIdx.x;
  const uint32_t p = blockIdx.y;
  const uint64_t n_in = min((uint64_t)cnt1[p * CNT1_STRIDE], cap1);
  const uint64_t start = (uint64_t)blockIdx.x * PART_CH;
  if (start >= n_in) return;
  const uint32_t m = (uint32_t)min((uint64_t)PART_CH, n_in - start);
  const uint32_t nd2 = 1u << B2;
  for (uint32_t d = tid; d < nd2; d += PART2_THREADS) hist[d] = 0;
  __syncthreads();
  uint64_t t[PART2_PER]; uint32_t pos[PART2_PER], rk[PART2_PER];
  const uint4 *src = trec + (uint64_t)p * cap1 + start;
  #pragma unroll
  for (int i = 0; i < PART2_PER; i++) {
    const uint32_t j = tid + i * PART2_THREADS;
    const uint4 v = j < m ? src[j] : make_uint4(0, 0, 0, 0);
    t[i] = (uint64_t)v.x | ((uint64_t)v.y << 32);
    pos[i] = v.z;
  }
  #pragma unroll
  for (int i = 0; i < PART2_PER; i++) {
    const uint32_t j = tid + i * PART2_THREADS;
    if (j < m) rk[i] = atomicAdd(&hist[(uint32_t)(t[i] >> sh2) & (nd2 - 1)], 1u);
  }
  __syncthreads();
  for (uint32_t d = tid; d < nd2; d += PART2_THREADS) { const uint32_t h = hist[d]; gbase[d] = h ? atomicAdd(&cnt2[(p << B2) + d], h) : 0u; }
  __syncthreads();
  bool over = false;
  #pragma unroll
  for (int i = 0; i < PART2_PER; i++) {
    const uint32_t j = tid + i * PART2_THREADS;
    if (j >= m) continue;
    const uint32_t d = (uint32_t)(t[i] >> sh2) & (nd2 - 1);
    const uint32_t o = gbase[d] + rk[i];
    if (o < bk_cap) btup[(uint64_t)((p << B2) + d) * bk_cap + o] = ((t[i] & lowmask) << posbits) | pos[i];
    else if (S.key) spill_add(S, t[i] >> 3);                             // census: the bucket's count now exceeds its capacity, which diverts all of it
    else over = true;
  }
  if (over) atomicOr(flag, 1ull);
}

//  exclusive prefix sum of min(cnt, cap) over nb buckets, one CTA; offs[nb] = total
__global__ void __launch_bounds__(1024)
k_bucket_scan(const unsigned int *__restrict__ cnt, uint32_t nb, uint32_t cap, uint32_t *__restrict__ offs) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t v = i < nb ? min(cnt[i], cap) : 0u;
    uint32_t inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = wsum[lane], wi = w;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t c0 = carry;
    if (i < nb) offs[i] = c0 + wsum[wid] + inc - v;
    __syncthreads();
    if (tid == 1023) carry = c0 + wsum[wid] + inc;
    __syncthreads();
  }
  if (tid == 0) offs[nb] = carry;
}

//  ---- bulk asynchronous copy + mbarrier (sm_90+; SASS: UBLKCP.S.G, SYNCS.*) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
//  one thread: expect `bytes`, then start the copy global -> shared; completion flips the barrier's phase
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

#define BK2_SMEM_OF(TB) (BK_TABLE_OF(TB) * 8 + BK_CAP * 8 + BK_CAP * 4 + BK_CAP * 2 + BK_TABLE_OF(TB) * 5 * 4 + BK_TABLE_OF(TB) * 4 + 64)

//  out[0]: distinct k-mers (slot records emitted), out[1]: bit 0 = a bucket held more distinct k-mers than the table
//  takes (retry with the larger table), bit 1 = a bucket or the slot scratch overflowed (sorted build)
template <int TB>
__global__ void __launch_bounds__(BK_THREADS, TB == 11 ? 2 : 1)
k_bucket_group2(const uint64_t *__restrict__ btup, const unsigned int *__restrict__ cnt2, const uint32_t *__restrict__ offs, uint32_t nb,
                int K, int B, int posbits, uint64_t mix_inv, uint32_t *__restrict__ occ, IndexSlot *__restrict__ tmp, uint32_t tmp_cap,
                uint32_t *__restrict__ first_pos, uint32_t *first_bitmap, unsigned long long *out) {
  extern __shared__ __align__(128) unsigned char bk_sm[];
  constexpr int BK_TABLE = BK_TABLE_OF(TB), BK_MAXDIST = BK_MAXDIST_OF(TB);
  uint64_t *hk = reinterpret_cast<uint64_t *>(bk_sm);                    // [BK_TABLE] low key bits (mixed k-mer below the bucket bits) of the slot
  uint64_t *tbuf = hk + BK_TABLE;                                        // [BK_CAP] the bucket's tuples, filled by the bulk copy
  uint32_t *gpos = reinterpret_cast<uint32_t *>(tbuf + BK_CAP);          // [BK_CAP] positions grouped by (slot, class)
  uint16_t *ts = reinterpret_cast<uint16_t *>(gpos + BK_CAP);            // [BK_CAP] slot << 3 | class
  uint32_t *hc = reinterpret_cast<uint32_t *>(ts + BK_CAP);              // [BK_TABLE * 5] count -> start -> end of (slot, class)
  uint32_t *hm = hc + BK_TABLE * 5;                                      // [BK_TABLE] first (smallest) position of the slot's k-mer
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int n_dist, bad, slot_base;
  __shared__ unsigned int wsum[BK_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint64_t kmask = (1ull << (2 * K)) - 1;
  const uint64_t posmask = (1ull << posbits) - 1;
  const int lowk = 2 * K - B;                                            // mixed k-mer bits below the bucket number

  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t b = blockIdx.x;
  while (b < nb && cnt2[b] == 0) b += gridDim.x;
  if (tid == 0 && b < nb) bulk_load(tbuf, btup + (uint64_t)b * BK_CAP, ((min(cnt2[b], (unsigned)BK_CAP) * 8u + 15u) & ~15u), &bar);
  uint32_t parity = 0;

  while (b < nb) {
    const uint32_t m = cnt2[b];
    uint32_t nxt = b + gridDim.x;
    while (nxt < nb && cnt2[nxt] == 0) nxt += gridDim.x;
    if (m > BK_CAP) {                                                    // cannot happen after a clean k_part2 (it flags the overflow itself)
      if (tid == 0) atomicOr(&out[1], 2ull);
    }
    const uint32_t s0 = offs[b];
    for (int i = tid; i < BK_TABLE; i += BK_THREADS) { hk[i] = ~0ull; hm[i] = 0xFFFFFFFFu; }
    for (int i = tid; i < BK_TABLE * 5; i += BK_THREADS) hc[i] = 0;
    if (tid == 0) { n_dist = 0; bad = 0; }
    mbar_wait(&bar, parity); parity ^= 1;                                // the bucket's tuples have landed
    uint64_t kreg[BK_PER];
    #pragma unroll
    for (int j = 0; j < BK_PER; j++) {
      const uint32_t i = tid + j * BK_THREADS;
      kreg[j] = i < m ? tbuf[i] : 0ull;
    }
    __syncthreads();                                                     // tbuf consumed, tables cleared
    if (tid == 0 && nxt < nb) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy reads of tbuf before the async-proxy overwrite
      bulk_load(tbuf, btup + (uint64_t)nxt * BK_CAP, ((min(cnt2[nxt], (unsigned)BK_CAP) * 8u + 15u) & ~15u), &bar);
    }
    //  find/insert the k-mer, count per (slot, class)
    #pragma unroll
    for (int j = 0; j < BK_PER; j++) {
      const uint32_t i = tid + j * BK_THREADS;
      if (i >= m) break;
      const uint64_t low = kreg[j] >> posbits;
      const uint32_t pos = (uint32_t)(kreg[j] & posmask);
      const uint32_t cls = (uint32_t)low & 7u;
      const uint64_t km = low >> 3;
      uint32_t h = (uint32_t)((km * 0xD6E8FEB86659FD93ull) >> (64 - TB));
      while (true) {
        const uint64_t cur = hk[h];
        if (cur == km) break;
        if (cur == ~0ull) {
          const unsigned long long old = atomicCAS((unsigned long long *)&hk[h], ~0ull, (unsigned long long)km);
          if (old == ~0ull) { if (atomicAdd(&n_dist, 1u) >= BK_MAXDIST) bad = 1; break; }
          if (old == km) break;
        }
        if (bad) break;
        h = (h + 1) & (BK_TABLE - 1);
      }
      ts[i] = (uint16_t)((h << 3) | cls);
      atomicAdd(&hc[h * 5 + cls], 1u);
      atomicMin(&hm[h], pos);
    }
    __syncthreads();
    if (bad) { if (tid == 0) atomicOr(&out[1], 1ull); b = nxt; continue; }
    //  exclusive prefix sum over hc[slot * 5 + class] in slot order: every thread owns 20 consecutive entries
    {
      const int per = BK_TABLE * 5 / BK_THREADS;
      uint32_t sum = 0;
      for (int j = 0; j < per; j++) sum += hc[tid * per + j];
      uint32_t inc = sum;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        uint32_t w = lane < BK_THREADS / 32 ? wsum[lane] : 0, wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
        if (lane < BK_THREADS / 32) wsum[lane] = wi - w;
      }
      __syncthreads();
      uint32_t run = wsum[wid] + inc - sum;
      for (int j = 0; j < per; j++) { const uint32_t c = hc[tid * per + j]; hc[tid * per + j] = run; run += c; }
    }
    if (tid == 0) {
      const unsigned int nd = n_dist;
      const unsigned long long base = atomicAdd(&out[0], (unsigned long long)nd);
      slot_base = (base + nd <= tmp_cap) ? (unsigned int)base : 0xFFFFFFFFu;
      if (base + nd > tmp_cap) atomicOr(&out[1], 2ull);
    }
    __syncthreads();
    //  scatter the positions to their segment: hc becomes the END of every (slot, class)
    #pragma unroll
    for (int j = 0; j < BK_PER; j++) {
      const uint32_t i = tid + j * BK_THREADS;
      if (i >= m) break;
      const uint32_t sc = ts[i];
      const uint32_t dst = atomicAdd(&hc[(sc >> 3) * 5 + (sc & 7u)], 1u);
      gpos[dst] = (uint32_t)(kreg[j] & posmask);
    }
    __syncthreads();
    for (uint32_t i = tid; i < m; i += BK_THREADS) occ[s0 + i] = gpos[i];
    //  one slot record per occupied table entry: every thread owns BK_TABLE / BK_THREADS consecutive entries
    if (slot_base != 0xFFFFFFFFu) {
      const int per = BK_TABLE / BK_THREADS;
      uint32_t mine = 0;
      #pragma unroll
      for (int j = 0; j < per; j++) mine += hk[tid * per + j] != ~0ull;
      uint32_t inc = mine;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        uint32_t w = lane < BK_THREADS / 32 ? wsum[lane] : 0, wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
        if (lane < BK_THREADS / 32) wsum[lane] = wi - w;
      }
      __syncthreads();
      uint32_t c = slot_base + wsum[wid] + inc - mine;
      #pragma unroll
      for (int j = 0; j < per; j++) {
        const int h = tid * per + j;
        const uint64_t kl = hk[h];
        if (kl == ~0ull) continue;
        const uint32_t st = h ? hc[h * 5 - 1] : 0u;
        const uint64_t km = ((uint64_t)b << lowk) | kl;                  // the mixed k-mer: bucket number on top
        const uint64_t kmer = (km * mix_inv) & kmask;
        uint4 *sp = reinterpret_cast<uint4 *>(&tmp[c]);
        sp[0] = make_uint4((uint32_t)kmer, (uint32_t)(kmer >> 32), s0 + st, s0 + hc[h * 5]);
        sp[1] = make_uint4(s0 + hc[h * 5 + 1], s0 + hc[h * 5 + 2], s0 + hc[h * 5 + 3], s0 + hc[h * 5 + 4]);
        const uint32_t fp = hm[h];
        first_pos[c] = fp;
        atomicOr(&first_bitmap[fp >> 5], 1u << (fp & 31));
        c++;
      }
    }
    __syncthreads();                                                     // everyone is done with the tables before they are cleared again
    b = nxt;
  }
}

//  Path order without a sort: rank of a first position among all first positions = prefix popcount of the bitmap.
//  wprefix[w] = number of set bits in words [0, w).  Three small kernels: per-block sums, scan of the sums, per-word prefix.
#define BM_BLOCK 1024
__global__ void __launch_bounds__(256)
k_bm_blocksum(const uint32_t *__restrict__ bm, uint64_t n_words, uint32_t *__restrict__ bsum) {
  __shared__ uint32_t ws[8];
  const uint64_t base = (uint64_t)blockIdx.x * BM_BLOCK;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < BM_BLOCK; i += 256) { const uint64_t w = base + i; if (w < n_words) s += __popc(bm[w]); }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < 8; i++) t += ws[i]; bsum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256)
k_bm_prefix(const uint32_t *__restrict__ bm, uint64_t n_words, const uint32_t *__restrict__ boff, uint32_t *__restrict__ wprefix) {
  __shared__ uint32_t ws[8];
  const uint64_t base = (uint64_t)blockIdx.x * BM_BLOCK;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  //  thread t owns words base + 4t .. base + 4t + 3
  uint32_t c[4]; uint32_t s = 0;
  #pragma unroll
  for (int j = 0; j < 4; j++) { const uint64_t w = base + 4 * threadIdx.x + j; c[j] = w < n_words ? __popc(bm[w]) : 0; s += c[j]; }
  uint32_t inc = s;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  if (lane == 31) ws[wid] = inc;
  __syncthreads();
  uint32_t wb = 0;
  for (int i = 0; i < wid; i++) wb += ws[i];
  uint32_t run = boff[blockIdx.x] + wb + inc - s;
  #pragma unroll
  for (int j = 0; j < 4; j++) { const uint64_t w = base + 4 * threadIdx.x + j; if (w < n_words) wprefix[w] = run; run += c[j]; }
}

//  one thread per distinct k-mer: its path index from the bitmap, the slot written in place, the hash-table entry
__global__ void __launch_bounds__(256)
k_path_slots2(const IndexSlot *__restrict__ tmp, const uint32_t *__restrict__ first_pos, uint32_t n,
              const uint32_t *__restrict__ bm, const uint32_t *__restrict__ wprefix,
              IndexSlot *__restrict__ slots, HashEntry *ht, uint64_t hcap, int K) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint32_t fp = first_pos[c];
  const uint32_t j = wprefix[fp >> 5] + __popc(bm[fp >> 5] & ((1u << (fp & 31)) - 1u));
  const uint4 *src = reinterpret_cast<const uint4 *>(&tmp[c]);
  const uint4 a = src[0], b = src[1];
  uint4 *dst = reinterpret_cast<uint4 *>(&slots[j]);
  dst[0] = a; dst[1] = b;
  ht_insert(ht, hcap, (uint64_t)a.x | ((uint64_t)a.y << 32), j, K);
}

// ------------------------------------------------------------------------------------------------
//  k-mer census for the skip list: what `meryl count` + `greater-than 1` + `print at-least
//  distinct=D / threshold=T` compute for Canu (src/pipelines/canu/Meryl.pm:529-533,603-607,663-671; threshold from the
//  count histogram: src/meryl/src/meryl/merylOp-nextMer.C:103-115).  Same partition kernels as the index build on the
//  CANONICAL k-mer; one CTA per bucket counts in a shared-memory hash table.  mode 0: histogram of the counts >= 2
//  (ghist[min(count, CENSUS_HMAX - 1)]); mode 1: the k-mers whose count reaches `thr` are appended to (okey, ocnt).
// ------------------------------------------------------------------------------------------------
#define CENSUS_HMAX 65536
#define CENSUS_TB 12
#define CENSUS_SMEM (BK_TABLE_OF(CENSUS_TB) * 8 + BK_CAP * 8 + BK_TABLE_OF(CENSUS_TB) * 4 + 64)
__global__ void __launch_bounds__(BK_THREADS, 2)
k_bucket_census(const uint64_t *__restrict__ btup, const unsigned int *__restrict__ cnt2, uint32_t nb,
                int K, int B, int posbits, uint64_t mix_inv, int mode, uint32_t thr,
                unsigned long long *ghist, uint64_t *okey, uint32_t *ocnt, uint64_t out_cap, unsigned long long *out, Spill S) {
  extern __shared__ __align__(128) unsigned char bk_sm[];
  constexpr int BK_TABLE = BK_TABLE_OF(CENSUS_TB), BK_MAXDIST = BK_MAXDIST_OF(CENSUS_TB);
  uint64_t *hk = reinterpret_cast<uint64_t *>(bk_sm);                    // [BK_TABLE]
  uint64_t *tbuf = hk + BK_TABLE;                                        // [BK_CAP]
  uint32_t *hn = reinterpret_cast<uint32_t *>(tbuf + BK_CAP);            // [BK_TABLE] occurrences
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int n_dist, bad;
  const int tid = threadIdx.x;
  const uint64_t kmask = (1ull << (2 * K)) - 1;
  const int lowk = 2 * K - B;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t b = blockIdx.x;
  while (b < nb && cnt2[b] == 0) b += gridDim.x;
  if (tid == 0 && b < nb) bulk_load(tbuf, btup + (uint64_t)b * BK_CAP, ((min(cnt2[b], (unsigned)BK_CAP) * 8u + 15u) & ~15u), &bar);
  uint32_t parity = 0;
  while (b < nb) {
    const uint32_t m = min(cnt2[b], (unsigned)BK_CAP);
    const bool spilled = cnt2[b] > (unsigned)BK_CAP || S.bucket_spill[b] != 0;
    uint32_t nxt = b + gridDim.x;
    while (nxt < nb && cnt2[nxt] == 0) nxt += gridDim.x;
    for (int i = tid; i < BK_TABLE; i += BK_THREADS) { hk[i] = ~0ull; hn[i] = 0; }
    if (tid == 0) { n_dist = 0; bad = 0; }
    mbar_wait(&bar, parity); parity ^= 1;
    uint64_t kreg[BK_PER];
    #pragma unroll
    for (int j = 0; j < BK_PER; j++) { const uint32_t i = tid + j * BK_THREADS; kreg[j] = i < m ? tbuf[i] : 0ull; }
    __syncthreads();
    if (tid == 0 && nxt < nb) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_load(tbuf, btup + (uint64_t)nxt * BK_CAP, ((min(cnt2[nxt], (unsigned)BK_CAP) * 8u + 15u) & ~15u), &bar);
    }
    #pragma unroll
    for (int j = 0; j < BK_PER; j++) {
      const uint32_t i = tid + j * BK_THREADS;
      if (i >= m) break;
      const uint64_t km = (kreg[j] >> posbits) >> 3;
      if (spilled) { spill_add(S, ((uint64_t)b << lowk) | km); continue; }   // a bucket with an overflowing k-mer is counted in the global table, all of it
      uint32_t h = (uint32_t)((km * 0xD6E8FEB86659FD93ull) >> (64 - CENSUS_TB));
      while (true) {
        const uint64_t cur = hk[h];
        if (cur == km) break;
        if (cur == ~0ull) {
          const unsigned long long old = atomicCAS((unsigned long long *)&hk[h], ~0ull, (unsigned long long)km);
          if (old == ~0ull) { if (atomicAdd(&n_dist, 1u) >= BK_MAXDIST) bad = 1; break; }
          if (old == km) break;
        }
        if (bad) break;
        h = (h + 1) & (BK_TABLE - 1);
      }
      atomicAdd(&hn[h], 1u);
    }
    __syncthreads();
    if (bad) { if (tid == 0) atomicOr(&out[1], 1ull); }
    else if (!spilled) {
      for (int h = tid; h < BK_TABLE; h += BK_THREADS) {
        const uint64_t kl = hk[h];
        if (kl == ~0ull) continue;
        const uint32_t n = hn[h];
        if (mode == 0) {
          if (n >= 2) atomicAdd(&ghist[n < CENSUS_HMAX ? n : CENSUS_HMAX - 1], 1ull);
          else atomicAdd(&ghist[1], 1ull);
        } else if (n >= thr) {
          const unsigned long long o = atomicAdd(&out[0], 1ull);
          if (o < out_cap) { okey[o] = ((((uint64_t)b << lowk) | kl) * mix_inv) & kmask; ocnt[o] = n; }
          else atomicOr(&out[1], 2ull);
        }
      }
    }
    __syncthreads();
    b = nxt;
  }
}

//  the k-mers counted in the spill table: same histogram / emission as the bucket entries
__global__ void __launch_bounds__(256)
k_spill_scan(Spill S, int K, uint64_t mix_inv, int mode, uint32_t thr, unsigned long long *ghist,
             uint64_t *okey, uint32_t *ocnt, uint64_t out_cap, unsigned long long *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > S.mask) return;
  const unsigned long long km = S.key[i];
  if (km == ~0ull) return;
  const uint32_t n = S.cnt[i];
  const uint64_t kmask = (1ull << (2 * K)) - 1;
  if (mode == 0) {
    if (n >= 2) atomicAdd(&ghist[n < CENSUS_HMAX ? n : CENSUS_HMAX - 1], 1ull);
    else atomicAdd(&ghist[1], 1ull);
  } else if (n >= thr) {
    const unsigned long long o = atomicAdd(&out[0], 1ull);
    if (o < out_cap) { okey[o] = (km * mix_inv) & kmask; ocnt[o] = n; }
    else atomicOr(&out[1], 2ull);
  }
}

//  one thread per skip k-mer (the host passes each k-mer once): flag its slot, or append a flagged slot with an empty
//  occurrence list if no hash read holds it (Add_Extra_Hash_String), and mark screened read ends
__global__ void k_index_skip(const uint64_t *__restrict__ skip, uint64_t n_skip, int K, IndexSlot *slots, uint32_t n_distinct,
                             HashEntry *ht, uint64_t hcap, unsigned long long *n_extra,
                             const uint32_t *__restrict__ occ, const uint64_t *__restrict__ hpbase,
                             const uint32_t *__restrict__ hgrp_read, const uint32_t *__restrict__ hlen, uint32_t *hflags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_skip) return;
  const uint64_t key = skip[i];
  const uint32_t idx = ht_find<false>(ht, hcap, slots, key, K);
  if (idx == HT_NOTFOUND) {
    const uint32_t j = n_distinct + (uint32_t)atomicAdd(n_extra, 1ull);
    uint4 *sp = reinterpret_cast<uint4 *>(&slots[j]);
    const uint64_t kk = key | OVL_SKIP_BIT;
    sp[0] = make_uint4((uint32_t)kk, (uint32_t)(kk >> 32), 0, 0);
    sp[1] = make_uint4(0, 0, 0, 0);
    ht_insert(ht, hcap, key, j, K);
    return;
  }
  atomicOr((unsigned long long *)&slots[idx].key, (unsigned long long)OVL_SKIP_BIT);
  const uint32_t st = slots[idx].start, en = slots[idx].end[4];
  for (uint32_t j = st; j < en; j++) {                   // Mark_Screened_Ends_Chain (Build_Hash_Index.C:98-121)
    const uint32_t pos = occ[j];
    const uint32_t r = hgrp_read[pos >> 5];
    const uint32_t q = (uint32_t)(pos - hpbase[r]);
    uint32_t f = 0;
    if (q < OVL_HOPELESS_MATCH) f |= 1u;
    if ((int)hlen[r] - (int)q - K + 1 < OVL_HOPELESS_MATCH) f |= 2u;
    if (f) atomicOr(&hflags[r], f);
  }
}

// ------------------------------------------------------------------------------------------------
//  K2a: probe every ref window (both orientations); emit the occurrence ranges that hold run heads
//
//  Replaces the window loop of Find_Overlaps + Hash_Find (Find_Overlaps.C:177-336).  A hit (p, q) on hash
//  read h continues the diagonal run of hit (p-1, q-1) iff window p-1 is itself a hit window and
//  h[q-1] == ref[p-1]; Add_Match folds such hits into the existing Match_Node (Find_Overlaps.C:45-55).
//  The occurrence list of a k-mer is split by h[q-1], so the hits that START a run are the occurrences in
//  the classes other than ref[p-1] (all of them when window p-1 is not a hit window).  Only those ranges
//  are passed on; on real read sets they are a small fraction of the list.
// ------------------------------------------------------------------------------------------------
struct SlotView { bool found, skip; uint32_t start, e0, e1, e2, e3, e4; };

//  the slot at path index `idx` if it holds `key`
__device__ __forceinline__ bool slot_at(const IndexSlot *__restrict__ slots, uint32_t idx, uint64_t key, SlotView &v) {
  //  one 256-bit load = the whole slot = one 32-byte sector (LDG.E.256 on sm_100)
  uint64_t k, q1, q2, q3;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(slots + idx));
  if ((k & ~OVL_SKIP_BIT) != key) return false;
  v.found = true; v.skip = (k & OVL_SKIP_BIT) != 0;
  v.start = (uint32_t)q1; v.e0 = (uint32_t)(q1 >> 32); v.e1 = (uint32_t)q2; v.e2 = (uint32_t)(q2 >> 32);
  v.e3 = (uint32_t)q3; v.e4 = (uint32_t)(q3 >> 32);
  return true;
}

//  Full lookup through the hash table; returns the path index or HT_NOTFOUND.  FR (the lookup of a FORWARD window)
//  also learns whether the reverse complement of the k-mer -- the k-mer of the reverse-strand window over the same
//  bases -- can be in the table: `rc_absent` is set only if it certainly is not (no entry of the probe sequence
//  carries the other strand's fingerprint).  The four entries of a bucket are compared first and the candidates (one,
//  as good as always) confirmed by their slot in a loop, so that the 32-byte slot load is in the code once.
template <bool FR>
__device__ __forceinline__ uint32_t slot_lookup(const IndexSlot *__restrict__ slots, const HashEntry *__restrict__ ht, uint64_t hcap,
                                                uint64_t key, int K, SlotView &v, bool &rc_absent) {
  //  keep the hash arithmetic (~45 instructions) inside the branch that looks up: it is loop-invariant, and the compiler
  //  otherwise hoists it in front of the probe kernel's round loop, where every lane of every group pays for it
  //  although most groups of a well-covered block never touch the table (ncu: 11 % of the kernel's instructions)
  asm volatile("" : "+l"(key));
  const uint64_t rkey = kmer_rc(key, K);
  const uint64_t h = ht_hash_of(key, rkey), nb = hcap >> 2;
  const uint32_t fp = ht_fp2(h, key, rkey);
  uint32_t found = HT_NOTFOUND;
  bool r_seen = rkey == key;                                             // a palindrome is its own reverse complement
  for (uint64_t b = ht_bucket_of(h, hcap);;) {
    uint64_t e[4];
    ht_load_bucket<true>(ht, b, e);
    bool open = false;
    unsigned cand = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t f = (uint32_t)(e[i] >> 32);
      open |= e[i] == HT_EMPTY;
      if (f == fp && e[i] != HT_EMPTY) cand |= 1u << i;
      if (FR && f == (fp ^ 1u)) r_seen = true;                            // may be the other strand's entry: no need to know for sure
    }
    while (cand && found == HT_NOTFOUND) {
      const int i = __ffs(cand) - 1; cand &= cand - 1;
      const uint32_t idx = (uint32_t)(i == 0 ? e[0] : i == 1 ? e[1] : i == 2 ? e[2] : e[3]);
      if (slot_at(slots, idx, key, v)) found = idx;
    }
    if (open || (found != HT_NOTFOUND && (!FR || r_seen))) break;
    if (++b == nb) b = 0;
  }
  rc_absent = FR && !r_seen;
  return found;
}

#define SMALL_ITEM_MAX 8u
#define STAGE_SMALL 96          // per-warp shared-memory staging of head ranges (uint4 entries)
#define STAGE_LARGE 64
#define PROBE_CHUNK 32          // consecutive 32-window groups handled by one warp in a row

//  Per-warp staging of items in shared memory: one global atomicAdd per ~64 items instead of one per group
//  (13 M same-address atomics per C2 tile were the probe kernel's bottleneck).
struct ItemStage { uint4 *buf; int n; int cap; uint4 *out; unsigned long long *counter; uint64_t out_cap; };

__device__ __forceinline__ void stage_flush(ItemStage &S, int lane) {
  if (S.n == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(S.counter, (unsigned long long)S.n);
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  for (int j = lane; j < S.n; j += 32)
    if (base + j < S.out_cap) S.out[base + j] = S.buf[j];
  __syncwarp();
  S.n = 0;
}

//  item = (ref position index, first occurrence, count << 1 | dir)
__device__ __forceinline__ void stage_append(ItemStage &S, bool has, uint4 item, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, has);
  if (!m) return;
  if (S.n + 32 > S.cap) stage_flush(S, lane);
  if (has) S.buf[S.n + __popc(m & ((1u << lane) - 1))] = item;
  S.n += __popc(m);
}

//  Persistent kernel: each warp walks PROBE_CHUNK consecutive groups at a time.
//
//  Lookup of the 32 windows of a group: lane l first tries path slot base + l, where `base` continues the slot of
//  the previous group's last window (or comes from a hash-table lookup by the first lane still unresolved); every
//  lane whose k-mer is there is done.  After PROBE_ROUNDS such rounds the lanes still unresolved (windows with read
//  errors, path breaks) look themselves up through the hash table, in parallel.
#define PROBE_ROUNDS 3
template <int DIR>
__global__ void __launch_bounds__(THREADS, 4)
k_ref_probe(const uint64_t *__restrict__ fwd, const uint64_t *__restrict__ rc, const uint64_t *__restrict__ woff,
            const uint32_t *__restrict__ len, const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read,
            uint64_t n_groups, uint64_t n_pos, int K, const IndexSlot *__restrict__ slots, uint32_t n_slots,
            const HashEntry *__restrict__ ht, uint64_t hcap,
            uint32_t *__restrict__ ref_valid, uint32_t *rflags,
            uint4 *__restrict__ item_small, uint4 *__restrict__ item_large, uint64_t item_cap, unsigned long long *work, int flags,
            uint32_t *rc_absent, uint64_t gg_lo, uint64_t gg_hi) {
  const bool one_list = flags & 1, adaptive = flags & 2;
  __shared__ uint4 stage[WARPS_PER_BLOCK][STAGE_SMALL + STAGE_LARGE];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  ItemStage SS, SL;
  SS.buf = stage[wib];               SS.n = 0; SS.cap = one_list ? STAGE_SMALL + STAGE_LARGE : STAGE_SMALL; SS.out = item_small; SS.counter = &work[3]; SS.out_cap = item_cap;
  SL.buf = stage[wib] + STAGE_SMALL; SL.n = 0; SL.cap = STAGE_LARGE; SL.out = item_large; SL.counter = &work[4]; SL.out_cap = item_cap;

  //  One launch per orientation, [gg_lo, gg_hi) = [0, n_groups) then [n_groups, 2 n_groups): the forward pass leaves
  //  in `rc_absent` (one bit per position of the ref batch, reverse-strand coordinates) the windows whose k-mer it has
  //  SEEN to be missing from the table while looking up the forward window over the same bases; the reverse pass does
  //  not look those up.
  const uint64_t n_chunks = (gg_hi - gg_lo + PROBE_CHUNK - 1) / PROBE_CHUNK;
  KmerWindow KW; KW.w = nullptr; KW.w0 = 0; KW.codes = 0; KW.inv = 0xFFFFu;
  for (uint64_t ch = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + wib; ch < n_chunks; ch += (uint64_t)gridDim.x * WARPS_PER_BLOCK) {
    const uint64_t gg_end = min(gg_hi, gg_lo + (ch + 1) * PROBE_CHUNK);
    uint32_t prev_r = 0xFFFFFFFFu; int prev_dir = -1; unsigned prev_top = 0;
    uint32_t next_idx = HT_NOTFOUND;                          // path slot expected for window p0 if the path continues
    bool miss_mode = false;
    for (uint64_t gg = gg_lo + ch * PROBE_CHUNK; gg < gg_end; gg++) {
      constexpr int dir = DIR;
      const uint64_t g = dir ? gg - n_groups : gg;
      const uint32_t absent_w = dir ? __ldcg(rc_absent + g) : 0u;   // issued first: nothing below depends on it until the lookups
      const uint32_t r = grp_read[g];
      const int L = (int)len[r];
      const int p0 = (int)(g * 32 - pbase[r]);
      const int p = p0 + lane;
      const uint64_t *w = (dir ? rc : fwd) + woff[r];
      const bool carried = (r == prev_r && dir == prev_dir);    // the previous iteration was window group p0-32 of this read

      uint64_t key; int cls;
      bool ok = warp_kmers_win(KW, w, p0, L, K, lane, key, cls);
      if (dir) ok = ok && !((absent_w >> lane) & 1u);
      bool r_absent = false;                                  // forward pass: my k-mer's reverse complement is not in the table
      SlotView v; v.found = false; v.skip = false; v.start = v.e0 = v.e1 = v.e2 = v.e3 = v.e4 = 0;
      uint32_t my_idx = HT_NOTFOUND;
      {
        unsigned need = __ballot_sync(0xffffffffu, ok);
        const unsigned n_ok = __popc(need);
        uint32_t base = carried ? next_idx : HT_NOTFOUND;
        //  miss_mode: most windows of the previous group were not in the table (a ref read from a part of the genome
        //  the hash block does not cover: 85 % of the windows of a human-size job).  There is no path to follow there,
        //  and one lane looking itself up per round is a chain of dependent DRAM latencies: all lanes go straight to
        //  the hash table, in parallel.
        for (int round = 0; need && round < PROBE_ROUNDS && !(miss_mode && base == HT_NOTFOUND); round++) {
          const int j = __ffs(need) - 1;
          if (base == HT_NOTFOUND) {                            // lane j finds its own slot; the others follow it
            if (lane == j) my_idx = slot_lookup<DIR == 0>(slots, ht, hcap, key, K, v, r_absent);
            const uint32_t ij = __shfl_sync(0xffffffffu, my_idx, j);
            need &= ~(1u << j);
            if (ij == HT_NOTFOUND && adaptive) break;           // no path here either: the rest in parallel
            if (ij == HT_NOTFOUND || !need) continue;
            base = ij - (uint32_t)j;
          }
          const uint32_t cand = base + (uint32_t)lane;
          bool hit = false;
          if (((need >> lane) & 1u) && cand < n_slots) hit = slot_at(slots, cand, key, v);
          if (hit) my_idx = cand;
          need &= ~__ballot_sync(0xffffffffu, hit);
          base = HT_NOTFOUND;
        }
        if ((need >> lane) & 1u) my_idx = slot_lookup<DIR == 0>(slots, ht, hcap, key, K, v, r_absent);
        const uint32_t last = __shfl_sync(0xffffffffu, my_idx, 31);
        next_idx = (last == HT_NOTFOUND) ? HT_NOTFOUND : last + 1;
        miss_mode = adaptive && 2 * __popc(__ballot_sync(0xffffffffu, my_idx != HT_NOTFOUND)) < n_ok;
      }
      if (!dir) {
        //  forward window p covers the bases of reverse-strand window L - K - p: lane i <-> position q - i, q = L - K - p0
        const unsigned am = __ballot_sync(0xffffffffu, r_absent);
        if (am && lane == 0) {
          const int q = L - K - p0;
          unsigned rev = __brev(am);                          // bit k <-> position q - 31 + k
          uint64_t base = pbase[r];
          if (q >= 31) base += (uint64_t)(q - 31); else rev >>= (31 - q);    // positions below 0 belong to lanes past the last window
          const uint64_t wd = base >> 5; const int sft = (int)(base & 31);
          const unsigned lo = rev << sft, hi = sft ? rev >> (32 - sft) : 0u;
          if (lo) atomicOr(&rc_absent[wd], lo);
          if (hi) atomicOr(&rc_absent[wd + 1], hi);
        }
      }
      if (v.found && v.skip) {                               // hi_hits (Find_Overlaps.C:274-276,310-316)
        uint32_t f = 0;
        if (p == 0) f = 1u;
        else {
          if (p < OVL_HOPELESS_MATCH) f |= 1u;
          if (L - p - K + 1 < OVL_HOPELESS_MATCH) f |= 2u;
        }
        if (f) atomicOr(&rflags[2 * r + dir], f);
      }
      const bool valid = v.found && !v.skip && v.e4 > v.start;
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) ref_valid[((uint64_t)dir * n_pos >> 5) + g] = vm;
      const unsigned carry_top = prev_top;
      prev_r = r; prev_dir = dir; prev_top = vm >> 31;
      if (vm == 0) continue;

      //  is window p-1 a hit window?  lanes 1..31 see it in the ballot; lane 0 takes it from the previous group of
      //  this chunk, or (first group of a chunk) looks window p0-1 up itself
      bool prev_valid = (lane > 0) ? ((vm >> (lane - 1)) & 1u) : (carried && carry_top);
      if (lane == 0 && !carried && valid && p0 > 0 && cls != 0) {
        //  k-mer of window p0-1 = my k-mer shifted up one base with ref[p0-1] in front (it cannot contain an N:
        //  its last K-1 bases are mine and cls != 0 says base p0-1 is ACGT)
        const uint64_t pk = ((key << 2) | (uint64_t)(cls - 1)) & ((1ull << (2 * K)) - 1);
        SlotView pv; pv.found = false; pv.skip = false; pv.start = pv.e4 = 0;
        bool dummy; slot_lookup<false>(slots, ht, hcap, pk, K, pv, dummy);
        prev_valid = pv.found && !pv.skip && pv.e4 > pv.start;
      }

      //  head ranges: the whole list, or the list minus the class of ref[p-1]
      uint32_t b0 = v.start, n0 = 0, b1 = 0, n1 = 0;
      if (valid) {
        if (prev_valid && cls != 0) {
          const uint32_t lo = (cls == 1) ? v.e0 : (cls == 2) ? v.e1 : (cls == 3) ? v.e2 : v.e3;     // start of class cls
          const uint32_t hi = (cls == 1) ? v.e1 : (cls == 2) ? v.e2 : (cls == 3) ? v.e3 : v.e4;     // end of class cls
          n0 = lo - v.start; b1 = hi; n1 = v.e4 - hi;
        } else {
          n0 = v.e4 - v.start;
        }
      }
      if (__any_sync(0xffffffffu, n0 | n1)) {
        const uint32_t pos = (uint32_t)(g * 32 + lane);
        if (one_list) {                                  // the cooperative expand kernel takes items of any size from one list
          stage_append(SS, n0 > 0, make_uint4(pos, b0, (n0 << 1) | (uint32_t)dir, 0), lane);
          stage_append(SS, n1 > 0, make_uint4(pos, b1, (n1 << 1) | (uint32_t)dir, 0), lane);
        } else {
          stage_append(SS, n0 > 0 && n0 <= SMALL_ITEM_MAX, make_uint4(pos, b0, (n0 << 1) | (uint32_t)dir, 0), lane);
          stage_append(SL, n0 > SMALL_ITEM_MAX,            make_uint4(pos, b0, (n0 << 1) | (uint32_t)dir, 0), lane);
          stage_append(SS, n1 > 0 && n1 <= SMALL_ITEM_MAX, make_uint4(pos, b1, (n1 << 1) | (uint32_t)dir, 0), lane);
          stage_append(SL, n1 > SMALL_ITEM_MAX,            make_uint4(pos, b1, (n1 << 1) | (uint32_t)dir, 0), lane);
        }
      }
    }
  }
  stage_flush(SS, lane);
  stage_flush(SL, lane);
}

// ------------------------------------------------------------------------------------------------
//  K2b: turn head occurrences into seed runs (Add_Ref + the node-creating branch of Add_Match,
//  Find_Overlaps.C:61-86,105-163): ID filter, run length, warp-aggregated append of (key, value)
// ------------------------------------------------------------------------------------------------
struct ExpandArgs {
  const uint64_t *rfwd, *rrc, *rwoff; const uint32_t *rlen; const uint64_t *rpbase; const uint32_t *rgrp_read;
  uint64_t r_npos; uint32_t ref_first_id;
  const uint64_t *hfwd, *hwoff; const uint32_t *hlen; const uint64_t *hpbase; const uint32_t *hgrp_read; uint32_t hash_first_id;
  int K;
  const uint32_t *occ; const uint32_t *ref_valid;
  uint64_t *run_key, *run_val; uint64_t run_cap;
  unsigned long long *n_runs;
};

//  one head occurrence: returns true and fills (key, val, runlen) if it yields a run
__device__ __forceinline__ bool head_to_run(const ExpandArgs &A, uint32_t pos, int dir, uint32_t hpos, uint64_t &rkey, uint64_t &rval, uint32_t &runlen) {
  const uint32_t r = A.rgrp_read[pos >> 5];
  const uint32_t hh = A.hgrp_read[hpos >> 5];
  if (A.ref_first_id + r >= A.hash_first_id + hh) return false;          // only refID < hashID pairs (Find_Overlaps.C:279,320)
  const int p = (int)(pos - A.rpbase[r]);
  const int q = (int)(hpos - A.hpbase[hh]);
  const int L = (int)A.rlen[r], HL = (int)A.hlen[hh];
  const uint64_t *rw = (dir ? A.rrc : A.rfwd) + A.rwoff[r];
  const uint64_t *hw = A.hfwd + A.hwoff[hh];
  //  run length: equal bases after the k-mer ...
  const int lim = min(L - (p + A.K), HL - (q + A.K));
  int e2 = 0;
  while (e2 < lim) {
    const int k = ovl_equal16(ovl_fetch16(rw, p + A.K + e2), ovl_fetch16(hw, q + A.K + e2));
    e2 += k;
    if (k < 16) break;
  }
  if (e2 > lim) e2 = lim;
  const int want = e2 + 1;
  //  ... and consecutive hit windows on the ref side (a skip k-mer or an N ends the run)
  const uint64_t vbase = (uint64_t)dir * A.r_npos;
  uint64_t wi = (vbase + pos) >> 5;
  const int b = (int)(pos & 31);
  const uint32_t rem = A.ref_valid[wi] >> b;
  int nv = __ffs(~rem) - 1;
  if (nv < 0 || nv > 32 - b) nv = 32 - b;
  if (nv == 32 - b) {
    wi++;
    while (nv < want) {
      const uint32_t x = A.ref_valid[wi++];
      if (x == 0xFFFFFFFFu) { nv += 32; continue; }
      nv += __ffs(~x) - 1;
      break;
    }
  }
  runlen = (uint32_t)min(want, nv);
  rkey = ovl_runkey(r, (uint32_t)dir, hh, (uint32_t)p);
  rval = (uint64_t)(uint32_t)q | ((uint64_t)runlen << 32);
  return true;
}

//  Runs are staged per warp in shared memory and appended RUN_STAGE at a time: one global atomic per ~100 runs
//  instead of one per loop iteration (a few million same-address atomics per C2 tile otherwise).
#define RUN_STAGE 128
struct RunStage { uint64_t *key, *val; int n; };

__device__ __forceinline__ void run_flush(const ExpandArgs &A, RunStage &S, int lane) {
  if (S.n == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(A.n_runs, (unsigned long long)S.n);
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  for (int j = lane; j < S.n; j += 32)
    if (base + j < A.run_cap) { A.run_key[base + j] = S.key[j]; A.run_val[base + j] = S.val[j]; }
  __syncwarp();
  S.n = 0;
}

__device__ __forceinline__ void append_run(const ExpandArgs &A, RunStage &S, bool has, uint64_t rkey, uint64_t rval, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, has);
  if (!m) return;
  if (S.n + 32 > RUN_STAGE) run_flush(A, S, lane);
  if (has) { const int i = S.n + __popc(m & ((1u << lane) - 1)); S.key[i] = rkey; S.val[i] = rval; }
  S.n += __popc(m);
}

//  thread per item, items of at most SMALL_ITEM_MAX occurrences
__global__ void __launch_bounds__(256)
k_expand_small(ExpandArgs A, const uint4 *__restrict__ items, const unsigned long long *n_items_p, uint64_t item_cap,
               unsigned long long *counters) {
  __shared__ uint64_t st_key[8][RUN_STAGE], st_val[8][RUN_STAGE];
  const int lane = threadIdx.x & 31;
  RunStage S; S.key = st_key[threadIdx.x >> 5]; S.val = st_val[threadIdx.x >> 5]; S.n = 0;
  unsigned long long n_items = *n_items_p; if (n_items > item_cap) n_items = item_cap;
  unsigned long long hits = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n_round = (n_items + 31) & ~31ull;      // whole warps stay together for the ballots
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    uint4 it = make_uint4(0, 0, 0, 0);
    if (i < n_items) it = items[i];
    const uint32_t cnt = it.z >> 1; const int dir = (int)(it.z & 1u);
    const uint32_t mx = __reduce_max_sync(0xffffffffu, cnt);
    for (uint32_t j = 0; j < mx; j++) {
      uint64_t rk = 0, rv = 0; uint32_t rl = 0;
      bool has = false;
      if (j < cnt) has = head_to_run(A, it.x, dir, A.occ[it.y + j], rk, rv, rl);
      if (has) hits += rl;
      append_run(A, S, has, rk, rv, lane);
    }
  }
  run_flush(A, S, lane);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync(0xffffffffu, hits, o);
  if (lane == 0 && hits) atomicAdd(&counters[CT_SEED_HITS], hits);
}

//  warp per item
__global__ void __launch_bounds__(256)
k_expand_large(ExpandArgs A, const uint4 *__restrict__ items, const unsigned long long *n_items_p, uint64_t item_cap,
               unsigned long long *counters) {
  __shared__ uint64_t st_key[8][RUN_STAGE], st_val[8][RUN_STAGE];
  const int lane = threadIdx.x & 31;
  RunStage S; S.key = st_key[threadIdx.x >> 5]; S.val = st_val[threadIdx.x >> 5]; S.n = 0;
  unsigned long long n_items = *n_items_p; if (n_items > item_cap) n_items = item_cap;
  unsigned long long hits = 0;
  const uint64_t wstride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_items; i += wstride) {
    const uint4 it = items[i];
    const uint32_t cnt = it.z >> 1; const int dir = (int)(it.z & 1u);
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
      const uint32_t j = j0 + lane;
      uint64_t rk = 0, rv = 0; uint32_t rl = 0;
      bool has = false;
      if (j < cnt) has = head_to_run(A, it.x, dir, A.occ[it.y + j], rk, rv, rl);
      if (has) hits += rl;
      append_run(A, S, has, rk, rv, lane);
    }
  }
  run_flush(A, S, lane);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync(0xffffffffu, hits, o);
  if (lane == 0 && hits) atomicAdd(&counters[CT_SEED_HITS], hits);
}

//  Round 2: one cooperative kernel for all items.  The thread-per-item / warp-per-item kernels above ran with 6-7 of 32
//  lanes active (ncu: thread_inst_executed_per_inst 6.2 / 7.5): on HiFi-like reads a run is hundreds of hits long and
//  every lane walked its own 16-bases-per-step comparison loop for a different number of steps.  Here a warp takes 32
//  items, expands them into (item, occurrence) candidates by a prefix sum (load-balanced: 32 candidates at a time
//  whatever the item sizes), filters them in parallel (refID < hashID, all metadata loads independent), and then
//  measures each surviving run with the WHOLE warp: 512 bases per step, like the extension kernel's slide.
struct ExpCand { const uint64_t *rw, *hw; int p, q, lim; uint32_t pos_dir; uint32_t r, hh; };   // 40 bytes

__global__ void __launch_bounds__(256)
k_expand_coop(ExpandArgs A, const uint4 *__restrict__ items_small, const unsigned long long *n_small_p,
              const uint4 *__restrict__ items_large, const unsigned long long *n_large_p, uint64_t item_cap,
              unsigned long long *counters) {
  __shared__ uint64_t st_key[8][RUN_STAGE], st_val[8][RUN_STAGE];
  __shared__ ExpCand cand[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  RunStage S; S.key = st_key[wib]; S.val = st_val[wib]; S.n = 0;
  ExpCand *C = cand[wib];
  unsigned long long n_small = *n_small_p, n_large = *n_large_p;
  if (n_small > item_cap) n_small = item_cap;
  if (n_large > item_cap) n_large = item_cap;
  const unsigned long long n_items = n_small + n_large;
  unsigned long long hits = 0;
  const uint64_t wstride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; i0 < n_items; i0 += wstride * 32) {
    const uint64_t i = i0 + lane;
    uint4 it = make_uint4(0, 0, 0, 0);
    if (i < n_items) it = i < n_small ? items_small[i] : items_large[i - n_small];
    const uint32_t cnt = it.z >> 1;
    //  per item (ref side): read, offset in the read, its words
    uint32_t r = 0; uint64_t rp = 0, rwo = 0; int rL = 0;
    if (cnt) { r = A.rgrp_read[it.x >> 5]; rp = A.rpbase[r]; rL = (int)A.rlen[r]; rwo = A.rwoff[r]; }
    uint32_t incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
    const uint32_t excl = incl - cnt;
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    for (uint32_t cb = 0; cb < total; cb += 32) {
      const uint32_t c = cb + lane;
      const bool active = c < total;
      //  owner item of candidate c: the largest l with excl_l <= c
      int l = 0;
      #pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int probe = l + step;
        const uint32_t v = __shfl_sync(0xffffffffu, excl, probe & 31);
        if (probe < 32 && v <= c) l = probe;
      }
      const uint32_t j = c - __shfl_sync(0xffffffffu, excl, l);
      const uint32_t pos_l = __shfl_sync(0xffffffffu, it.x, l), y_l = __shfl_sync(0xffffffffu, it.y, l), z_l = __shfl_sync(0xffffffffu, it.z, l);
      const uint32_t r_l = __shfl_sync(0xffffffffu, r, l);
      const uint64_t rp_l = __shfl_sync(0xffffffffu, rp, l), rwo_l = __shfl_sync(0xffffffffu, rwo, l);
      const int rL_l = __shfl_sync(0xffffffffu, rL, l);
      bool pass = false;
      if (active) {
        const uint32_t hpos = A.occ[y_l + j];
        const uint32_t hh = A.hgrp_read[hpos >> 5];
        if (A.ref_first_id + r_l < A.hash_first_id + hh) {                // only refID < hashID pairs (Find_Overlaps.C:279,320)
          pass = true;
          const int dir = (int)(z_l & 1u);
          ExpCand e;
          e.p = (int)(pos_l - rp_l);
          e.q = (int)(hpos - A.hpbase[hh]);
          e.lim = min(rL_l - (e.p + A.K), (int)A.hlen[hh] - (e.q + A.K));
          e.rw = (dir ? A.rrc : A.rfwd) + rwo_l;
          e.hw = A.hfwd + A.hwoff[hh];
          e.pos_dir = pos_l; e.r = (r_l << 1) | (uint32_t)dir; e.hh = hh;
          C[lane] = e;
          //  the survivors are measured one after the other by the whole warp: have their first lines on the way
          asm volatile("prefetch.global.L1 [%0];" :: "l"(e.rw + ((e.p + A.K) >> 4)));
          asm volatile("prefetch.global.L1 [%0];" :: "l"(e.hw + ((e.q + A.K) >> 4)));
          if (e.lim > 240) {                               // a 512-base step reads two 128-byte lines of either read
            asm volatile("prefetch.global.L1 [%0];" :: "l"(e.rw + ((e.p + A.K) >> 4) + 16));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(e.hw + ((e.q + A.K) >> 4) + 16));
          }
          asm volatile("prefetch.global.L1 [%0];" :: "l"(A.ref_valid + (((uint64_t)dir * A.r_npos + pos_l) >> 5)));
        }
      }
      __syncwarp();
      for (unsigned surv = __ballot_sync(0xffffffffu, pass); surv; surv &= surv - 1) {
        const ExpCand e = C[__ffs(surv) - 1];
        const int dir = (int)(e.r & 1u);
        //  run length: equal bases after the k-mer, 512 per step ...
        int e2 = 0;
        while (e2 < e.lim) {
          const int off = e2 + 16 * lane;
          int k = 16;
          if (off < e.lim) k = ovl_equal16(ovl_fetch16(e.rw, e.p + A.K + off), ovl_fetch16(e.hw, e.q + A.K + off));
          const unsigned nf = __ballot_sync(0xffffffffu, k < 16);
          if (nf == 0) { e2 += 512; continue; }
          const int first = __ffs(nf) - 1;
          e2 += 16 * first + __shfl_sync(0xffffffffu, k, first);
          break;
        }
        if (e2 > e.lim) e2 = e.lim;
        const int want = e2 + 1;
        //  ... and consecutive hit windows on the ref side (a skip k-mer or an N ends the run)
        const uint64_t vbase = (uint64_t)dir * A.r_npos;
        uint64_t wi = (vbase + e.pos_dir) >> 5;
        const int b = (int)(e.pos_dir & 31);
        const uint32_t rem = A.ref_valid[wi] >> b;
        int nv = __ffs(~rem) - 1;
        if (nv < 0 || nv > 32 - b) nv = 32 - b;
        if (nv == 32 - b) {
          wi++;
          while (nv < want) {
            const uint32_t x = A.ref_valid[wi + lane];
            const unsigned part = __ballot_sync(0xffffffffu, x != 0xFFFFFFFFu);
            if (part == 0) { nv += 1024; wi += 32; continue; }
            const int f = __ffs(part) - 1;
            const uint32_t xf = __shfl_sync(0xffffffffu, x, f);
            nv += 32 * f + (__ffs(~xf) - 1);
            break;
          }
        }
        const uint32_t runlen = (uint32_t)min(want, nv);
        if (lane == 0) {
          S.key[S.n] = ovl_runkey(e.r >> 1, (uint32_t)dir, e.hh, (uint32_t)e.p);
          S.val[S.n] = (uint64_t)(uint32_t)e.q | ((uint64_t)runlen << 32);
          hits += runlen;
        }
        S.n++;
        if (S.n == RUN_STAGE) { __syncwarp(); run_flush(A, S, lane); }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  run_flush(A, S, lane);
  if (lane == 0 && hits) atomicAdd(&counters[CT_SEED_HITS], hits);
}

// ------------------------------------------------------------------------------------------------
//  K3: pairs and seed lists
// ------------------------------------------------------------------------------------------------
__global__ void k_pair_heads(const uint64_t *__restrict__ key, uint64_t n, uint32_t *__restrict__ flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || ovl_runkey_pair(key[i]) != ovl_runkey_pair(key[i - 1])) ? 1u : 0u;
}

//  pair_idx = exclusive scan of flag.  Heads write the pair's first run; run records are unpacked.
__global__ void k_pair_scatter(const uint64_t *__restrict__ key, const uint64_t *__restrict__ val, uint64_t n,
                               const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pair_idx,
                               PairRec *__restrict__ pairs, OvlRun *__restrict__ runs) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = key[i], v = val[i];
  OvlRun rr;
  rr.start = (int32_t)(k & ((1u << OVL_RUNKEY_POS_BITS) - 1));
  rr.q = (int32_t)(uint32_t)v;
  rr.len = (int32_t)(v >> 32);
  runs[i] = rr;
  if (flag[i]) {
    PairRec pr;
    pr.ref_idx  = (uint32_t)(k >> 46);
    pr.dir      = (int32_t)((k >> 45) & 1);
    pr.hash_idx = (uint32_t)((k >> 21) & ((1u << OVL_RUNKEY_HASH_BITS) - 1));
    pr.consistent = 1; pr.diag_ct = 0; pr.diag_bgn = 0; pr.diag_end = 0; pr.n_seeds = 0;
    pr.seed_begin = (int64_t)i;
    pairs[pair_idx[i]] = pr;
  }
}

//  one thread per pair: replay Add_Match over the runs, apply --minkmers and the hopeless check
__global__ void __launch_bounds__(128)
k_chain_pairs(PairRec *pairs, uint64_t n_pairs, uint64_t n_runs, const OvlRun *__restrict__ runs,
              int32_t *nxt, int32_t *hits, int32_t *act, int32_t *order,
              int32_t *__restrict__ seed_start, int32_t *__restrict__ seed_off, int32_t *__restrict__ seed_len,
              uint8_t *__restrict__ seed_alive,
              DevParams P, const uint32_t *__restrict__ rlen, const uint32_t *__restrict__ rflags,
              const uint32_t *__restrict__ hlen, const uint32_t *__restrict__ hflags, unsigned long long *counters) {
  uint64_t pi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= n_pairs) return;
  PairRec pr = pairs[pi];
  const int64_t b = pr.seed_begin;
  const int64_t e = (pi + 1 < n_pairs) ? pairs[pi + 1].seed_begin : (int64_t)n_runs;
  const int n = (int)(e - b);
  const OvlRun *rr = runs + b;

  int consistent = 1;
  if (n == 1) {
    order[b] = 0;
  } else {
    consistent = ovl_chain_simulate(rr, n, P.K, nxt + b, hits + b, act + b, order + b);
  }

  int diag_ct = 0, diag_end = 0;
  for (int i = 0; i < n; i++) {
    diag_ct += rr[i].len;
    int last = rr[i].start + rr[i].len - 1;
    if (last > diag_end) diag_end = last;
  }
  const int diag_bgn = rr[0].start;

  for (int k = 0; k < n; k++) {
    const OvlRun x = rr[order[b + k]];
    seed_start[b + k] = x.start;
    seed_off[b + k]   = x.q;
    seed_len[b + k]   = P.K + x.len - 1;
    seed_alive[b + k] = 1;
  }

  pr.consistent = consistent;
  pr.diag_ct = diag_ct; pr.diag_bgn = diag_bgn; pr.diag_end = diag_end;
  pr.n_seeds = n;

  //  --minkmers (Process_String_Overlaps.C:22-33,618-621)
  if (P.filter_by_kmer_count != 0) {
    double ovl_len = (double)(diag_end - diag_bgn);
    unsigned long long expected = 0;
    if (!(ovl_len < (double)P.K))
      expected = (unsigned long long)(int)floor(__dmul_rn(P.minkmers_factor, ovl_len - (double)P.K + 1.0));
    unsigned long long need = P.filter_by_kmer_count > expected ? P.filter_by_kmer_count : expected;
    if (need > (unsigned long long)diag_ct) {
      atomicAdd(&counters[CT_HITS_SKIPPED], 1ull);
      pr.n_seeds = 0;
      pairs[pi] = pr;
      return;
    }
  }

  atomicAdd(&counters[CT_PAIRS], 1ull);

  //  hopeless check on singleton seeds (Process_String_Overlaps.C:384-415)
  if (P.use_hopeless && n == 1 && !P.partial) {
    const int s_len = (int)rlen[pr.ref_idx], t_len = (int)hlen[pr.hash_idx];
    const uint32_t rf = rflags[2 * pr.ref_idx + pr.dir], hf = hflags[pr.hash_idx];
    const int s_head = seed_start[b], t_head = seed_off[b], ln = seed_len[b];
    bool hopeless = false;
    if (s_head <= t_head) { if (s_head > OVL_HOPELESS_MATCH && !(rf & 1u)) hopeless = true; }
    else                  { if (t_head > OVL_HOPELESS_MATCH && !(hf & 1u)) hopeless = true; }
    const int s_tail = s_len - s_head - ln + 1, t_tail = t_len - t_head - ln + 1;
    if (s_tail <= t_tail) { if (s_tail > OVL_HOPELESS_MATCH && !(rf & 2u)) hopeless = true; }
    else                  { if (t_tail > OVL_HOPELESS_MATCH && !(hf & 2u)) hopeless = true; }
    if (hopeless) {
      atomicAdd(&counters[CT_HITS_WITHOUT], 1ull);
      pr.n_seeds = 0;
    }
  }
  pairs[pi] = pr;
}

//  Scheduling key of a pair for the extension kernel: pairs with many seeds (non-consistent pairs on noisy reads: one
//  Extend_Alignment per surviving seed) go first, so that the persistent kernel does not end on a few warps grinding
//  through the heaviest pairs (longest-processing-time-first).  Ascending sort on ~cost.
__global__ void k_pair_cost(const PairRec *__restrict__ pairs, uint32_t n_pairs, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const PairRec pr = pairs[i];
  const uint32_t cost = pr.n_seeds <= 0 ? 0u : (pr.consistent ? 1u : (uint32_t)pr.n_seeds);
  key[i] = ~cost; val[i] = i;
}

// ------------------------------------------------------------------------------------------------
//  host-side launch helpers
// ------------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ovl_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return OVLB_ERR_CUDA; } } while (0)

template <typename T>
static int ensure(T *&ptr, size_t &cap, size_t need, size_t slack_num = 5, size_t slack_den = 4) {
  if (need <= cap && ptr) return OVLB_OK;
  if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
  size_t n = need * slack_num / slack_den + 64;
  cudaError_t e = cudaMalloc((void **)&ptr, n * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) { ovl_set_error("device memory exhausted (" + std::to_string(n * sizeof(T)) + " bytes wanted): use a smaller hash block / ref batch"); return OVLB_ERR_CAPACITY; }
    ovl_set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); return OVLB_ERR_CUDA;
  }
  cap = n;
  return OVLB_OK;
}

struct EvTimer {
  cudaEvent_t a, b; cudaStream_t s;
  EvTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
  float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b); return ms; }
};

//  Upload + encode a read set (host pointers) into `dst`.  `is_hash`: flags are per read, else per (read,dir).
//  slot: 0 = hash block, 1 = ref batch, 2 = the NEXT ref batch (second slot, uploaded while slot 1 runs)
int ovl_upload_reads(ovlb_ctx *c, const ovlb_reads *in, DevReads &dst, int slot, float *upload_ms, float *encode_ms) {
  const bool is_hash = slot == 0;
  bool &pending = slot == 2 ? c->next_pending : c->ref_pending;
  cudaEvent_t ev_ready = slot == 2 ? c->next_ready : c->ref_ready, ev_up0 = slot == 2 ? c->next_up0 : c->ref_up0, ev_up1 = slot == 2 ? c->next_up1 : c->ref_up1;
  if (!in || (in->n_reads && (!in->byte_offset || !in->len))) { ovl_set_error("ovl_upload_reads: null argument"); return OVLB_ERR_ARG; }
  const uint32_t n = in->n_reads;
  std::vector<uint64_t> woff(n + 1), pbase(n + 1);
  uint64_t nw = 0, np = 0, tb = 0, nwin = 0; uint32_t maxlen = 0;
  const uint32_t K = c->P.kmer_len;
  for (uint32_t i = 0; i < n; i++) {
    woff[i] = nw; pbase[i] = np;
    uint32_t L = in->len[i];
    if (L > OVLB_MAX_READLEN) { ovl_set_error("read longer than AS_MAX_READLEN"); return OVLB_ERR_ARG; }
    nw += (uint64_t)(L + 15) / 16 + 2;
    np += ((uint64_t)L + 31) / 32 * 32;
    tb += L;
    if (L >= K) nwin += L - K + 1;
    if (L > maxlen) maxlen = L;
  }
  woff[n] = nw; pbase[n] = np;
  if (np >= 0xFFFFFFF0ull) { ovl_set_error("read set too large for one block (>= 2^32 positions); split it"); return OVLB_ERR_CAPACITY; }
  if (maxlen > c->P.max_read_len) { ovl_set_error("read longer than ovlb_params.max_read_len"); return OVLB_ERR_ARG; }

  int rc;
  if (!is_hash && pending) { CK(cudaStreamSynchronize(c->copy_stream)); pending = false; }   // previous upload into this slot still in flight
  if (nw + 4 > dst.cap_words) {
    if (dst.fwd) cudaFree(dst.fwd); if (dst.rc) cudaFree(dst.rc);
    dst.fwd = dst.rc = nullptr; dst.cap_words = 0;
    size_t want = (size_t)(nw + 4) * 9 / 8 + 1024;
    CK(cudaMalloc((void **)&dst.fwd, want * 8));
    CK(cudaMalloc((void **)&dst.rc, want * 8));
    dst.cap_words = want;
  }
  if ((size_t)n + 1 > dst.cap_reads) {
    if (dst.woff) cudaFree(dst.woff); if (dst.len) cudaFree(dst.len); if (dst.pbase) cudaFree(dst.pbase); if (dst.flags) cudaFree(dst.flags);
    size_t want = (size_t)(n + 1) * 9 / 8 + 64;
    CK(cudaMalloc((void **)&dst.woff, want * 8));
    CK(cudaMalloc((void **)&dst.len, want * 4));
    CK(cudaMalloc((void **)&dst.pbase, want * 8));
    CK(cudaMalloc((void **)&dst.flags, want * 8));       // 2 x uint32 per read
    dst.cap_reads = want;
  }
  dst.n = n; dst.first_id = in->first_read_id; dst.total_bases = tb; dst.n_words = nw; dst.n_pos = np; dst.n_windows = nwin; dst.max_len = maxlen;

  //  Hash side: on the compute stream, synchronous (the index build follows).  Ref side: on the copy stream, asynchronous
  //  when the caller's buffers are page-locked -- it overlaps the index build or the previous batch's host work;
  //  ovlb_run_staged makes the compute stream wait for `ref_ready`.
  ovlb_ctx::Staging &S = slot == 2 ? c->stg_next : c->stg[is_hash ? 0 : 1];
  cudaStream_t st = is_hash ? c->stream : c->copy_stream;
  S.h_woff.swap(woff); S.h_pbase.swap(pbase);
  if ((rc = ensure(S.d_packed, S.packed_cap, (size_t)in->packed_bytes + 16))) return rc;
  if ((rc = ensure(S.d_boff, S.boff_cap, (size_t)n + 1))) return rc;
  if (in->n_n > S.nn_cap) {
    if (S.d_nread) cudaFree(S.d_nread); if (S.d_npos) cudaFree(S.d_npos);
    S.nn_cap = in->n_n * 5 / 4 + 64;
    CK(cudaMalloc((void **)&S.d_nread, S.nn_cap * 4));
    CK(cudaMalloc((void **)&S.d_npos, S.nn_cap * 4));
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (is_hash) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); CK(cudaEventRecord(e0, st)); }
  else CK(cudaEventRecord(ev_up0, st));
  //  the small per-read arrays first (from pageable memory these copies are synchronous: they must not queue behind
  //  the large one), then the packed bases, asynchronous when the caller page-locked them
  if (n) {
    CK(cudaMemcpyAsync(S.d_boff, in->byte_offset, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dst.len, in->len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(dst.woff, S.h_woff.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dst.pbase, S.h_pbase.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (in->packed_bytes) CK(cudaMemcpyAsync(S.d_packed, in->packed, in->packed_bytes, cudaMemcpyHostToDevice, st));
  if (in->n_n) {
    CK(cudaMemcpyAsync(S.d_nread, in->n_read, in->n_n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_npos, in->n_pos, in->n_n * 4, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(dst.flags, 0, (size_t)(n + 1) * 8, st));
  if (is_hash) CK(cudaEventRecord(e1, st));
  if (n) {
    const uint32_t *d_src_len = nullptr, *d_clear = nullptr;
    if (in->src_len) {                                 // blobs as stored: zero the words, then compress / trim on the device
      if ((rc = ensure(S.d_srclen, S.srclen_cap, (size_t)2 * n + 2))) return rc;
      CK(cudaMemcpyAsync(S.d_srclen, in->src_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
      if (in->clear_bgn) CK(cudaMemcpyAsync(S.d_srclen + n, in->clear_bgn, (size_t)n * 4, cudaMemcpyHostToDevice, st));
      else CK(cudaMemsetAsync(S.d_srclen + n, 0, (size_t)n * 4, st));
      d_src_len = S.d_srclen; d_clear = S.d_srclen + n;
      CK(cudaMemsetAsync(dst.fwd, 0, (size_t)(nw + 4) * 8, st));
    }
    k_encode_fwd<<<div_up(n, WARPS_PER_BLOCK), THREADS, 0, st>>>(S.d_packed, S.d_boff, dst.len, dst.woff, dst.fwd, n, d_src_len); c->launches++;
    if (d_src_len) {
      k_encode_raw<<<div_up(n, WARPS_PER_BLOCK), THREADS, 0, st>>>(S.d_packed, S.d_boff, dst.len, dst.woff, dst.fwd, n, d_src_len, d_clear, (int)in->homopoly_compress);
      c->launches++;
    }
    if (in->n_n) { k_apply_n<<<div_up(in->n_n, 256), 256, 0, st>>>(S.d_nread, S.d_npos, in->n_n, dst.woff, dst.len, dst.fwd); c->launches++; }
    k_encode_rc<<<div_up(n, WARPS_PER_BLOCK), THREADS, 0, st>>>(dst.fwd, dst.len, dst.woff, dst.rc, n); c->launches++;
  }
  CK(cudaGetLastError());
  if (is_hash) {
    CK(cudaEventRecord(e2, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1); if (upload_ms) *upload_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2); if (encode_ms) *encode_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  } else {
    CK(cudaEventRecord(ev_up1, st));
    CK(cudaEventRecord(ev_ready, st));
    pending = true;
  }
  (void)is_hash;
  return OVLB_OK;
}

static int ensure_groups(ovlb_ctx *c, DevReads &d) {
  int rc;
  if ((rc = ensure(d.grp_read, d.cap_groups, (size_t)(d.n_pos / 32) + 1))) return rc;
  if (d.n) { k_fill_groups<<<div_up(d.n, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(d.pbase, d.grp_read, d.n); c->launches++; }
  return OVLB_OK;
}

int ovl_build_index(ovlb_ctx *c) {
  DevReads &H = c->hash;
  DevIndex &X = c->index;
  int rc;
  const int K = (int)c->P.kmer_len;
  const uint64_t n_groups = H.n_pos / 32;
  const uint64_t n = H.n_pos;
  const uint64_t sentinel = 1ull << (2 * K + 3);

  if ((rc = ensure_groups(c, H))) return rc;
  {
    const char *ev = getenv("OVLB_HT_FPMASK");                          // read per build: a test narrows it for one case
    c->ht_fpmask = ev ? (uint32_t)strtoul(ev, nullptr, 0) : 0xFFFFFFFFu;
    CK(cudaMemcpyToSymbolAsync(g_ht_fpmask, &c->ht_fpmask, 4, 0, cudaMemcpyHostToDevice, c->stream));
  }

  //  Bucketed build (default) or sorted build (OVLB_BUCKETED=0, or after a bucket overflowed)
  bool bucketed = true;
  if (const char *ev = getenv("OVLB_BUCKETED")) bucketed = atoi(ev) != 0;   // read per build: the tests run both builds in one process
  const uint64_t mixc = 0x9E3779B97F4A7C15ull;
  uint64_t mix_inv = mixc;                                              // inverse mod 2^64 by Newton iteration
  for (int i = 0; i < 6; i++) mix_inv *= 2 - mixc * mix_inv;
  bool tmp_ready = false;

  for (int attempt = 0; attempt < 2; attempt++) {
    if (bucketed && n < (1u << 16)) bucketed = false;                     // tiny block: the sorted build
    //  bucket bits from the position count (an upper bound of the tuple count); the 8-byte bucket tuple needs the key
    //  bits below the bucket number and the position to share 64 bits: 2K + 3 - B + posbits <= 64 (K <= 24 always fits)
    int B = 2 * K + 3 < 8 ? 2 * K + 3 : 8; while (B < 2 * K && (n >> B) > BK_TARGET) B++;
    int posbits = 1; while (posbits < 32 && (n >> posbits)) posbits++;
    if (bucketed && (2 * K + 3 - B + posbits > 64 || B > 20 || B < 2)) bucketed = false;
    //  the bucketed build sizes its slot scratch before it knows the number of distinct k-mers: for every tuple distinct
    //  (a block of a large job covers its part of the genome less than once) if that fits a third of the memory budget,
    //  else for 3/4 of them; if that does not fit either use the sorted build, which counts first
    uint64_t tmp_cap = n + (1u << 20);
    if (tmp_cap * (sizeof(IndexSlot) + 8) > c->mem_budget / 3) tmp_cap = n / 4 * 3 + (1u << 20);
    if (bucketed && tmp_cap * (sizeof(IndexSlot) + 8) > c->mem_budget / 3) bucketed = false;

    if (bucketed) {
      const int B1 = (B + 1) / 2, B2 = B - B1;
      const uint32_t nd1 = 1u << B1, nb = 1u << B;
      const uint64_t mean1 = n >> B1;
      uint64_t cap1 = mean1 + 16 * (uint64_t)sqrt((double)mean1) + PART_CH;
      cap1 = (cap1 + 1) & ~1ull;
      const uint64_t n_words = (n + 31) / 32 + 1, n_bmblk = (n_words + BM_BLOCK - 1) / BM_BLOCK;
      if ((rc = ensure(X.tkey, X.tkey_cap, (size_t)(nd1 * cap1) * 2 + 32, 1, 1))) return rc;   // 16-byte tuple records (key, position)
      if ((rc = ensure(X.tkey2, X.tkey2_cap, (size_t)nb * BK_CAP + 32, 1, 1))) return rc;      // the 8-byte bucket tuples
      if ((rc = ensure(X.occ, X.occ_cap, (size_t)n + 32))) return rc;
      if ((rc = ensure(X.tmp_slots, X.tmp_cap, (size_t)tmp_cap + 1, 1, 1))) return rc;
      if ((rc = ensure(X.gk, X.gk_cap, (size_t)tmp_cap + 1, 1, 1))) return rc;                 // first position of every distinct k-mer
      if ((rc = ensure(X.gv, X.gv_cap, (size_t)(2 * n_words + 2 * n_bmblk + 8), 1, 1))) return rc;   // bitmap | word prefix | block sums | block offsets
      //  small integer scratch: cnt1[nd1] | cnt2[nb] | offs[nb + 1]
      const size_t iscr = (size_t)nd1 * CNT1_STRIDE + nb + nb + 1 + 16;
      if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, iscr * 4 + 256))) return rc;
      unsigned int *cnt1 = reinterpret_cast<unsigned int *>(c->cub_temp), *cnt2 = cnt1 + (size_t)nd1 * CNT1_STRIDE;
      uint32_t *offs = cnt2 + nb;
      uint32_t *bitmap = X.gv, *wprefix = bitmap + n_words, *bsum = wprefix + n_words, *boff = bsum + n_bmblk;
      CK(cudaMemsetAsync(cnt1, 0, ((size_t)nd1 * CNT1_STRIDE + nb) * 4, c->stream));
      CK(cudaMemsetAsync(bitmap, 0, n_words * 4, c->stream));
      CK(cudaMemsetAsync(&c->d_work[5], 0, 24, c->stream));              // [5] distinct, [6] overflow flags, [7] partition overflow
      if (!c->bucket_attr_set) {
        CK(cudaFuncSetAttribute(k_bucket_group2<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK2_SMEM_OF(11)));
        CK(cudaFuncSetAttribute(k_bucket_group2<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK2_SMEM_OF(12)));
        c->bucket_attr_set = true;
      }
      const int sh1 = 2 * K + 3 - B1, sh2 = 2 * K + 3 - B;
      const uint64_t lowmask = (1ull << sh2) - 1;

      EvTimer t1(c->stream);
      k_part1<false><<<div_up(n_groups, PART1_GROUPS), PART_THREADS, 0, c->stream>>>(H.fwd, H.woff, H.len, H.pbase, H.grp_read, n_groups, K, mixc,
                                                                                     sh1, nd1, cap1, reinterpret_cast<uint4 *>(X.tkey), cnt1, &c->d_work[7], 0, 0u, Spill{nullptr, nullptr, 0, nullptr, 0, nullptr});
      c->launches++;
      CK(cudaGetLastError());
      c->timings.index_tuples_ms = t1.stop();

      EvTimer t2(c->stream);
      k_part2<<<dim3(div_up(cap1, PART_CH), nd1), PART2_THREADS, 0, c->stream>>>(reinterpret_cast<const uint4 *>(X.tkey), cnt1, cap1, sh2, B2, lowmask, posbits, BK_CAP,
                                                                                X.tkey2, cnt2, &c->d_work[7], Spill{nullptr, nullptr, 0, nullptr, 0, nullptr});
      k_bucket_scan<<<1, 1024, 0, c->stream>>>(cnt2, nb, BK_CAP, offs);
      c->launches += 2;
      CK(cudaGetLastError());
      c->timings.index_sort_ms = t2.stop();

      EvTimer t3(c->stream);
      unsigned long long h3[3] = {0, 0, 0}; uint32_t n_occ32 = 0;
      //  the small table first (two CTAs per SM) unless the last block needed the large one; a bucket with more distinct
      //  k-mers than it takes -- a block that covers its genome only a few times -- reruns this one kernel with TB = 12
      for (int tb = c->bucket_tb; tb <= 12; tb++) {
        if (tb == 11) k_bucket_group2<11><<<2 * c->sm_count, BK_THREADS, BK2_SMEM_OF(11), c->stream>>>(X.tkey2, cnt2, offs, nb, K, B, posbits, mix_inv, X.occ, X.tmp_slots,
                                                                         (uint32_t)std::min<uint64_t>(tmp_cap, 0xFFFFFFFFull), X.gk, bitmap, &c->d_work[5]);
        else          k_bucket_group2<12><<<c->sm_count, BK_THREADS, BK2_SMEM_OF(12), c->stream>>>(X.tkey2, cnt2, offs, nb, K, B, posbits, mix_inv, X.occ, X.tmp_slots,
                                                                         (uint32_t)std::min<uint64_t>(tmp_cap, 0xFFFFFFFFull), X.gk, bitmap, &c->d_work[5]);
        c->launches++;
        CK(cudaMemcpyAsync(h3, &c->d_work[5], 24, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(&n_occ32, offs + nb, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
        if (h3[1] != 1 || h3[2] != 0 || tb == 12) break;
        c->bucket_tb = 12;
        CK(cudaMemsetAsync(bitmap, 0, n_words * 4, c->stream));
        CK(cudaMemsetAsync(&c->d_work[5], 0, 16, c->stream));
      }
      c->timings.index_table_ms = t3.stop();
      if (h3[1] != 0 || h3[2] != 0) { bucketed = false; continue; }       // a partition or bucket did not fit: redo with the sorted build
      X.n_distinct = h3[0];
      X.n_occ = n_occ32;
      X.bucketed = true;
      tmp_ready = true;
      break;
    }
    X.bucketed = false;
    {
      EvTimer t1(c->stream);
      if ((rc = ensure(X.tkey, X.tkey_cap, (size_t)n + 32))) return rc;
      if ((rc = ensure(X.tkey2, X.tkey2_cap, (size_t)n + 32))) return rc;
      if ((rc = ensure(X.tval, X.tval_cap, (size_t)n + 32))) return rc;
      if ((rc = ensure(X.occ, X.occ_cap, (size_t)n + 32))) return rc;
      if (n_groups) {
        k_hash_tuples<<<div_up(n_groups, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(H.fwd, H.woff, H.len, H.pbase, H.grp_read, n_groups, K, X.tkey, X.tval);
        c->launches++;
      }
      CK(cudaGetLastError());
      c->timings.index_tuples_ms = t1.stop();
    }

    EvTimer t2(c->stream);
    if (n) {
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, X.tkey, X.tkey2, X.tval, X.occ, (int64_t)n, 0, 2 * K + 4, c->stream);
      if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
      size_t tb2 = c->cub_temp_cap;
      CK(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb2, X.tkey, X.tkey2, X.tval, X.occ, (int64_t)n, 0, 2 * K + 4, c->stream));
      c->launches += 2 + (2 * K + 4 + 7) / 8;
    }
    CK(cudaGetLastError());
    c->timings.index_sort_ms = t2.stop();

    unsigned long long h2[2] = {0, 0};
    CK(cudaMemsetAsync(&c->d_work[5], 0, 16, c->stream));
    if (n) {
      k_count_distinct<<<div_up(n, 256 * CNT_PER_THREAD), 256, 0, c->stream>>>(X.tkey2, n, sentinel, &c->d_work[5]);
      c->launches++;
    }
    CK(cudaMemcpyAsync(h2, &c->d_work[5], 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    X.n_distinct = h2[0];
    X.n_occ = h2[1];
    break;
  }

  EvTimer t3(c->stream);
  //  slots in discovery order + the key of the path order (first position), then the path sort and the gather
  std::vector<uint64_t> skip(c->skip_keys);
  std::sort(skip.begin(), skip.end());
  skip.erase(std::unique(skip.begin(), skip.end()), skip.end());
  const uint64_t nd = X.n_distinct, ns = nd + skip.size();
  if (ns >= 0xFFFFFFF0ull) { ovl_set_error("too many distinct k-mers for one index (>= 2^32); use a smaller hash block"); return OVLB_ERR_CAPACITY; }
  uint64_t hcap = (4 * ns + 64) & ~3ull;                   // 8-byte entries in buckets of four, a quarter of them used
  if (const char *ev = getenv("OVLB_HT_PERCENT")) {        // tests: a table filled to `percent` (e.g. 90) makes most buckets overflow into the next ones
    const long pc = atol(ev);
    if (pc >= 25 && pc <= 98) hcap = (ns * 100 / (uint64_t)pc + 8) & ~3ull;
  }
  if ((rc = ensure(X.slots, X.slots_cap, (size_t)ns + 1, 9, 8))) return rc;
  if (!tmp_ready) {
    if ((rc = ensure(X.tmp_slots, X.tmp_cap, (size_t)nd + 1, 9, 8))) return rc;
    if ((rc = ensure(X.gk, X.gk_cap, (size_t)nd + 1, 9, 8))) return rc;
    if ((rc = ensure(X.gv, X.gv_cap, (size_t)nd + 1, 9, 8))) return rc;
  }
  if ((rc = ensure(X.htab, X.htab_cap, (size_t)hcap, 9, 8))) return rc;
  if (!tmp_ready) {
    if ((rc = ensure(X.gk2, X.gk2_cap, (size_t)nd + 1, 9, 8))) return rc;
    if ((rc = ensure(X.gv2, X.gv2_cap, (size_t)nd + 1, 9, 8))) return rc;
  }
  X.hcap = hcap; X.n_slots = (uint32_t)nd;
  CK(cudaMemsetAsync(X.htab, 0xFF, hcap * sizeof(HashEntry), c->stream));
  CK(cudaMemsetAsync(&c->d_work[5], 0, 16, c->stream));
  if (X.n_occ && tmp_ready) {
    //  path order from the bitmap of first positions (set by k_bucket_group2): prefix popcount, then every distinct k-mer
    //  places its own slot and publishes it in the hash table
    const uint64_t n_words = (n + 31) / 32 + 1, n_bmblk = (n_words + BM_BLOCK - 1) / BM_BLOCK;
    uint32_t *bitmap = X.gv, *wprefix = bitmap + n_words, *bsum = wprefix + n_words, *boff = bsum + n_bmblk;
    k_bm_blocksum<<<(unsigned)n_bmblk, 256, 0, c->stream>>>(bitmap, n_words, bsum);
    k_bucket_scan<<<1, 1024, 0, c->stream>>>(bsum, (uint32_t)n_bmblk, 0xFFFFFFFFu, boff);
    k_bm_prefix<<<(unsigned)n_bmblk, 256, 0, c->stream>>>(bitmap, n_words, boff, wprefix);
    k_path_slots2<<<div_up(nd, 256), 256, 0, c->stream>>>(X.tmp_slots, X.gk, (uint32_t)nd, bitmap, wprefix, X.slots, X.htab, hcap, K);
    c->launches += 4;
  } else if (X.n_occ) {
    if (!tmp_ready) {
      k_group_heads<<<div_up(X.n_occ, 256), 256, 0, c->stream>>>(X.tkey2, X.occ, (uint32_t)X.n_occ, X.tmp_slots, X.gk, X.gv, &c->d_work[5]);
      c->launches++;
    }
    int end_bit = 1; while (end_bit < 32 && (H.n_pos >> end_bit)) end_bit++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, X.gk, X.gk2, X.gv, X.gv2, (int64_t)nd, 0, end_bit, c->stream);
    if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
    size_t tb2 = c->cub_temp_cap;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb2, X.gk, X.gk2, X.gv, X.gv2, (int64_t)nd, 0, end_bit, c->stream));
    c->launches += 2 + (end_bit + 7) / 8;
    k_path_slots<<<div_up(nd, 256), 256, 0, c->stream>>>(X.tmp_slots, X.gv2, (uint32_t)nd, X.slots, X.htab, hcap, K);
    c->launches++;
  }
  CK(cudaGetLastError());
  c->timings.index_table_ms = (tmp_ready ? c->timings.index_table_ms : 0.0f) + t3.stop();
  c->host_counters[CT_HASH_KMERS] += X.n_occ;

  EvTimer t4(c->stream);
  if (!skip.empty()) {
    uint64_t *d_skip = nullptr;
    CK(cudaMalloc((void **)&d_skip, skip.size() * 8));
    CK(cudaMemcpyAsync(d_skip, skip.data(), skip.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(&c->d_work[5], 0, 8, c->stream));
    k_index_skip<<<div_up(skip.size(), 128), 128, 0, c->stream>>>(d_skip, skip.size(), K, X.slots, (uint32_t)nd, X.htab, hcap, &c->d_work[5],
                                                                  X.occ, H.pbase, H.grp_read, H.len, H.flags);
    c->launches++;
    unsigned long long n_extra = 0;
    CK(cudaMemcpyAsync(&n_extra, &c->d_work[5], 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_skip);
    X.n_slots = (uint32_t)(nd + n_extra);
  }
  CK(cudaGetLastError());
  c->timings.index_skip_ms = t4.stop();
  X.built = true;
  return OVLB_OK;
}

//  Device pipeline for the staged ref batch up to and including chaining.
int ovl_seed_ref_batch(ovlb_ctx *c) {
  DevReads &R = c->ref, &H = c->hash;
  DevIndex &X = c->index;
  int rc;
  const int K = (int)c->P.kmer_len;
  const uint64_t n_groups = R.n_pos / 32;

  if (R.n >= (1u << OVL_RUNKEY_REF_BITS)) { ovl_set_error("ref batch has too many reads (max 262143); split it"); return OVLB_ERR_CAPACITY; }
  if (H.n >= (1u << OVL_RUNKEY_HASH_BITS)) { ovl_set_error("hash block has too many reads (max 16777215); split it"); return OVLB_ERR_CAPACITY; }

  //  hit words of both orientations | 80 words of slack | rc_absent bits (k_ref_probe)
  if ((rc = ensure(c->ref_valid, c->ref_valid_cap, (size_t)3 * n_groups + 96))) return rc;
  uint32_t *rc_absent = c->ref_valid + 2 * n_groups + 80;
  if ((rc = ensure_groups(c, R))) return rc;

  //  run and item buffers: sized from the memory budget once; overflow -> OVLB_ERR_CAPACITY
  if (c->run_cap == 0) {
    uint64_t want = c->mem_budget / 12 / 88;               // ~1/12 of the budget over 88 B/run of run-side arrays
    if (want < (1u << 20)) want = 1u << 20;
    if (want > (1ull << 31)) want = 1ull << 31;
    if (const char *ev = getenv("OVLB_RUN_CAP")) { const long long v = atoll(ev); if (v > 0) want = (uint64_t)v; }   // tests: force the overflow path
    size_t cap0 = 0, cap1 = 0, cap2 = 0, cap3 = 0, cap4 = 0, cap5 = 0;
    if ((rc = ensure(c->run_key, cap0, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_val, cap1, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_key2, cap2, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_val2, cap3, want, 1, 1))) return rc;
    if ((rc = ensure(c->item_small, cap4, want, 1, 1))) return rc;
    if ((rc = ensure(c->item_large, cap5, want, 1, 1))) return rc;
    c->run_cap = want;
  }

  CK(cudaMemsetAsync(c->d_work, 0, 40, c->stream));       // [0] n_runs, [1] extend work cursor, [2] n_records, [3] small items, [4] large items
  CK(cudaMemsetAsync(R.flags, 0, (size_t)(R.n + 1) * 8, c->stream));
  CK(cudaMemsetAsync(c->ref_valid + 2 * n_groups, 0, (size_t)(n_groups + 96) * 4, c->stream));  // the run-length scan peeks up to 32 words past the word it needs; rc_absent starts empty

  EvTimer t1(c->stream);
  if (n_groups) {
    int per_sm = 4;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ref_probe<0>, THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    const int pflags = (getenv("OVLB_EXPAND_OLD") == nullptr ? 1 : 0) | (getenv("OVLB_PROBE_SERIAL") == nullptr ? 2 : 0);
    //  forward windows first: they tell the reverse pass what it need not look up
    k_ref_probe<0><<<c->sm_count * per_sm, THREADS, 0, c->stream>>>(
        R.fwd, R.rc, R.woff, R.len, R.pbase, R.grp_read, n_groups, R.n_pos, K, X.slots, X.n_slots, X.htab, X.hcap,
        c->ref_valid, R.flags, c->item_small, c->item_large, c->run_cap, c->d_work, pflags, rc_absent, 0, n_groups);
    k_ref_probe<1><<<c->sm_count * per_sm, THREADS, 0, c->stream>>>(
        R.fwd, R.rc, R.woff, R.len, R.pbase, R.grp_read, n_groups, R.n_pos, K, X.slots, X.n_slots, X.htab, X.hcap,
        c->ref_valid, R.flags, c->item_small, c->item_large, c->run_cap, c->d_work, pflags, rc_absent, n_groups, 2 * n_groups);
    c->launches += 2;
  }
  CK(cudaGetLastError());
  c->timings.probe_ms = t1.stop();
  c->host_counters[CT_REF_KMERS] += R.n_windows * 2;

  EvTimer t2(c->stream);
  if (n_groups) {
    ExpandArgs A;
    A.rfwd = R.fwd; A.rrc = R.rc; A.rwoff = R.woff; A.rlen = R.len; A.rpbase = R.pbase; A.rgrp_read = R.grp_read;
    A.r_npos = R.n_pos; A.ref_first_id = R.first_id;
    A.hfwd = H.fwd; A.hwoff = H.woff; A.hlen = H.len; A.hpbase = H.pbase; A.hgrp_read = H.grp_read; A.hash_first_id = H.first_id;
    A.K = K; A.occ = X.occ; A.ref_valid = c->ref_valid;
    A.run_key = c->run_key; A.run_val = c->run_val; A.run_cap = c->run_cap; A.n_runs = &c->d_work[0];
    const int grid = c->sm_count * 8;
    static const int old_expand = [] { const char *ev = getenv("OVLB_EXPAND_OLD"); return ev ? atoi(ev) : 0; }();
    if (old_expand) {
      k_expand_large<<<grid, 256, 0, c->stream>>>(A, c->item_large, &c->d_work[4], c->run_cap, c->d_counters->v);
      k_expand_small<<<grid, 256, 0, c->stream>>>(A, c->item_small, &c->d_work[3], c->run_cap, c->d_counters->v);
      c->launches += 2;
    } else {
      k_expand_coop<<<grid, 256, 0, c->stream>>>(A, c->item_small, &c->d_work[3], c->item_large, &c->d_work[4], c->run_cap, c->d_counters->v);
      c->launches++;
    }
  }
  CK(cudaGetLastError());
  unsigned long long w5[5] = {0, 0, 0, 0, 0};
  CK(cudaMemcpyAsync(w5, c->d_work, 40, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->timings.expand_ms = t2.stop();
  const unsigned long long nr = w5[0];
  if (w5[3] > c->run_cap || w5[4] > c->run_cap || nr > c->run_cap) {
    ovl_set_error("seed buffer overflow (" + std::to_string(nr) + " runs, " + std::to_string(w5[3]) + "+" + std::to_string(w5[4]) +
                  " head ranges > capacity " + std::to_string(c->run_cap) + "); use a smaller ref batch");
    c->n_runs = 0; c->n_pairs = 0;
    return OVLB_ERR_CAPACITY;
  }
  c->n_runs = nr;
  c->n_pairs = 0;
  c->host_counters[CT_SEED_RUNS] += nr;
  c->timings.sort_ms = 0; c->timings.chain_ms = 0;
  if (nr == 0) return OVLB_OK;

  //  sort runs by (ref, dir, hash, ref start)
  EvTimer t3(c->stream);
  {
    int end_bit = 46;
    uint32_t nref = R.n; while (nref) { end_bit++; nref >>= 1; }
    if (end_bit > 64) end_bit = 64;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, c->run_key, c->run_key2, c->run_val, c->run_val2, (int64_t)nr, 0, end_bit, c->stream);
    if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
    size_t tb2 = c->cub_temp_cap;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb2, c->run_key, c->run_key2, c->run_val, c->run_val2, (int64_t)nr, 0, end_bit, c->stream));
    c->launches += 8;
  }
  CK(cudaGetLastError());
  c->timings.sort_ms = t3.stop();

  EvTimer t4(c->stream);
  //  seed-side arrays (one entry per run)
  if (nr > c->seed_cap) {
    int32_t **arrs[] = { &c->seed_start, &c->seed_off, &c->seed_len, &c->sim_nxt, &c->sim_hits, &c->sim_act, &c->sim_order };
    for (auto a : arrs) { if (*a) cudaFree(*a); *a = nullptr; }
    if (c->seed_alive) cudaFree(c->seed_alive); c->seed_alive = nullptr;
    if (c->pair_flag) cudaFree(c->pair_flag); c->pair_flag = nullptr;
    if (c->pair_idx) cudaFree(c->pair_idx); c->pair_idx = nullptr;
    uint64_t want = nr * 5 / 4 + 1024;
    for (auto a : arrs) CK(cudaMalloc((void **)a, want * 4));
    CK(cudaMalloc((void **)&c->seed_alive, want));
    CK(cudaMalloc((void **)&c->pair_flag, want * 4));
    CK(cudaMalloc((void **)&c->pair_idx, want * 4));
    c->seed_cap = want;
  }
  k_pair_heads<<<div_up(nr, 256), 256, 0, c->stream>>>(c->run_key2, nr, c->pair_flag); c->launches++;
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, c->pair_flag, c->pair_idx, (int64_t)nr, c->stream);
    if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
    size_t tb2 = c->cub_temp_cap;
    CK(cub::DeviceScan::ExclusiveSum(c->cub_temp, tb2, c->pair_flag, c->pair_idx, (int64_t)nr, c->stream));
    c->launches += 2;
  }
  uint32_t last_idx = 0, last_flag = 0;
  CK(cudaMemcpyAsync(&last_idx, c->pair_idx + (nr - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&last_flag, c->pair_flag + (nr - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const uint64_t np = (uint64_t)last_idx + last_flag;
  if (np > c->pair_cap) {
    if (c->pairs) cudaFree(c->pairs); c->pairs = nullptr;
    uint64_t want = np * 5 / 4 + 1024;
    CK(cudaMalloc((void **)&c->pairs, want * sizeof(PairRec)));
    c->pair_cap = want;
  }
  c->n_pairs = np;
  //  the unpacked runs reuse run_key (24 B/run of key+val space is enough for 12 B OvlRun)
  OvlRun *runs = reinterpret_cast<OvlRun *>(c->run_key);
  //  run_key is an input of nothing after the sort (sorted data is in run_key2/run_val2), but OvlRun[nr] needs 12*nr <= 8*cap: guaranteed if nr <= 2/3 cap
  if (nr * 12 > c->run_cap * 8) runs = nullptr;
  if (!runs) { if ((rc = ensure(c->runs_extra, c->runs_extra_cap, (size_t)nr))) return rc; runs = c->runs_extra; }
  k_pair_scatter<<<div_up(nr, 256), 256, 0, c->stream>>>(c->run_key2, c->run_val2, nr, c->pair_flag, c->pair_idx, c->pairs, runs); c->launches++;
  k_chain_pairs<<<div_up(np, 128), 128, 0, c->stream>>>(c->pairs, np, nr, runs, c->sim_nxt, c->sim_hits, c->sim_act, c->sim_order,
                                                         c->seed_start, c->seed_off, c->seed_len, c->seed_alive, c->dp,
                                                         R.len, R.flags, H.len, H.flags, c->d_counters->v);
  c->launches++;
  c->pair_order = nullptr;
  if (nr > 3 * np && np > 1 && np < 0xFFFFFFFFull) {      // many seeds per pair: noisy reads, pair costs differ by orders of magnitude
    uint32_t *k1 = c->pair_flag, *v1 = c->pair_idx, *k2 = (uint32_t *)c->sim_nxt, *v2 = (uint32_t *)c->sim_hits;   // all free now, >= nr entries
    k_pair_cost<<<div_up(np, 256), 256, 0, c->stream>>>(c->pairs, (uint32_t)np, k1, v1); c->launches++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k1, k2, v1, v2, (int64_t)np, 0, 32, c->stream);
    if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
    size_t tb2 = c->cub_temp_cap;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb2, k1, k2, v1, v2, (int64_t)np, 0, 32, c->stream));
    c->launches += 6;
    c->pair_order = v2;
  }
  CK(cudaGetLastError());
  c->timings.chain_ms = t4.stop();
  return OVLB_OK;
}


//  k-mer census over the hash reads currently loaded (ovlb_kmer_census): see k_bucket_census.
//  stats[0] distinct k-mers with count >= 2, [1] their total occurrences, [2] k-mers with count 1, [3] the threshold used.
int ovl_kmer_census(ovlb_ctx *c, uint32_t slice_bits, double distinct_fraction, uint64_t min_count,
                    uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n_out, uint64_t stats[4]) {
  NvtxRange nvtx_("ovlb_kmer_census");
  DevReads &H = c->hash;
  DevIndex &X = c->index;
  const int K = (int)c->P.kmer_len;
  const uint64_t n_groups = H.n_pos / 32, n = H.n_pos;
  int rc;
  if (slice_bits > 8) { ovl_set_error("ovlb_kmer_census: at most 256 slices"); return OVLB_ERR_ARG; }
  *n_out = 0;
  if (n == 0) { if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0; return OVLB_OK; }
  if ((rc = ensure_groups(c, H))) return rc;
  const uint64_t per_slice = (n >> slice_bits) + 1;
  int B = 2 * K + 3 < 8 ? 2 * K + 3 : 8; while (B < 2 * K && (per_slice >> B) > BK_TARGET) B++;
  if (B < 2) B = 2;
  const int posbits = 1;                                                  // census tuples carry no position
  if (2 * K + 3 - B + posbits > 64 || B > 20) { ovl_set_error("ovlb_kmer_census: block too large for one slice; use more slices"); return OVLB_ERR_CAPACITY; }
  const int B1 = (B + 1) / 2, B2 = B - B1;
  const uint32_t nd1 = 1u << B1, nb = 1u << B;
  const uint64_t mean1 = per_slice >> B1;
  uint64_t cap1 = mean1 + 16 * (uint64_t)sqrt((double)mean1) + PART_CH; cap1 = (cap1 + 1) & ~1ull;
  if ((rc = ensure(X.tkey, X.tkey_cap, (size_t)(nd1 * cap1) * 2 + 32, 1, 1))) return rc;
  if ((rc = ensure(X.tkey2, X.tkey2_cap, (size_t)nb * BK_CAP + 32, 1, 1))) return rc;
  const size_t iscr = (size_t)nd1 * CNT1_STRIDE + nb + 16;
  if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, iscr * 4 + 256))) return rc;
  unsigned int *cnt1 = reinterpret_cast<unsigned int *>(c->cub_temp), *cnt2 = cnt1 + (size_t)nd1 * CNT1_STRIDE;
  unsigned long long *ghist = nullptr; uint64_t *okey = nullptr; uint32_t *ocnt = nullptr;
  uint64_t scap = 1ull << 22; while (scap < per_slice / 16) scap <<= 1;
  Spill S; S.mask = scap - 1; S.sh2 = 2 * K + 3 - B; S.flag = &c->d_work[6];
  CK(cudaMalloc((void **)&S.key, scap * 8));
  CK(cudaMalloc((void **)&S.cnt, scap * 4));
  CK(cudaMalloc((void **)&S.bucket_spill, (size_t)nb * 4));
  CK(cudaMalloc((void **)&ghist, CENSUS_HMAX * 8));
  CK(cudaMemsetAsync(ghist, 0, CENSUS_HMAX * 8, c->stream));
  const uint64_t mixc = 0x9E3779B97F4A7C15ull;
  uint64_t mix_inv = mixc; for (int i = 0; i < 6; i++) mix_inv *= 2 - mixc * mix_inv;
  static bool attr_set = false;
  if (!attr_set) { CK(cudaFuncSetAttribute(k_bucket_census, cudaFuncAttributeMaxDynamicSharedMemorySize, CENSUS_SMEM)); attr_set = true; }
  const int sh1 = 2 * K + 3 - B1, sh2 = 2 * K + 3 - B;
  const uint64_t lowmask = (1ull << sh2) - 1;
  std::vector<unsigned long long> hist(CENSUS_HMAX, 0);
  uint32_t thr = 0;
  int result = OVLB_OK;
  for (int mode = 0; mode < 2 && result == OVLB_OK; mode++) {
    if (mode == 1) {
      //  threshold: smallest count v with #{distinct k-mers of count in [2, v]} >= fraction x #{distinct, count >= 2}
      //  (merylOp-nextMer.C:103-115 over the `greater-than 1` database), AND count >= min_count
      CK(cudaMemcpyAsync(hist.data(), ghist, CENSUS_HMAX * 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      unsigned long long nd = 0, tot = 0;
      for (uint32_t v = 2; v < CENSUS_HMAX; v++) { nd += hist[v]; tot += hist[v] * v; }
      uint64_t t_d = 0;
      if (distinct_fraction >= 0.0) {
        const unsigned long long target = (unsigned long long)(distinct_fraction * (double)nd);
        unsigned long long cum = 0;
        for (uint32_t v = 2; v < CENSUS_HMAX; v++) { if (!hist[v]) continue; cum += hist[v]; if (cum >= target) { t_d = v; break; } }
      }
      uint64_t t = std::max<uint64_t>(t_d, min_count);
      if (t < 2) t = 2;
      if (t >= CENSUS_HMAX) t = CENSUS_HMAX - 1;      // counts are clamped in the histogram only; emission compares real counts
      thr = (uint32_t)t;
      if (stats) { stats[0] = nd; stats[1] = tot; stats[2] = hist[1]; stats[3] = thr; }
      CK(cudaMalloc((void **)&okey, (cap + 1) * 8));
      CK(cudaMalloc((void **)&ocnt, (cap + 1) * 4));
    }
    CK(cudaMemsetAsync(&c->d_work[5], 0, 24, c->stream));
    for (uint32_t slice = 0; slice < (1u << slice_bits) && result == OVLB_OK; slice++) {
      CK(cudaMemsetAsync(cnt1, 0, ((size_t)nd1 * CNT1_STRIDE + nb) * 4, c->stream));
      CK(cudaMemsetAsync(S.key, 0xFF, scap * 8, c->stream));
      CK(cudaMemsetAsync(S.cnt, 0, scap * 4, c->stream));
      CK(cudaMemsetAsync(S.bucket_spill, 0, (size_t)nb * 4, c->stream));
      k_part1<true><<<div_up(n_groups, PART1_GROUPS), PART_THREADS, 0, c->stream>>>(H.fwd, H.woff, H.len, H.pbase, H.grp_read, n_groups, K, mixc,
                                                                                    sh1, nd1, cap1, reinterpret_cast<uint4 *>(X.tkey), cnt1, &c->d_work[7], (int)slice_bits, slice, S);
      k_part2<<<dim3(div_up(cap1, PART_CH), nd1), PART2_THREADS, 0, c->stream>>>(reinterpret_cast<const uint4 *>(X.tkey), cnt1, cap1, sh2, B2, lowmask, posbits, BK_CAP,
                                                                                X.tkey2, cnt2, &c->d_work[7], S);
      k_bucket_census<<<2 * c->sm_count, BK_THREADS, CENSUS_SMEM, c->stream>>>(X.tkey2, cnt2, nb, K, B, posbits, mix_inv, mode, thr, ghist, okey, ocnt, cap, &c->d_work[5], S);
      k_spill_scan<<<div_up(scap, 256), 256, 0, c->stream>>>(S, K, mix_inv, mode, thr, ghist, okey, ocnt, cap, &c->d_work[5]);
      c->launches += 4;
      unsigned long long h3[3] = {0, 0, 0};
      CK(cudaMemcpyAsync(h3, &c->d_work[5], 24, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      CK(cudaGetLastError());
      if (h3[2] != 0 || (h3[1] & 5)) { ovl_set_error("ovlb_kmer_census: the spill table of high-copy k-mers is full; use more slices"); result = OVLB_ERR_CAPACITY; }
      else if (h3[1] & 2) { *n_out = h3[0]; ovl_set_error("ovlb_kmer_census: output buffer too small"); result = OVLB_ERR_CAPACITY; }
      else if (mode == 1) *n_out = h3[0];
    }
  }
  if (result == OVLB_OK && *n_out) {
    CK(cudaMemcpyAsync(kmers, okey, *n_out * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(counts, ocnt, *n_out * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  cudaFree(ghist); if (okey) cudaFree(okey); if (ocnt) cudaFree(ocnt);
  cudaFree(S.key); cudaFree(S.cnt); cudaFree(S.bucket_spill);
  return result;
}
