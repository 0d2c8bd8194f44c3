//  ovl_index.cu -- read encoding, k-mer index build, lookup and seed-run emission, chaining.
//
//  Kernels (SURVEY.md 2.1 numbering):
//    K0  encode_fwd / apply_n / encode_rc       sqStore 2-bit bytes -> dp4 words, both orientations
//    K1  index_count / index_fill               Build_Hash_Index + Put_String_In_Hash + Hash_Insert
//                                               (overlapInCore-Build_Hash_Index.C:267-404,415-631)
//    K1b index_skip                             Mark_Skip_Kmers / Hash_Mark_Empty / Mark_Screened_Ends (:98-257)
//    K2a ref_probe                              Find_Overlaps window loop + Hash_Find (Find_Overlaps.C:177-336)
//    K2b ref_expand                             chain walk + Add_Ref, collapsed to maximal diagonal runs
//    K3  pair_heads / runs_unpack / chain_pairs Add_Match replay, hopeless check, --minkmers filter
//                                               (Find_Overlaps.C:26-163, Process_String_Overlaps.C:384-415,581-637)
//
//  All HBM-bound integer work: coalesced streaming of dp4 words and position arrays, random 8/16-byte
//  probes into the open-addressed table, warp-aggregated appends.  No tensor cores.
#include "ovl_ctx.h"

#include <cub/cub.cuh>

#define WARPS_PER_BLOCK 8
#define THREADS (WARPS_PER_BLOCK * 32)

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
//  K0: encoding
// ------------------------------------------------------------------------------------------------

//  one warp per read: packed 2-bit (4 bases/byte, first base in the top bits) -> dp4 forward words
__global__ void __launch_bounds__(THREADS)
k_encode_fwd(const uint8_t *__restrict__ packed, const uint64_t *__restrict__ boff, const uint32_t *__restrict__ len,
             const uint64_t *__restrict__ woff, uint64_t *__restrict__ fwd, uint32_t n) {
  uint32_t r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
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
//  k-mer of the window starting at base p (needs K <= 31): key with base j in bits [2j, 2j+1]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kmer_at(const uint64_t *__restrict__ w, int p, int K, uint64_t &key) {
  uint32_t i0, i1;
  uint64_t c0 = ovl_codes16(ovl_fetch16(w, p), &i0);
  uint64_t c1 = ovl_codes16(ovl_fetch16(w, p + 16), &i1);
  uint64_t all = c0 | (c1 << 32);
  uint32_t inv = i0 | (i1 << 16);
  key = all & ((1ull << (2 * K)) - 1);
  return (inv & ((1u << K) - 1)) == 0;
}

// ------------------------------------------------------------------------------------------------
//  K1: index build
// ------------------------------------------------------------------------------------------------

//  one warp per 32-position group of the hash block; one k-mer per lane
__global__ void __launch_bounds__(THREADS)
k_index_count(const uint64_t *__restrict__ fwd, const uint64_t *__restrict__ woff, const uint32_t *__restrict__ len,
              const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read, uint64_t n_groups,
              int K, uint64_t *keys, uint32_t *cnt, uint64_t mask, uint32_t *__restrict__ slot_of,
              unsigned long long *counters) {
  uint64_t g = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= n_groups) return;
  const int lane = threadIdx.x & 31;
  const uint32_t r = grp_read[g];
  const int L = (int)len[r];
  const int p = (int)(g * 32 - pbase[r]) + lane;
  uint32_t slot = 0xFFFFFFFFu;
  uint64_t key;
  if (p + K <= L && kmer_at(fwd + woff[r], p, K, key)) {
    uint64_t h = ovl_mix64(key) & mask;
    while (true) {
      uint64_t k = keys[h];
      if (k == key) break;
      if (k == OVL_EMPTY_KEY) {
        unsigned long long old = atomicCAS((unsigned long long *)&keys[h], (unsigned long long)OVL_EMPTY_KEY, (unsigned long long)key);
        if (old == OVL_EMPTY_KEY || old == key) break;
      }
      h = (h + 1) & mask;
    }
    atomicAdd(&cnt[h], 1u);
    slot = (uint32_t)h;
  }
  slot_of[g * 32 + lane] = slot;
  unsigned m = __ballot_sync(0xffffffffu, slot != 0xFFFFFFFFu);
  if (lane == 0 && m) atomicAdd(&counters[CT_HASH_KMERS], (unsigned long long)__popc(m));
}

//  one thread per position: append the occurrence to its k-mer's list
__global__ void __launch_bounds__(THREADS)
k_index_fill(const uint32_t *__restrict__ slot_of, const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read,
             uint64_t n_pos, const uint32_t *__restrict__ start, uint32_t *cursor, uint64_t *__restrict__ occ) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pos) return;
  uint32_t s = slot_of[i];
  if (s == 0xFFFFFFFFu) return;
  uint32_t r = grp_read[i >> 5];
  uint32_t p = (uint32_t)(i - pbase[r]);
  uint32_t o = atomicAdd(&cursor[s], 1u);
  occ[(uint64_t)start[s] + o] = ((uint64_t)r << 32) | p;
}

//  one thread per skip k-mer: flag the slot (insert it if absent) and mark screened read ends
__global__ void k_index_skip(const uint64_t *__restrict__ skip, uint64_t n_skip, int K, uint64_t *keys, uint32_t *cnt,
                             const uint32_t *__restrict__ start, uint64_t mask, const uint64_t *__restrict__ occ,
                             const uint32_t *__restrict__ hlen, uint32_t *hflags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_skip) return;
  uint64_t key = skip[i];
  uint64_t h = ovl_mix64(key) & mask;
  while (true) {
    uint64_t k = keys[h];
    if (k == key) break;
    if (k == OVL_EMPTY_KEY) {
      unsigned long long old = atomicCAS((unsigned long long *)&keys[h], (unsigned long long)OVL_EMPTY_KEY, (unsigned long long)key);
      if (old == OVL_EMPTY_KEY || old == key) break;
    }
    h = (h + 1) & mask;
  }
  uint32_t c = atomicOr(&cnt[h], OVL_SKIP_FLAG);
  if (c & OVL_SKIP_FLAG) return;                        // duplicate in the skip list: already handled
  uint32_t n = c & ~OVL_SKIP_FLAG;
  uint64_t st = start[h];
  for (uint32_t j = 0; j < n; j++) {                     // Mark_Screened_Ends_Chain (Build_Hash_Index.C:98-121)
    uint64_t e = occ[st + j];
    uint32_t r = (uint32_t)(e >> 32), q = (uint32_t)e;
    uint32_t f = 0;
    if (q < OVL_HOPELESS_MATCH) f |= 1u;
    if ((int)hlen[r] - (int)q - K + 1 < OVL_HOPELESS_MATCH) f |= 2u;
    if (f) atomicOr(&hflags[r], f);
  }
}

// ------------------------------------------------------------------------------------------------
//  K2a: probe every ref window (both orientations)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
k_ref_probe(const uint64_t *__restrict__ fwd, const uint64_t *__restrict__ rc, const uint64_t *__restrict__ woff,
            const uint32_t *__restrict__ len, const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ grp_read,
            uint64_t n_groups, uint64_t n_pos, int K, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ cnt,
            uint64_t mask, int32_t *__restrict__ ref_slot, uint32_t *__restrict__ ref_valid, uint32_t *rflags,
            unsigned long long *counters) {
  uint64_t gg = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (gg >= 2 * n_groups) return;
  const int lane = threadIdx.x & 31;
  const int dir = gg >= n_groups;
  const uint64_t g = dir ? gg - n_groups : gg;
  const uint32_t r = grp_read[g];
  const int L = (int)len[r];
  const int p = (int)(g * 32 - pbase[r]) + lane;
  const uint64_t *w = (dir ? rc : fwd) + woff[r];
  int32_t slot = -1;
  bool inrange = (p + K <= L);
  uint64_t key;
  if (inrange && kmer_at(w, p, K, key)) {
    uint64_t h = ovl_mix64(key) & mask;
    while (true) {
      uint64_t k = keys[h];
      if (k == key) {
        uint32_t c = cnt[h];
        if (c & OVL_SKIP_FLAG) {                          // hi_hits (Find_Overlaps.C:274-276,310-316)
          uint32_t f = 0;
          if (p == 0) f = 1u;
          else {
            if (p < OVL_HOPELESS_MATCH) f |= 1u;
            if (L - p - K + 1 < OVL_HOPELESS_MATCH) f |= 2u;
          }
          if (f) atomicOr(&rflags[2 * r + dir], f);
        } else if (c != 0) {
          slot = (int32_t)h;
        }
        break;
      }
      if (k == OVL_EMPTY_KEY) break;
      h = (h + 1) & mask;
    }
  }
  ref_slot[(uint64_t)dir * n_pos + g * 32 + lane] = slot;
  unsigned vm = __ballot_sync(0xffffffffu, slot >= 0);
  unsigned im = __ballot_sync(0xffffffffu, inrange);
  if (lane == 0) {
    ref_valid[((uint64_t)dir * n_pos >> 5) + g] = vm;
    if (im) atomicAdd(&counters[CT_REF_KMERS], (unsigned long long)__popc(im));
  }
}

// ------------------------------------------------------------------------------------------------
//  K2b: expand hits, keep only the first hit of every maximal diagonal run, emit (key, value) runs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
k_ref_expand(const uint64_t *__restrict__ rfwd, const uint64_t *__restrict__ rrc, const uint64_t *__restrict__ rwoff,
             const uint32_t *__restrict__ rlen, const uint64_t *__restrict__ rpbase, const uint32_t *__restrict__ grp_read,
             uint64_t n_groups, uint64_t n_pos, uint32_t ref_first_id,
             const uint64_t *__restrict__ hfwd, const uint64_t *__restrict__ hwoff, const uint32_t *__restrict__ hlen,
             uint32_t hash_first_id, int K,
             const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ start, const uint64_t *__restrict__ occ,
             const int32_t *__restrict__ ref_slot, const uint32_t *__restrict__ ref_valid,
             uint64_t *__restrict__ run_key, uint64_t *__restrict__ run_val, uint64_t run_cap,
             unsigned long long *n_runs, unsigned long long *counters) {
  uint64_t gg = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (gg >= 2 * n_groups) return;
  const int lane = threadIdx.x & 31;
  const int dir = gg >= n_groups;
  const uint64_t g = dir ? gg - n_groups : gg;
  const uint64_t vword = ((uint64_t)dir * n_pos >> 5) + g;
  const uint32_t vw = ref_valid[vword];
  if (vw == 0) return;

  const uint32_t r = grp_read[g];
  const int L = (int)rlen[r];
  const int p0 = (int)(g * 32 - rpbase[r]);
  const uint64_t *rw = (dir ? rrc : rfwd) + rwoff[r];
  const uint32_t ref_id = ref_first_id + r;
  const uint32_t prev_top = (p0 > 0) ? (ref_valid[vword - 1] >> 31) : 0u;

  const int32_t my_slot = ref_slot[(uint64_t)dir * n_pos + g * 32 + lane];
  uint32_t my_cnt = 0, my_start = 0;
  if (my_slot >= 0) { my_cnt = cnt[my_slot] & ~OVL_SKIP_FLAG; my_start = start[my_slot]; }

  unsigned long long hits_acc = 0, runs_acc = 0;

  for (uint32_t bits = vw; bits; bits &= bits - 1) {
    const int b = __ffs(bits) - 1;
    const int p = p0 + b;
    const uint32_t c  = __shfl_sync(0xffffffffu, my_cnt, b);
    const uint32_t st = __shfl_sync(0xffffffffu, my_start, b);
    const bool prev_valid = (b > 0) ? ((vw >> (b - 1)) & 1u) : (prev_top != 0);
    const uint32_t ref_prev_nib = (p > 0) ? ovl_nibble(rw, p - 1) : 0u;

    for (uint32_t j0 = 0; j0 < c; j0 += 32) {
      const uint32_t j = j0 + lane;
      bool is_start = false;
      uint32_t hh = 0, q = 0, runlen = 0;
      if (j < c) {
        uint64_t e = occ[(uint64_t)st + j];
        hh = (uint32_t)(e >> 32); q = (uint32_t)e;
        if (ref_id < hash_first_id + hh) {                 // only refID < hashID pairs (Find_Overlaps.C:279,320)
          const uint64_t *hw = hfwd + hwoff[hh];
          is_start = true;
          if (prev_valid && q > 0 && ovl_nibble(hw, (int)q - 1) == ref_prev_nib) is_start = false;
          if (is_start) {
            //  run length: equal bases after the k-mer, and consecutive valid ref windows
            const int HL = (int)hlen[hh];
            int lim = min(L - (p + K), HL - ((int)q + K));
            int e2 = 0;
            while (e2 < lim) {
              int k = ovl_equal16(ovl_fetch16(rw, p + K + e2), ovl_fetch16(hw, (int)q + K + e2));
              e2 += k;
              if (k < 16) break;
            }
            if (e2 > lim) e2 = lim;
            int want = e2 + 1;
            uint32_t rem = vw >> b;
            int nv = __ffs(~rem) - 1;                       // ones run inside this word (<= 32-b)
            if (nv < 0) nv = 32;
            if (nv == 32 - b) {
              uint64_t wi = vword + 1;
              while (nv < want) {
                uint32_t x = ref_valid[wi++];
                if (x == 0xFFFFFFFFu) { nv += 32; continue; }
                nv += __ffs(~x) - 1;
                break;
              }
            }
            runlen = (uint32_t)min(want, nv);
          }
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, is_start);
      if (m) {
        unsigned long long base = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(n_runs, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (is_start) {
          unsigned long long idx = base + __popc(m & ((1u << lane) - 1));
          if (idx < run_cap) {
            run_key[idx] = ovl_runkey(r, (uint32_t)dir, hh, (uint32_t)p);
            run_val[idx] = (uint64_t)q | ((uint64_t)runlen << 32);
          }
          hits_acc += runlen;
          runs_acc += 1;
        }
      }
    }
  }
  //  warp totals
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hits_acc += __shfl_down_sync(0xffffffffu, hits_acc, o);
    runs_acc += __shfl_down_sync(0xffffffffu, runs_acc, o);
  }
  if (lane == 0 && runs_acc) {
    atomicAdd(&counters[CT_SEED_HITS], hits_acc);
    atomicAdd(&counters[CT_SEED_RUNS], runs_acc);
  }
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
  if (e != cudaSuccess) { ovl_set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); cudaGetLastError(); return OVLB_ERR_CUDA; }
  cap = n;
  return OVLB_OK;
}

struct EvTimer {
  cudaEvent_t a, b; cudaStream_t s;
  EvTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
  float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b); return ms; }
};

//  Upload + encode a read set (host pointers) into `dst`.  `is_hash`: flags are per read, else per (read,dir).
int ovl_upload_reads(ovlb_ctx *c, const ovlb_reads *in, DevReads &dst, bool is_hash, float *upload_ms, float *encode_ms) {
  if (!in || (in->n_reads && (!in->byte_offset || !in->len))) { ovl_set_error("ovl_upload_reads: null argument"); return OVLB_ERR_ARG; }
  const uint32_t n = in->n_reads;
  std::vector<uint64_t> woff(n + 1), pbase(n + 1);
  uint64_t nw = 0, np = 0, tb = 0; uint32_t maxlen = 0;
  for (uint32_t i = 0; i < n; i++) {
    woff[i] = nw; pbase[i] = np;
    uint32_t L = in->len[i];
    if (L > OVLB_MAX_READLEN) { ovl_set_error("read longer than AS_MAX_READLEN"); return OVLB_ERR_ARG; }
    nw += (uint64_t)(L + 15) / 16 + 2;
    np += ((uint64_t)L + 31) / 32 * 32;
    tb += L;
    if (L > maxlen) maxlen = L;
  }
  woff[n] = nw; pbase[n] = np;
  if (np >= 0xFFFFFFF0ull) { ovl_set_error("read set too large for one block (>= 2^32 positions); split it"); return OVLB_ERR_CAPACITY; }
  if (maxlen > c->P.max_read_len) { ovl_set_error("read longer than ovlb_params.max_read_len"); return OVLB_ERR_ARG; }

  int rc;
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
  dst.n = n; dst.first_id = in->first_read_id; dst.total_bases = tb; dst.n_words = nw; dst.n_pos = np; dst.max_len = maxlen;

  EvTimer tu(c->stream);
  if ((rc = ensure(c->d_packed, c->packed_cap, (size_t)in->packed_bytes + 16))) return rc;
  if ((rc = ensure(c->d_boff, c->boff_cap, (size_t)n + 1))) return rc;
  if (in->packed_bytes) CK(cudaMemcpyAsync(c->d_packed, in->packed, in->packed_bytes, cudaMemcpyHostToDevice, c->stream));
  if (n) {
    CK(cudaMemcpyAsync(c->d_boff, in->byte_offset, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dst.len, in->len, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  }
  CK(cudaMemcpyAsync(dst.woff, woff.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(dst.pbase, pbase.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (in->n_n) {
    if (in->n_n > c->nn_cap) {
      if (c->d_nread) cudaFree(c->d_nread); if (c->d_npos) cudaFree(c->d_npos);
      c->nn_cap = in->n_n * 5 / 4 + 64;
      CK(cudaMalloc((void **)&c->d_nread, c->nn_cap * 4));
      CK(cudaMalloc((void **)&c->d_npos, c->nn_cap * 4));
    }
    CK(cudaMemcpyAsync(c->d_nread, in->n_read, in->n_n * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_npos, in->n_pos, in->n_n * 4, cudaMemcpyHostToDevice, c->stream));
  }
  CK(cudaMemsetAsync(dst.flags, 0, (size_t)(n + 1) * 8, c->stream));
  CK(cudaStreamSynchronize(c->stream));                  // host vectors woff/pbase go out of scope
  if (upload_ms) *upload_ms = tu.stop(); else tu.stop();

  EvTimer te(c->stream);
  if (n) {
    k_encode_fwd<<<div_up(n, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(c->d_packed, c->d_boff, dst.len, dst.woff, dst.fwd, n); c->launches++;
    if (in->n_n) { k_apply_n<<<div_up(in->n_n, 256), 256, 0, c->stream>>>(c->d_nread, c->d_npos, in->n_n, dst.woff, dst.len, dst.fwd); c->launches++; }
    k_encode_rc<<<div_up(n, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(dst.fwd, dst.len, dst.woff, dst.rc, n); c->launches++;
  }
  CK(cudaGetLastError());
  if (encode_ms) *encode_ms = te.stop(); else te.stop();
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

  //  capacity: at most one distinct k-mer per base, load factor <= 0.5, plus the skip list
  uint64_t want = 2 * (H.total_bases + c->skip_keys.size()) + 1024;
  uint64_t cap = 1024; while (cap < want) cap <<= 1;
  if (cap > X.cap_alloc) {
    if (X.keys) cudaFree(X.keys); if (X.cnt) cudaFree(X.cnt); if (X.start) cudaFree(X.start);
    X.keys = nullptr; X.cnt = nullptr; X.start = nullptr; X.cap_alloc = 0;
    CK(cudaMalloc((void **)&X.keys, cap * 8));
    CK(cudaMalloc((void **)&X.cnt, cap * 4));
    CK(cudaMalloc((void **)&X.start, cap * 4));
    X.cap_alloc = cap;
  }
  X.cap = cap;
  if ((rc = ensure(X.slot_of, X.slot_alloc, (size_t)H.n_pos + 32))) return rc;
  if ((rc = ensure_groups(c, H))) return rc;

  EvTimer t1(c->stream);
  CK(cudaMemsetAsync(X.keys, 0xFF, cap * 8, c->stream));
  CK(cudaMemsetAsync(X.cnt, 0, cap * 4, c->stream));
  if (n_groups) {
    k_index_count<<<div_up(n_groups, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(
        H.fwd, H.woff, H.len, H.pbase, H.grp_read, n_groups, K, X.keys, X.cnt, cap - 1, X.slot_of, c->d_counters->v);
    c->launches++;
  }
  CK(cudaGetLastError());
  c->timings.index_count_ms = t1.stop();

  EvTimer t2(c->stream);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, X.cnt, X.start, (int64_t)cap, c->stream);
  if ((rc = ensure((uint8_t *&)c->cub_temp, c->cub_temp_cap, tb + 256))) return rc;
  size_t tb2 = c->cub_temp_cap;
  CK(cub::DeviceScan::ExclusiveSum(c->cub_temp, tb2, X.cnt, X.start, (int64_t)cap, c->stream));
  c->launches += 2;
  unsigned long long hk = 0;
  CK(cudaMemcpyAsync(&hk, &c->d_counters->v[CT_HASH_KMERS], 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  //  CT_HASH_KMERS is cumulative over the context; the occurrences of THIS block are start[cap-1]+cnt[cap-1]
  uint32_t last_start = 0, last_cnt = 0;
  CK(cudaMemcpy(&last_start, X.start + (cap - 1), 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&last_cnt, X.cnt + (cap - 1), 4, cudaMemcpyDeviceToHost));
  X.n_occ = (uint64_t)last_start + last_cnt;
  c->timings.index_scan_ms = t2.stop();

  if ((rc = ensure(X.occ, X.occ_alloc, (size_t)X.n_occ + 32))) return rc;

  EvTimer t3(c->stream);
  CK(cudaMemsetAsync(X.cnt, 0, cap * 4, c->stream));
  if (H.n_pos) {
    k_index_fill<<<div_up(H.n_pos, THREADS), THREADS, 0, c->stream>>>(X.slot_of, H.pbase, H.grp_read, H.n_pos, X.start, X.cnt, X.occ);
    c->launches++;
  }
  CK(cudaGetLastError());
  c->timings.index_fill_ms = t3.stop();

  EvTimer t4(c->stream);
  if (!c->skip_keys.empty()) {
    uint64_t *d_skip = nullptr;
    CK(cudaMalloc((void **)&d_skip, c->skip_keys.size() * 8));
    CK(cudaMemcpyAsync(d_skip, c->skip_keys.data(), c->skip_keys.size() * 8, cudaMemcpyHostToDevice, c->stream));
    k_index_skip<<<div_up(c->skip_keys.size(), 128), 128, 0, c->stream>>>(d_skip, c->skip_keys.size(), K, X.keys, X.cnt, X.start, cap - 1, X.occ, H.len, H.flags);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_skip);
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

  if ((rc = ensure(c->ref_slot, c->ref_slot_cap, (size_t)2 * R.n_pos + 64))) return rc;
  if ((rc = ensure(c->ref_valid, c->ref_valid_cap, (size_t)2 * n_groups + 8))) return rc;
  if ((rc = ensure_groups(c, R))) return rc;

  CK(cudaMemsetAsync(c->d_work, 0, 64, c->stream));       // [0] n_runs, [1] extend work cursor, [2] n_records
  CK(cudaMemsetAsync(R.flags, 0, (size_t)(R.n + 1) * 8, c->stream));

  EvTimer t1(c->stream);
  if (n_groups) {
    k_ref_probe<<<div_up(2 * n_groups, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(
        R.fwd, R.rc, R.woff, R.len, R.pbase, R.grp_read, n_groups, R.n_pos, K, X.keys, X.cnt, X.cap - 1,
        c->ref_slot, c->ref_valid, R.flags, c->d_counters->v);
    c->launches++;
  }
  CK(cudaGetLastError());
  c->timings.probe_ms = t1.stop();

  //  run buffers: sized from the memory budget once; overflow -> OVLB_ERR_CAPACITY
  if (c->run_cap == 0) {
    uint64_t want = c->mem_budget / 12 / 56;               // ~1/12 of the budget over 56 B/run of run-side arrays
    if (want < (1u << 20)) want = 1u << 20;
    if (want > (1ull << 31)) want = 1ull << 31;
    size_t cap0 = 0, cap1 = 0, cap2 = 0, cap3 = 0;
    if ((rc = ensure(c->run_key, cap0, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_val, cap1, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_key2, cap2, want, 1, 1))) return rc;
    if ((rc = ensure(c->run_val2, cap3, want, 1, 1))) return rc;
    c->run_cap = want;
  }

  EvTimer t2(c->stream);
  if (n_groups) {
    k_ref_expand<<<div_up(2 * n_groups, WARPS_PER_BLOCK), THREADS, 0, c->stream>>>(
        R.fwd, R.rc, R.woff, R.len, R.pbase, R.grp_read, n_groups, R.n_pos, R.first_id,
        H.fwd, H.woff, H.len, H.first_id, K, X.cnt, X.start, X.occ, c->ref_slot, c->ref_valid,
        c->run_key, c->run_val, c->run_cap, &c->d_work[0], c->d_counters->v);
    c->launches++;
  }
  CK(cudaGetLastError());
  unsigned long long nr = 0;
  CK(cudaMemcpyAsync(&nr, &c->d_work[0], 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->timings.expand_ms = t2.stop();
  if (nr > c->run_cap) {
    ovl_set_error("seed-run buffer overflow (" + std::to_string(nr) + " runs > capacity " + std::to_string(c->run_cap) + "); use a smaller ref batch");
    c->n_runs = 0; c->n_pairs = 0;
    return OVLB_ERR_CAPACITY;
  }
  c->n_runs = nr;
  c->n_pairs = 0;
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
  CK(cudaGetLastError());
  c->timings.chain_ms = t4.stop();
  return OVLB_OK;
}
