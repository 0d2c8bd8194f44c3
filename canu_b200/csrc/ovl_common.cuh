//  ovl_common.cuh -- helpers shared by every kernel of the B200 ovl path.
//
//  Everything here is `OVL_HD` (host+device) sequential code so that the same
//  source can be exercised on the CPU by tests/model (g++) and on the GPU (nvcc).
//
//  Device data layout ("dp4"): every read is stored twice in HBM, forward and
//  reverse-complemented, as one-hot nibbles  A=1 C=2 G=4 T=8 N=15 , 16 bases per
//  64-bit word, base j of a read in bits [4*(j%16), 4*(j%16)+3] of word j/16.
//  Two bases "match" in the reference's alignment sense
//  (A==T || A=='n' || T=='n', prefixEditDistance-forward.C:164) iff the AND of
//  their nibbles is non-zero; bases past the end of a read are 0 and match
//  nothing, which terminates every slide without a bounds test.  Reverse
//  complementing 16 bases is a 64-bit bit reversal.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define OVL_HD __host__ __device__ __forceinline__
#else
#define OVL_HD inline
#endif

#define OVL_HOPELESS_MATCH      90      // overlapInCore.H:78
#define OVL_MAX_DISTINCT_OLAPS  3       // overlapInCore.H:107
#define OVL_MIN_INTERSECTION    10      // overlapInCore.H:120
#define OVL_SHIFT_SLACK         1       // overlapInCore.H:146
#define OVL_MIN_BRANCH_END_DIST 20      // prefixEditDistance.C:28

#define OVL_EMPTY_KEY  0xFFFFFFFFFFFFFFFFull
#define OVL_SKIP_FLAG  0x80000000u      // top bit of a slot's count: k-mer is in the skip list ("Empty")

enum { OVL_NONE = 0, OVL_LEFT_BRANCH_PT = 1, OVL_RIGHT_BRANCH_PT = 2, OVL_DOVETAIL = 3 };   // prefixEditDistance.H:35-40

OVL_HD int ovl_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

//  Number of leading (lowest) nibble positions at which a AND b is non-zero, 0..16.
OVL_HD int ovl_match16(uint64_t a, uint64_t b) {
  uint64_t t = a & b;
  uint64_t z = (t - 0x1111111111111111ull) & ~t & 0x8888888888888888ull;   // lowest flagged nibble is exact
  return z ? (ovl_ctz64(z) >> 2) : 16;
}

//  32-bit flavour for the DP inner loop: 8 bases per test (most cells of a noisy band stop within 2 bases).
OVL_HD int ovl_match8(uint32_t a, uint32_t b) {
  uint32_t t = a & b;
  uint32_t z = (t - 0x11111111u) & ~t & 0x88888888u;
#if defined(__CUDA_ARCH__)
  return z ? ((__ffs((int)z) - 1) >> 2) : 8;
#else
  return z ? (__builtin_ctz(z) >> 2) : 8;
#endif
}

//  8 nibbles starting at base x (x >= 0); w32 is the same dp4 storage viewed as 32-bit words.
OVL_HD uint32_t ovl_fetch8(const uint32_t *w32, int x) {
  const uint32_t *p = w32 + (x >> 3);
  const int sh = (x & 7) << 2;
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(p[0], p[1], sh);
#else
  return sh ? ((p[0] >> sh) | (p[1] << (32 - sh))) : p[0];
#endif
}

//  Number of leading nibble positions at which a == b exactly, 0..16 (for seeds: N never seeds).
OVL_HD int ovl_equal16(uint64_t a, uint64_t b) {
  uint64_t x = a ^ b;
  uint64_t m = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x1111111111111111ull;
  return m ? (ovl_ctz64(m) >> 2) : 16;
}

//  16 nibbles starting at base x (x >= 0) of a read whose words start at w; reads words x/16 and x/16+1.
OVL_HD uint64_t ovl_fetch16(const uint64_t *w, int x) {
  const uint64_t *p = w + (x >> 4);
  int sh = (x & 15) << 2;
  uint64_t lo = p[0];
  if (sh == 0) return lo;
  return (lo >> sh) | (p[1] << (64 - sh));
}

OVL_HD uint32_t ovl_nibble(const uint64_t *w, int x) {
  return (uint32_t)(w[x >> 4] >> ((x & 15) << 2)) & 0xFu;
}

//  16 one-hot nibbles -> 16 two-bit codes (A0 C1 G2 T3), base j in bits [2j,2j+1]; and a 16-bit
//  mask of positions that are NOT a plain base (N or past-the-end).
OVL_HD uint32_t ovl_codes16(uint64_t w, uint32_t *invalid) {
  const uint64_t M = 0x1111111111111111ull;
  uint64_t b0 = ((w >> 1) | (w >> 3)) & M;
  uint64_t b1 = ((w >> 2) | (w >> 3)) & M;
  uint64_t x  = b0 | (b1 << 1);
  x = (x | (x >> 2))  & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x >> 4))  & 0x00FF00FF00FF00FFull;
  x = (x | (x >> 8))  & 0x0000FFFF0000FFFFull;
  x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
  uint64_t isn  = (w & (w >> 1) & (w >> 2) & (w >> 3)) & M;      // nibble == 15
  uint64_t is0  = ~(w | (w >> 1) | (w >> 2) | (w >> 3)) & M;     // nibble == 0
  uint64_t v = isn | is0;
  v = (v | (v >> 3))  & 0x0303030303030303ull;
  v = (v | (v >> 6))  & 0x000F000F000F000Full;
  v = (v | (v >> 12)) & 0x000000FF000000FFull;
  v = (v | (v >> 24)) & 0x000000000000FFFFull;
  *invalid = (uint32_t)v;
  return (uint32_t)x;
}

OVL_HD uint64_t ovl_mix64(uint64_t k) {         // murmur3 finaliser: slot hash of a k-mer key
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

//  Sort key of one seed run: | ref index:18 | dir:1 | hash index:24 | ref start:21 |
#define OVL_RUNKEY_POS_BITS   21
#define OVL_RUNKEY_HASH_BITS  24
#define OVL_RUNKEY_REF_BITS   18
OVL_HD uint64_t ovl_runkey(uint32_t ref_idx, uint32_t dir, uint32_t hash_idx, uint32_t pos) {
  return ((uint64_t)ref_idx << 46) | ((uint64_t)dir << 45) | ((uint64_t)hash_idx << 21) | pos;
}
OVL_HD uint64_t ovl_runkey_pair(uint64_t key) { return key >> 21; }

//  ---------------------------------------------------------------------------------------------
//  Seed-list construction at RUN granularity.
//
//  The reference feeds every exact k-mer hit of a (ref read, orientation, hash read) pair through
//  Add_Match (overlapInCore-Find_Overlaps.C:26-96) in the order: ref offset ascending, and within
//  one ref offset hash offset descending (chain head = last inserted, Build_Hash_Index.C:288-294).
//  A maximal run of hits on one diagonal at consecutive ref offsets becomes ONE Match_Node
//  (Start = first ref offset, Offset = first hash offset, Len = K + hits - 1); what depends on the
//  arrival order is (a) the order of the nodes in the list (extension may move a node to the
//  front) and (b) the `consistent` flag.  We get the runs from the lookup kernel and replay
//  Add_Match's list dynamics here, but only at the offsets where something can change: a run is
//  born, a run dies, or the previous offset reordered the list.  If an offset produced no move, no
//  birth and no death, the following offsets up to the next birth/death are provably identical
//  no-ops (the list then starts with the active runs in processing order), so they are skipped.
//
//  runs[0..n): sorted by start ascending.  scratch: nxt, hits, act (n ints each).
//  out: order[0..n) = run indices in final list order (head first); returns `consistent`.
//  ---------------------------------------------------------------------------------------------
struct OvlRun { int32_t start, q, len; };      // ref start, hash start, number of k-mer hits

OVL_HD int ovl_chain_simulate(const OvlRun *runs, int n, int K, int32_t *nxt, int32_t *hits, int32_t *act, int32_t *order) {
  int consistent = 1;
  int head = -1;
  int nact = 0;
  int r = 0;
  int o = runs[0].start;

  while (true) {
    int nb = 0;
    while (r + nb < n && runs[r + nb].start == o) { hits[r + nb] = 0; act[nact + nb] = r + nb; nb++; }
    int tot = nact + nb;

    //  hits at offset o, hash offset descending
    for (int i = 1; i < tot; i++) {
      int x = act[i];
      int kx = runs[x].q + (o - runs[x].start);
      int j = i - 1;
      while (j >= 0) {
        int y = act[j];
        int ky = runs[y].q + (o - runs[y].start);
        if (ky >= kx) break;
        act[j + 1] = y; j--;
      }
      act[j + 1] = x;
    }

    bool changed = (nb > 0);

    for (int i = 0; i < tot; i++) {
      int x = act[i];
      if (hits[x] == 0) {
        //  Add_Match falls through its loop and creates a node (Find_Overlaps.C:61-86)
        int new_diag = runs[x].q - runs[x].start;
        int diag = 0, expected = 0, num_checked = 0;
        for (int p = head; p != -1; p = nxt[p]) {
          expected = runs[p].start + hits[p];
          diag     = runs[p].q - runs[p].start;
          if (expected < o) break;
          num_checked++;
        }
        int dd = diag - new_diag; if (dd < 0) dd = -dd;
        if (head != -1 && (num_checked > 0 || dd > 3 || o < expected + K - 2))
          consistent = 0;
        nxt[x] = head; head = x; hits[x] = 1;
      } else {
        //  Add_Match extends node x (Find_Overlaps.C:45-55), moving it to the front if a node
        //  with the same expected start but another diagonal was met on the way.
        bool mtf = false;
        int prev = -1, p = head;
        while (p != x) {
          if (runs[p].start + hits[p] == o) mtf = true;
          prev = p; p = nxt[p];
        }
        hits[x]++;
        if (mtf) {
          nxt[prev] = nxt[x];
          nxt[x] = head; head = x;
          changed = true;
        }
      }
    }

    //  drop runs that just received their last hit
    int na = 0;
    for (int i = 0; i < tot; i++) { int x = act[i]; if (hits[x] < runs[x].len) act[na++] = x; else changed = true; }
    nact = na;
    r += nb;

    if (nact == 0 && r == n) break;

    int next_birth = (r < n) ? runs[r].start : 0x7fffffff;
    if (nact == 0) { o = next_birth; continue; }

    if (changed) { o = o + 1; continue; }

    //  steady state: jump to the next birth or death
    int next_death = 0x7fffffff;
    for (int i = 0; i < nact; i++) { int x = act[i]; int dth = runs[x].start + runs[x].len; if (dth < next_death) next_death = dth; }
    int o2 = next_birth < next_death ? next_birth : next_death;
    int adv = o2 - 1 - o;                      // offsets o+1 .. o2-1 are skipped
    if (adv > 0)
      for (int i = 0; i < nact; i++) hits[act[i]] += adv;
    o = o2;
    na = 0;
    for (int i = 0; i < nact; i++) { int x = act[i]; if (hits[x] < runs[x].len) act[na++] = x; }
    nact = na;
    if (nact == 0 && r == n) break;
    if (nact == 0) o = next_birth;
  }

  int k = 0;
  for (int p = head; p != -1; p = nxt[p]) order[k++] = p;
  return consistent;
}

//  ---------------------------------------------------------------------------------------------
//  ovOverlap record packing (stores/ovOverlap.H:49-67) and Output_Overlap / Output_Partial_Overlap
//  (overlapInCore-Output.C:27-264).
//  ---------------------------------------------------------------------------------------------
struct OvlOlap {                       // Olap_Info_t without the delta array (only delta_ct is consumed)
  int s_lo, s_hi, t_lo, t_hi;
  double quality;
  int delta_ct;
  int s_left_boundary, s_right_boundary, t_left_boundary, t_right_boundary;
  int min_diag, max_diag;
};

OVL_HD void ovl_pack_record(uint32_t a, uint32_t b, uint32_t ahg5, uint32_t ahg3, uint32_t bhg5, uint32_t bhg3,
                            uint32_t span, uint32_t evalue, uint32_t flipped, uint32_t obt, uint32_t dup, uint32_t utg,
                            uint32_t *oa, uint32_t *ob, uint64_t *w0, uint64_t *w1) {
  const uint64_t M = (1ull << 21) - 1;
  *oa = a; *ob = b;
  *w0 = ((uint64_t)ahg5 & M) | (((uint64_t)ahg3 & M) << 21) | (((uint64_t)evalue & 0xffff) << 42) |
        ((uint64_t)(flipped & 1) << 58) | ((uint64_t)(obt & 1) << 59) | ((uint64_t)(dup & 1) << 60) | ((uint64_t)(utg & 1) << 61);
  *w1 = ((uint64_t)bhg5 & M) | (((uint64_t)bhg3 & M) << 21) | (((uint64_t)span & M) << 42);
}

//  evalue = AS_OVS_encodeEvalue(quality) (stores/ovOverlap.H:31-35).  No FMA: two roundings.
OVL_HD uint32_t ovl_encode_evalue(double q) {
#if defined(__CUDA_ARCH__)
  return (q < 65535 / 100000.0) ? (uint32_t)__dadd_rn(__dmul_rn(100000.0, q), 0.5) : 65535u;
#else
  volatile double t = 100000.0 * q;
  return (q < 65535 / 100000.0) ? (uint32_t)(t + 0.5) : 65535u;
#endif
}

//  Returns 1 if the overlap is "contained" (bhg <= 0) else 0 (dovetail), Output.C:175-178.
OVL_HD int ovl_output_overlap(uint32_t s_id, int s_len, int dir, uint32_t t_id, int t_len, const OvlOlap &o,
                              uint32_t *oa, uint32_t *ob, uint64_t *w0, uint64_t *w1) {
  uint32_t span = (uint32_t)(((o.s_hi - o.s_lo) + (o.t_hi - o.t_lo) + o.delta_ct) / 2);
  int s_right_hang = s_len - o.s_hi - 1;
  int t_right_hang = t_len - o.t_hi - 1;
  bool sleft = (o.s_lo > o.t_lo) || (o.s_lo == o.t_lo && s_right_hang > t_right_hang);
  uint32_t a = sleft ? s_id : t_id, b = sleft ? t_id : s_id;
  int orient, ahg, bhg;                                  // orient: 0 'N', 1 'I', 2 'O'
  if (sleft) { orient = (dir == 0) ? 0 : 2; ahg = o.s_lo; bhg = t_right_hang - s_right_hang; }
  else       { orient = (dir == 0) ? 0 : 1; ahg = o.t_lo; bhg = s_right_hang - t_right_hang; }
  if (orient == 2 && s_right_hang >= t_right_hang) {
    orient = 1;
    ahg = -(t_right_hang - s_right_hang);
    bhg = -(o.s_lo);
  }
  int ah = (orient == 2) ? -bhg : ahg;
  int bh = (orient == 2) ? -ahg : bhg;
  uint32_t ahg5 = (ah < 0) ? 0 : (uint32_t)ah, bhg5 = (ah < 0) ? (uint32_t)-ah : 0;      // ovOverlap::a_hang(int32)
  uint32_t bhg3 = (bh < 0) ? 0 : (uint32_t)bh, ahg3 = (bh < 0) ? (uint32_t)-bh : 0;      // ovOverlap::b_hang(int32)
  ovl_pack_record(a, b, ahg5, ahg3, bhg5, bhg3, span, ovl_encode_evalue(o.quality), orient != 0, 0, 0, 1, oa, ob, w0, w1);
  return bhg <= 0;
}

OVL_HD void ovl_output_partial(uint32_t s_id, uint32_t t_id, int dir, const OvlOlap &o, int s_len, int t_len,
                               uint32_t *oa, uint32_t *ob, uint64_t *w0, uint64_t *w1) {
  uint32_t span = (uint32_t)(((o.s_hi - o.s_lo) + (o.t_hi - o.t_lo) + o.delta_ct) / 2);
  uint32_t ahg5, ahg3, bhg5, bhg3;
  if (dir == 0) {
    ahg5 = (uint32_t)o.s_lo;                 ahg3 = (uint32_t)(s_len - (o.s_hi + 1));
    bhg5 = (uint32_t)o.t_lo;                 bhg3 = (uint32_t)(t_len - (o.t_hi + 1));
  } else {
    ahg5 = (uint32_t)(s_len - (o.s_hi + 1)); ahg3 = (uint32_t)o.s_lo;
    bhg5 = (uint32_t)(t_len - (o.t_hi + 1)); bhg3 = (uint32_t)o.t_lo;
  }
  ovl_pack_record(s_id, t_id, ahg5, ahg3, bhg5, bhg3, span, ovl_encode_evalue(o.quality), dir != 0, 1, 1, 0, oa, ob, w0, w1);
}

//  ---------------------------------------------------------------------------------------------
//  Store-ingest step (ovl_ingest.cu): the mirrored twin of an overlap and the error-rate filter.
//  ovOverlap::swapIDs (stores/ovOverlap.C:215-246) + ovStoreFilter::filterOverlap (stores/ovStoreFilter.C:71-150).
//  In: one record (dat0, dat1).  Out: dat0 of the forward record after the filter, dat0/dat1 of the twin (whose IDs are
//  the forward record's, swapped); returns 1 if the pair still carries a forUTG/forOBT/forDUP flag (is kept).
//  Host+device so that tests/model can run the very code the kernel runs.
//  ---------------------------------------------------------------------------------------------
#define OVL_ING_M21    ((1ull << 21) - 1)
#define OVL_ING_FLAGS  (7ull << 59)                    // forOBT (59) | forDUP (60) | forUTG (61)

OVL_HD int ovl_ingest_twin(uint64_t w0, uint64_t w1, uint32_t max_evalue, uint64_t *f0, uint64_t *r0, uint64_t *r1) {
  const uint64_t ahg5 = w0 & OVL_ING_M21, ahg3 = (w0 >> 21) & OVL_ING_M21, bhg5 = w1 & OVL_ING_M21, bhg3 = (w1 >> 21) & OVL_ING_M21;
  const bool flipped = ((w0 >> 58) & 1ull) != 0;
  const uint32_t evalue = (uint32_t)(w0 >> 42) & 0xFFFFu;
  if (evalue > max_evalue) w0 &= ~OVL_ING_FLAGS;                       // both twins lose their flags
  const uint64_t hi0 = w0 & ~((1ull << 42) - 1), hi1 = w1 & ~((1ull << 42) - 1);
  *f0 = w0;
  *r0 = hi0 | (flipped ? bhg3 : bhg5) | ((flipped ? bhg5 : bhg3) << 21);
  *r1 = hi1 | (flipped ? ahg3 : ahg5) | ((flipped ? ahg5 : ahg3) << 21);
  return (w0 & OVL_ING_FLAGS) != 0;
}
