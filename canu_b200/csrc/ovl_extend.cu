//  ovl_extend.cu -- K4/K5: warp-per-candidate-pair banded prefix-edit-distance extension.
//
//  Replaces, bit-exactly (paths under /root/reference/src/overlapInCore):
//    Process_Matches                     overlapInCore-Process_String_Overlaps.C:355-547
//    Extend_Alignment                    liboverlap/prefixEditDistance-extend.C:36-183
//    prefixEditDistance::forward/reverse liboverlap/prefixEditDistance-forward.C:94-314, -reverse.C:114-330
//    Set_Right_Delta / Set_Left_Delta    -forward.C:32-71, -reverse.C:36-93
//    Add_Overlap, Combine_Into_One_Olap, Merge_Intersecting_Olaps, Choose_Best_Partial,
//    Lies_On_Alignment                   Process_String_Overlaps.C:42-311
//    Output_Overlap / Output_Partial_Overlap   overlapInCore-Output.C:27-264
//
//  Design.  One warp owns one oriented read pair.  The O(ND) edit array is computed one error row
//  at a time; lanes stride the diagonals of the row.  Only the previous and the current row are
//  kept (shared-memory rings indexed by diagonal modulo the ring size; a per-warp ring in HBM takes
//  over for bands wider than the shared ring).  For the traceback the reference keeps the whole
//  triangle of int32 cells; we keep 2 bits per cell ("from" code: 0 = same diagonal/mismatch,
//  1 = from d-1, 2 = from d+1, decided with the reference's tie order) as two ballot bit-planes per
//  32 cells in a per-warp HBM arena, walk the path back through the codes, then re-slide forward
//  along the path to recover the row values the delta encoding needs.  The reverse extension runs
//  the same forward code on the reverse-complemented copies of both reads.
//
//  Integer DP: no tensor cores.  The only floating point is the branch-point score, evaluated once
//  per row in FP64 with separate multiply and subtract (the reference build has no FMA).
#include "ovl_ctx.h"
#include <cstdlib>

#define EXT_WARPS  8
#define EXT_THREADS (EXT_WARPS * 32)
#define SRING      256                    // ints per shared ring (per row, per warp)
#define SRING_PAD  64                     // slack behind every ring: the branch-free cell front reads predecessors of masked lanes
#define SRING_STRIDE (SRING + SRING_PAD)
#define FULL       0xffffffffu

struct DpOut { int a_end, t_end, errors, leftover, match_to_end, delta_len; };

//  Everything a warp needs that is the same in all 32 lanes lives ONCE per warp in shared memory: scratch pointers, the
//  result of the last extension, the overlaps collected for the current pair, statistics.  Kept in per-thread variables
//  these were spilled to local memory (696 B of stack per thread = 530 KB per SM at 24 warps), which evicted the read
//  data from L1 and put an L2 round trip behind every access (ncu: 42 % long-scoreboard stalls, most of them after LDL).
struct WarpCtl {
  int      *gring0, *gring1;              // HBM rings (bands wider than the shared ring)
  uint2    *arena;                        // from-code bit planes
  int2     *row_meta;                     // per row: (leftmost diagonal, first arena word)
  uint8_t  *path; int32_t *ival; uint32_t *ikc;
  int32_t  *ldelta, *rdelta;
  uint64_t  arena_cap;
  uint32_t  gring_cap;
  int       emax;
  DpOut     o;                            // result of the last warp_dp
  const uint64_t *s_fwd, *s_rc, *t_fwd, *t_rc;   // current pair: ref read in its search orientation (and its reverse complement), hash read
  int       s_len, t_len;
  uint32_t  s_id, t_id;
  int64_t   seed_begin;
  int       n_seeds, dir, consistent;
  int       distinct_ct;
  OvlOlap   distinct[OVL_MAX_DISTINCT_OLAPS];
  unsigned long long cells, calls, c_with, c_without, c_multi, c_total, c_cont, c_dove;
};

//  dynamic shared memory of the extension kernels: | DevParams | WarpCtl x EXT_WARPS | rings: EXT_WARPS x 2 x SRING ints |
#define EXT_SM_PARAMS 0
#define EXT_SM_CTL    128
#define EXT_SM_RINGS  (EXT_SM_CTL + ((EXT_WARPS * (int)sizeof(WarpCtl) + 127) / 128) * 128)
#define EXT_SM_BYTES  (EXT_SM_RINGS + EXT_WARPS * 2 * SRING_STRIDE * 4)
static_assert(sizeof(DevParams) <= EXT_SM_CTL, "DevParams must fit its shared-memory slot");
extern __shared__ __align__(16) unsigned char ext_sm[];

__device__ __forceinline__ const DevParams &sh_params() { return *reinterpret_cast<const DevParams *>(ext_sm + EXT_SM_PARAMS); }
__device__ __forceinline__ WarpCtl &sh_ctl(int wib) { return reinterpret_cast<WarpCtl *>(ext_sm + EXT_SM_CTL)[wib]; }
__device__ __forceinline__ WarpCtl &sh_ctl() { return sh_ctl(threadIdx.x >> 5); }
__device__ __forceinline__ int *sh_ring(int which, int wib) { return reinterpret_cast<int *>(ext_sm + EXT_SM_RINGS) + (wib * 2 + which) * SRING_STRIDE; }

//  Number of leading positions (< lim) where A[a..] and T[t..] match; all 32 lanes cooperate, 512 bases per round.
__device__ __forceinline__ int warp_slide(const uint64_t *A, int a, const uint64_t *T, int t, int lim, int lane) {
  int total = 0;
  while (total < lim) {
    int pos = total + 16 * lane;
    int k = 0;
    if (pos < lim) k = ovl_match16(ovl_fetch16(A, a + pos), ovl_fetch16(T, t + pos));
    unsigned nf = __ballot_sync(FULL, k < 16);
    if (nf == 0) { total += 512; continue; }
    int first = __ffs(nf) - 1;
    int kk = __shfl_sync(FULL, k, first);
    total += 16 * first + kk;
    break;
  }
  return total < lim ? total : lim;
}

//  Pull the read lines the DP is about to walk over into L1 (mode 1) or L2 (mode 2): lanes 0..15 take 16 consecutive
//  128-byte lines (256 bases each) of A starting at the line of base `a`, lanes 16..31 the same for T.  On HiFi-like
//  reads an extension is a chain of dependent row -> slide -> row steps, each waiting a full DRAM round trip for bases
//  that lie a few hundred bytes further down the same two reads; one burst of prefetches per ~4 kb turns all but the
//  first of those waits into L1 hits.  Lines past the end of the read (`a_end`, `t_end`: exclusive base limits) are skipped.
#define PF_SPAN 3840                      // bases certainly covered by one burst (15 whole lines)
__device__ __forceinline__ void dp_prefetch(int mode, const uint64_t *A, int a, int a_end, const uint64_t *T, int t, int t_end, int lane) {
  const bool is_t = lane >= 16;
  const uintptr_t base = (uintptr_t)(is_t ? T : A);
  int from = is_t ? t : a; if (from < 0) from = 0;
  const int end = is_t ? t_end : a_end;
  const uintptr_t p = ((base + ((uintptr_t)from >> 1)) & ~(uintptr_t)127) + 128u * (uint32_t)(lane & 15);
  if (p < base + (((uintptr_t)end + 1) >> 1)) {
    if (mode == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
    else           asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
  }
}

//  i-th value pushed by Set_Right_Delta / Set_Left_Delta's loop (k descending): V_i = row value below indel i.
__device__ __forceinline__ int push_at(const int32_t *ival, const uint32_t *ikc, int i, int v_start) {
  int vi = ival[i];
  int vp = (i == 0) ? v_start : ival[i - 1];
  return ((ikc[i] >> 30) == 1u) ? (vi - vp - 1) : (vp - vi);
}

//  Traceback through the from-codes, then delta encoding.  Returns delta_len; writes deltas to `out`.
//  fwd_rules: Set_Right_Delta conventions; else Set_Left_Delta (sets leftover, may bump t_mag).
__device__ __noinline__ int warp_traceback(const uint64_t *A, int a0, int m, const uint64_t *T, int t0, int n,
                              int e_start, int d_start, int v_start, int row0, bool fwd_rules, int first_code,
                              int32_t *out, int lane) {
  WarpCtl &C = sh_ctl();
  const int2 *row_meta = C.row_meta; const uint2 *arena = C.arena;
  uint8_t *path = C.path; int32_t *ival = C.ival; uint32_t *ikc = C.ikc;
  //  Phase A: walk down, 32 rows per round
  int n_ind = 0;
  int dcur = d_start;
  for (int kb = e_start; kb >= 1; kb -= 32) {
    const int kk = kb - lane;                      // my row
    int rl = 0; uint32_t glo = 0; uint2 w0 = make_uint2(0, 0), w1 = w0, w2 = w0;
    if (kk >= 1) {
      const int2 rm = row_meta[kk];
      rl = rm.x;
      uint32_t ro = (uint32_t)rm.y;
      int idx_lo = dcur - lane - rl; if (idx_lo < 0) idx_lo = 0;
      glo = (uint32_t)idx_lo >> 5;
      const uint2 *ap = arena + ro + glo;
      w0 = ap[0]; w1 = ap[1]; w2 = ap[2];
    }
    int mycode = 0;
    const int steps = kb < 32 ? kb : 32;
    for (int l = 0; l < steps; l++) {
      int rl_l = __shfl_sync(FULL, rl, l);
      uint32_t glo_l = __shfl_sync(FULL, glo, l);
      int idx = dcur - rl_l;
      int gi = (idx >> 5) - (int)glo_l;
      unsigned x0 = gi == 0 ? w0.x : (gi == 1 ? w1.x : w2.x);
      unsigned y0 = gi == 0 ? w0.y : (gi == 1 ? w1.y : w2.y);
      unsigned x = __shfl_sync(FULL, x0, l), y = __shfl_sync(FULL, y0, l);
      //  note: gi computed from the broadcast (uniform) values, so every lane selects the same word index
      int bit = idx & 31;
      int code = ((x >> bit) & 1) | (((y >> bit) & 1) << 1);
      if (first_code >= 0 && kb == e_start && l == 0) code = first_code;   // cell the DP never evaluated (forced mismatch)
      if (lane == l) mycode = code;
      if (code == 1) dcur--; else if (code == 2) dcur++;
    }
    if (kk >= 1) path[kk] = (uint8_t)mycode;
    unsigned im = __ballot_sync(FULL, kk >= 1 && mycode != 0);
    if (kk >= 1 && mycode != 0) {
      int idx = n_ind + __popc(im & ((1u << lane) - 1));
      ikc[idx] = (uint32_t)kk | ((uint32_t)mycode << 30);
    }
    n_ind += __popc(im);
  }
  __syncwarp();

  //  Phase B: re-slide upward to recover v_{k-1} at every indel step k
  if (n_ind > 0) {
    int v = row0, d = 0, seen = 0;
    for (int kb = 1; kb <= e_start; kb += 32) {
      int myc = 0;
      if (kb + lane <= e_start) myc = path[kb + lane];
      const int steps = (e_start - kb + 1) < 32 ? (e_start - kb + 1) : 32;
      for (int l = 0; l < steps; l++) {
        const int k = kb + l;
        const int c = __shfl_sync(FULL, myc, l);
        if (c != 0) {
          if (lane == 0) ival[n_ind - 1 - seen] = v;
          seen++;
        }
        if (k == e_start || seen == n_ind) { kb = e_start + 1; break; }   // nothing above the last indel is needed
        d += (c == 1) ? 1 : ((c == 2) ? -1 : 0);
        int pre = (c == 1) ? v : v + 1;
        int lim = min(m - pre, n - d - pre);
        v = pre + warp_slide(A, a0 + pre, T, t0 + pre + d, lim, lane);
      }
    }
  }
  __syncwarp();

  //  Phase C: deltas.  push_i = (code 1) ? V_i - V_{i-1} - 1 : V_{i-1} - V_i, with V_{-1} = v_start.
  const int last = n_ind ? ival[n_ind - 1] : v_start;
#define PUSH(i) push_at(ival, ikc, (i), v_start)
  int len = n_ind;
  if (fwd_rules) {
    //  stack S[0..n_ind-1] = pushes, S[n_ind] = last+1;  Right_Delta[j] = |S[n_ind-j]| * sign(S[n_ind-j-1])
    for (int j = lane; j < n_ind; j += 32) {
      int i = n_ind - j;
      int si  = (i == n_ind) ? (last + 1) : PUSH(i);
      int sim = PUSH(i - 1);
      int a = si < 0 ? -si : si;
      out[j] = a * ((sim > 0) - (sim < 0));
    }
  } else {
    int t_mag = C.o.t_end;
    bool fix = false;
    if (n_ind > 1) {
      int p0 = PUSH(0);
      fix = (p0 == 1) && (t_mag < n);                 // Left_Delta[0] == 1 && t_end + t_len > 0
    }
    if (fix) {
      for (int j = lane; j < n_ind - 1; j += 32) {
        int pj = PUSH(j + 1);
        out[j] = (j == 0) ? ((pj > 0) ? pj + 1 : pj - 1) : pj;
      }
      len = n_ind - 1;
      t_mag += 1;
    } else {
      for (int j = lane; j < n_ind; j += 32) out[j] = PUSH(j);
    }
    __syncwarp();
    if (lane == 0) { C.o.leftover = last; C.o.t_end = t_mag; }
  }
#undef PUSH
  __syncwarp();
  return len;
}

//  The cells of one error row: 32 diagonals per round.  SH = the two row rings live in shared memory (bands up to
//  SRING-4 wide), addressed from the kernel's shared-memory symbol so that the compiler emits LDS/STS; otherwise
//  they are the warp's HBM rings.  `psel` says which of the two rings holds row e-1.
//  Ring layout: a row computed for diagonals [Lu, Ru] is stored at index d - Lu + 2 (no modular wrap: a row is
//  at most ring size - 4 wide), so cell d of the next row finds EA[e-1][d-1], [d], [d+1] in three consecutive
//  words at index d - pbase + 1, `pbase` being the Lu of the row below; `pidx` = Lu - pbase + 1.
//  Band pruning and best-cell bookkeeping are folded into the cell loop (forward.C:245-299): a cell survives the
//  pruning iff EA[e][d] + max(d,0) >= Edit_Match_Limit[e]; the new band is [min, max] surviving d, and the longest
//  cell of that band is always a surviving one (a pruned cell inside the band is shorter than the nearest survivor
//  on its left), so one pass over the row gives Left, Right, Longest and Best_d.
struct RowOut { int mn, mx, bv, bd, term_d, term_row; };

//  Read bases are only ever read by this kernel: fetch them through the non-coherent path (LDG.E.CONSTANT) -- the
//  pointers come out of the per-warp control block in shared memory, so the compiler would otherwise emit generic LD.
__device__ __forceinline__ uint32_t fetch8_nc(const uint32_t *w32, int x) {
  const uint32_t *p = w32 + (x >> 3);
  return __funnelshift_r(__ldg(p), __ldg(p + 1), (x & 7) << 2);
}

//  One cell, front half: predecessor max, slide limit, first 8 bases.
struct CellState { int row, code, lim, cnt; bool act, more; };

//  Branch-free on purpose: with divergent `if (act)` / `if (lim > 0)` regions the compiler serialises the fronts of the
//  two groups of an iteration and every load is waited for before the next group starts.  Inactive lanes (past the
//  right edge of the band: only in a row's last group) compute on row 0 / diagonal 0 and are masked; the rings are
//  padded so that their three predecessor loads stay in bounds.  lim >= 0 for every active cell (a cell at the end of
//  A or T ends the extension in its own row), so cnt = min(cnt, lim) also covers lim == 0.
__device__ __forceinline__ void cell_front(CellState &c, const int *prev, int d, int Ru,
                                           const uint32_t *A32, int a0, int m, const uint32_t *T32, int t0, int n) {
  const bool act = d <= Ru;
  const int a = prev[0], b = prev[1], c2 = prev[2];
  int row = 1 + b, code = 0;
  if (a > row) { row = a; code = 1; }
  if (1 + c2 > row) { row = 1 + c2; code = 2; }
  row = act ? row : 0;
  const int dd = act ? d : 0;
  int lim = min(m - row, n - dd - row);
  lim = act ? lim : 0;
  //  every lane slides its own diagonal over the first 8 bases (32-bit arithmetic) ...
  const int c8 = ovl_match8(fetch8_nc(A32, a0 + row), fetch8_nc(T32, t0 + row + dd));
  c.act = act; c.row = row; c.code = code; c.lim = lim;
  c.more = (c8 == 8) && (lim > 8);
  c.cnt = min(c8, lim);
}

//  ... and the few diagonals that are still matching (on real overlaps: the true one) are finished by the whole
//  warp, 512 bases per round, instead of one lane chasing dependent loads
__device__ __forceinline__ void cell_slides(CellState &c, int d, const uint64_t *A, int a0, const uint64_t *T, int t0, int lane) {
  for (unsigned pend = __ballot_sync(FULL, c.more); pend; pend &= pend - 1) {
    const int l = __ffs(pend) - 1;
    const int r_l = __shfl_sync(FULL, c.row, l), d_l = __shfl_sync(FULL, d, l), lim_l = __shfl_sync(FULL, c.lim, l);
    const int ext = warp_slide(A, a0 + r_l + 8, T, t0 + r_l + d_l + 8, lim_l - 8, lane);
    if (lane == l) c.cnt = 8 + ext;
  }
}

//  Back half: store, pruning / best bookkeeping, from-code planes, termination.  Returns true if a cell reached the
//  end of A or T (cnt == lim <=> row == m or row + d == n, because lim = min(m - row, n - d - row) >= 0).
__device__ __forceinline__ bool cell_back(const CellState &c, int *cur, int d, int dbase, int lim_e, uint2 *arena_slot, int lane,
                                          int &mn, int &mx, int &bv, int &bd, RowOut &ro) {
  const int row = c.row + c.cnt;
  if (c.act) {
    cur[0] = row;
    if (!(row + (d > 0 ? d : 0) < lim_e)) {
      mn = min(mn, d); mx = max(mx, d);
      if (row > bv) { bv = row; bd = d; }
    }
  }
  const unsigned b0 = __ballot_sync(FULL, c.act && (c.code & 1));
  const unsigned b1 = __ballot_sync(FULL, c.act && (c.code >> 1));
  if (lane == 0) *arena_slot = make_uint2(b0, b1);
  const unsigned hb = __ballot_sync(FULL, c.act && (c.cnt == c.lim));
  if (hb) {
    const int tl = __ffs(hb) - 1;
    ro.term_d = dbase + tl;
    ro.term_row = __shfl_sync(FULL, row, tl);
    return true;
  }
  return false;
}

//  U = groups of 32 cells in flight per iteration: with U = 2 the loads and the dependent max/shift/compare chains of
//  two groups overlap (ncu: the single-group loop left 24 % of the issue slots empty on long-scoreboard / wait stalls
//  behind the base fetches at 30 warps per SM).
template <bool SH, int U>
__device__ __forceinline__ void dp_row_cells(int psel, int pidx, const uint64_t *A, int a0, int m, const uint64_t *T, int t0, int n,
                                             int Lu, int Ru, uint32_t ngroups, int lim_e, uint2 *arena_row, int lane, int wib, RowOut &ro) {
  const int *prev; int *cur;
  if (SH) { int *ring = sh_ring(0, wib); prev = ring + psel * SRING_STRIDE; cur = ring + (psel ^ 1) * SRING_STRIDE; }
  else    { WarpCtl &C = sh_ctl(wib); prev = psel ? C.gring1 : C.gring0; cur = psel ? C.gring0 : C.gring1; }
  const uint32_t *A32 = reinterpret_cast<const uint32_t *>(A), *T32 = reinterpret_cast<const uint32_t *>(T);
  int mn = 0x7fffffff, mx = -0x7fffffff, bv = -1, bd = 0x7fffffff;
  ro.term_d = 0x7fffffff; ro.term_row = 0;
  prev += pidx + lane; cur += 2 + lane;
  uint32_t g = 0;
  if (U == 2) {
    for (; g + 2 <= ngroups; g += 2, prev += 64, cur += 64) {
      const int dbase = Lu + (int)(g << 5);
      const int d0 = dbase + lane, d1 = d0 + 32;
      CellState c0, c1;
      cell_front(c0, prev, d0, Ru, A32, a0, m, T32, t0, n);
      cell_front(c1, prev + 32, d1, Ru, A32, a0, m, T32, t0, n);
      cell_slides(c0, d0, A, a0, T, t0, lane);
      cell_slides(c1, d1, A, a0, T, t0, lane);
      if (cell_back(c0, cur, d0, dbase, lim_e, arena_row + g, lane, mn, mx, bv, bd, ro)) goto done;
      if (cell_back(c1, cur + 32, d1, dbase + 32, lim_e, arena_row + g + 1, lane, mn, mx, bv, bd, ro)) goto done;
    }
  }
  for (; g < ngroups; g++, prev += 32, cur += 32) {
    const int dbase = Lu + (int)(g << 5);
    const int d = dbase + lane;
    CellState c;
    cell_front(c, prev, d, Ru, A32, a0, m, T32, t0, n);
    cell_slides(c, d, A, a0, T, t0, lane);
    if (cell_back(c, cur, d, dbase, lim_e, arena_row + g, lane, mn, mx, bv, bd, ro)) break;
  }
done:
  ro.mn = mn; ro.mx = mx; ro.bv = bv; ro.bd = bd;
}

#ifdef OVL_DP_NOINLINE
#define OVL_DP_LINKAGE __noinline__
#else
#define OVL_DP_LINKAGE __forceinline__
#endif

//  One banded extension (forward(), or reverse() on reverse-complemented strings).
//  A: shorter string (m <= n), starting at base a0 of the dp4 words A; T likewise.
//  It has exactly ONE call site (the side loop of warp_extend_alignment), so inlining it does not duplicate it:
//  an earlier version called it from four places and, fully inlined, was 17.6 k SASS instructions that spent 79 % of
//  their stall samples waiting for instruction fetch; as a separate function every memory access re-materialised
//  its descriptor (R2UR) from the call ABI's vector registers.
template <int ILP>
__device__ OVL_DP_LINKAGE void warp_dp(const uint64_t *A, int a0, int m, const uint64_t *T, int t0, int n,
                        int error_limit, bool fwd_rules, unsigned long long *err_flags, int lane, int wib) {
  const DevParams &P = sh_params();
  WarpCtl &C = sh_ctl(wib);
  DpOut &o = C.o;                                    // every lane stores the same values
  unsigned int cells = 0;
  __syncwarp();                                      // every lane has read the previous extension's result
  o.leftover = 0; o.delta_len = 0;
  int pf_next = PF_SPAN;
  if (P.ext_prefetch) dp_prefetch(P.ext_prefetch, A, a0, a0 + m, T, t0, t0 + n, lane);
  int row0 = warp_slide(A, a0, T, t0, m, lane);
  if (row0 == m) {                                   // exact match to the end of A
    o.a_end = m; o.t_end = m; o.leftover = m; o.match_to_end = 1; o.errors = 0;
    return;
  }
  if (lane == 0) C.calls++;

  int psel = 0;                                      // ring holding row e-1
  int pbase = 0;                                     // diagonal stored at index 2 of that ring
  uint32_t rcap = SRING;
  bool in_shared = true;
  if (lane == 0) sh_ring(0, wib)[2] = row0;
  int L = 0, R = 0;
  int longest = 0, best_d = 0, best_e = 0;
  int ms_len = 0, ms_d = 0, ms_e = 0;
  double max_score = 0.0;
  uint32_t aoff = 0;
  int e;
  bool reached_end = false;
  int tb_e = 0, tb_d = 0, tb_v = 0, first_code = -1;
  __syncwarp();

  for (e = 1; e <= error_limit; e++) {
    const int Lu = L - 1, Ru = R + 1;
    const int width = Ru - Lu + 1;
    if (in_shared && width + 4 > SRING) {            // migrate the previous row to the HBM ring
      int *g0 = psel ? C.gring1 : C.gring0;
      const int *sp = sh_ring(psel, wib);
      for (int i = lane; i <= R - pbase + 4; i += 32) g0[i] = sp[i];
      __syncwarp();
      rcap = C.gring_cap; in_shared = false;
    }
    const uint32_t ngroups = (uint32_t)(width + 31) >> 5;
    if ((!in_shared && (uint32_t)(width + 4) > rcap) || (uint64_t)aoff + ngroups + 4 > C.arena_cap || e > C.emax) {
      if (lane == 0) atomicOr(err_flags, 4ull);       // scratch too small: reported as an error by the host
      break;
    }
    //  sentinels at L-1, L-2, R+1, R+2: one store.  Shared and global rings are addressed in separate branches: a pointer
    //  that may be either is a generic pointer, and forming one from a shared address costs special-register reads per row
    if (lane < 4) {
      const int si = (lane < 2 ? L - 1 - lane : R - 1 + lane) + 2 - pbase;
      if (in_shared) sh_ring(psel, wib)[si] = -2; else (psel ? C.gring1 : C.gring0)[si] = -2;
    }
    if (lane == 4) C.row_meta[e] = make_int2(Lu, (int)aoff);
    __syncwarp();

    RowOut ro;
    if (in_shared) dp_row_cells<true, ILP>(psel, Lu - pbase + 1, A, a0, m, T, t0, n, Lu, Ru, ngroups, P.eml[e], C.arena + aoff, lane, wib, ro);
    else           dp_row_cells<false, 1>(psel, Lu - pbase + 1, A, a0, m, T, t0, n, Lu, Ru, ngroups, P.eml[e], C.arena + aoff, lane, wib, ro);
    const int term_d = ro.term_d, term_row = ro.term_row;
    int mn = ro.mn, mx = ro.mx; const int bv = ro.bv, bd = ro.bd;
    __syncwarp();

    if (term_d != 0x7fffffff) {
      cells += (unsigned int)(term_d - Lu + 1);
      //  reached the end of A or T: branch-point test (forward.C:170-212)
      double score = __dsub_rn(__dmul_rn((double)term_row, P.bmv), (double)e);
      int tail_len = term_row - ms_len;
      bool abort_ = false;
      if (P.partial && score < max_score) abort_ = true;
      if (e > OVL_MIN_BRANCH_END_DIST / 2 && tail_len >= OVL_MIN_BRANCH_END_DIST) {
        double slope = __ddiv_rn(__dsub_rn(max_score, score), (double)tail_len);
        if (slope >= P.min_tail_slope) abort_ = true;
      }
      if (abort_) break;                                // best-so-far result, assembled after the loop
      int d = term_d;
      //  generic pointer (rare accesses only), biased so that prev[d] = EA[e-1][d]
      const int *prev = (in_shared ? sh_ring(psel, wib) : (psel ? C.gring1 : C.gring0)) + (2 - pbase);
      if (fwd_rules && term_row == m && d < Ru && 1 + prev[d + 1] == term_row) {
        //  Force the last error to be a mismatch (forward.C:215-221): the path now starts in cell (e, d+1), which
        //  the DP never evaluated (it may lie in a 32-cell group after the one that terminated), so its from-code
        //  is derived here from row e-1 exactly as Set_Right_Delta does (forward.C:45-55).
        d++;
        const int pa = prev[d - 1], pb = prev[d], pc = prev[d + 1];
        int mx = 1 + pb; first_code = 0;
        if (pa > mx) { mx = pa; first_code = 1; }
        if (1 + pc > mx) first_code = 2;
      }
      o.a_end = term_row; o.t_end = term_row + d; o.match_to_end = 1; o.errors = e;
      tb_e = e; tb_d = d; tb_v = term_row;
      reached_end = true;
      break;
    }
    cells += (unsigned int)width;
    aoff += ngroups;

    mn = __reduce_min_sync(FULL, mn);
    mx = __reduce_max_sync(FULL, mx);
    if (mn > mx) break;                               // Left > Right
    L = mn; R = mx;
    int vmax = __reduce_max_sync(FULL, bv);
    int dmin = __reduce_min_sync(FULL, bv == vmax ? bd : 0x7fffffff);
    if (vmax > longest) { longest = vmax; best_d = dmin; best_e = e; }
    if (P.ext_prefetch && longest + 1536 > pf_next) {
      dp_prefetch(P.ext_prefetch, A, a0 + pf_next, a0 + m, T, t0 + pf_next + best_d, t0 + n, lane);
      pf_next += PF_SPAN;
    }

    double score = __dsub_rn(__dmul_rn((double)longest, P.bmv), (double)e);
    if (score > max_score) { max_score = score; ms_len = longest; ms_d = best_d; ms_e = best_e; }

    psel ^= 1; pbase = Lu;
    __syncwarp();
  }

  if (!reached_end) {
    //  branch point, error limit exhausted or band closed: the best-scoring prefix so far
    o.a_end = ms_len; o.t_end = ms_len + ms_d; o.match_to_end = 0; o.errors = ms_e;
    tb_e = ms_e; tb_d = ms_d; tb_v = (ms_e == 0) ? row0 : ms_len;
  }
  if (lane == 0) C.cells += cells;
  __syncwarp();
  const int dl = warp_traceback(A, a0, m, T, t0, n, tb_e, tb_d, tb_v, row0, fwd_rules, first_code, fwd_rules ? C.rdelta : C.ldelta, lane);
  o.delta_len = dl;
}

//  Extend_Alignment (prefixEditDistance-extend.C:36-183).  S is the ref read in its search orientation
//  (S.fwd = oriented sequence, S.rc = its reverse complement); T the hash read.
//  On return ldelta[0..ldelta_len) (the warp's ldelta scratch) is the merged Left_Delta.
template <int ILP>
__device__ int warp_extend_alignment(int m_start, int m_offset, int m_len,
                                     int &s_lo, int &s_hi, int &t_lo, int &t_hi, int &errors, int &ldelta_len,
                                     unsigned long long *err_flags, int lane, int wib) {
  const DevParams &P = sh_params();
  WarpCtl &C = sh_ctl(wib);
  int right_errors = 0, left_errors = 0, leftover = 0;
  bool r_to_end = true, l_to_end = true;
  const int S_len = C.s_len, T_len = C.t_len;
  const int s_left_begin = m_start - 1, s_right_begin = m_start + m_len, s_right_len = S_len - s_right_begin;
  const int t_left_begin = m_offset - 1, t_right_begin = m_offset + m_len, t_right_len = T_len - t_right_begin;
  const int total_olap = min(m_start, m_offset) + m_len + min(s_right_len, t_right_len);
  const int error_limit = (int)ceil(__dmul_rn((double)total_olap, P.erate));      // Error_Bound[Total_Olap]

  int rlen = 0, llen = 0;
  bool r_negate = false, l_negate = false;
  s_hi = 0; t_hi = 0; s_lo = 0; t_lo = 0;

  //  side 0: forward(S + s_right_begin, T + t_right_begin); side 1: reverse(S + s_left_begin, T + t_left_begin), which is
  //  the same forward code on the reverse complements starting at the mirrored position.  In both the shorter string plays
  //  "A" (extend.C:86-121,129-162).  One loop, hence ONE call site of warp_dp: it is inlined without being duplicated.
  #pragma unroll 1
  for (int side = 0; side < 2; side++) {
    const bool right = (side == 0);
    const bool skip = right ? (s_right_len == 0 || t_right_len == 0) : (s_left_begin < 0 || t_left_begin < 0);
    if (skip) continue;
    const bool swap = right ? !(s_right_len <= t_right_len) : !(s_right_begin <= t_right_begin);     // T plays "A"
    const uint64_t *sw = right ? C.s_fwd : C.s_rc, *tw = right ? C.t_fwd : C.t_rc;
    const int s0 = right ? s_right_begin : S_len - 1 - s_left_begin, sl = right ? s_right_len : s_left_begin + 1;
    const int t0 = right ? t_right_begin : T_len - 1 - t_left_begin, tl = right ? t_right_len : t_left_begin + 1;
    warp_dp<ILP>(swap ? tw : sw, swap ? t0 : s0, swap ? tl : sl, swap ? sw : tw, swap ? s0 : t0, swap ? sl : tl,
            right ? error_limit : error_limit - right_errors, right, err_flags, lane, wib);
    const DpOut o = C.o;
    const int s_end = swap ? o.t_end : o.a_end, t_end = swap ? o.a_end : o.t_end;
    if (right) { right_errors = o.errors; s_hi = s_end; t_hi = t_end; r_to_end = o.match_to_end; rlen = o.delta_len; r_negate = !swap; }
    else       { left_errors = o.errors; s_lo = -s_end; t_lo = -t_end; l_to_end = o.match_to_end; llen = o.delta_len; leftover = o.leftover; l_negate = swap; }
  }
  s_hi += s_right_begin - 1;
  t_hi += t_right_begin - 1;
  s_lo += s_left_begin + 1;
  t_lo += t_left_begin + 1;
  errors = left_errors + right_errors;

  const int kind = !r_to_end ? (!l_to_end ? OVL_NONE : OVL_RIGHT_BRANCH_PT) : (!l_to_end ? OVL_LEFT_BRANCH_PT : OVL_DOVETAIL);

  //  merge: Left_Delta (sign-flipped when T played "A"), then the right deltas, negated (extend.C:164-175)
  __syncwarp();
  int32_t *ldelta = C.ldelta; const int32_t *rdelta = C.rdelta;
  if (l_negate) for (int i = lane; i < llen; i += 32) ldelta[i] = -ldelta[i];
  for (int i = lane; i < rlen; i += 32) {
    int rd = rdelta[i]; if (r_negate) rd = -rd;
    int v;
    if (i == 0) v = (rd > 0) ? -(rd + leftover + m_len) : -(rd - leftover - m_len);
    else        v = -rd;
    ldelta[llen + i] = v;
  }
  ldelta_len = llen + rlen;
  __syncwarp();
  return kind;
}

//  Lies_On_Alignment (Process_String_Overlaps.C:262-281); each lane walks the shared delta list for its own seed.
__device__ __forceinline__ bool lies_on_alignment(const int32_t *ld, int ld_len, int start, int offset, int s_lo, int t_lo) {
  int diag = t_lo - s_lo, new_diag = offset - start;
  for (int i = 0; i < ld_len; i++) {
    int dl = ld[i];
    s_lo += dl < 0 ? -dl : dl;
    if (start < s_lo) break;
    if (dl < 0) diag++;
    else { s_lo++; diag--; }
  }
  int x = new_diag - diag; if (x < 0) x = -x;
  return x <= OVL_SHIFT_SLACK;
}

//  Add_Overlap (Process_String_Overlaps.C:177-244)
__device__ void add_overlap(const DevParams &P, int s_lo, int s_hi, int t_lo, int t_hi, double qual, int delta_ct, OvlOlap *o, int &ct) {
  if (!P.partial) {
    int new_diag = t_lo - s_lo;
    for (int i = 0; i < ct; i++) {
      int old_diag = o[i].t_lo - o[i].s_lo;
      if ((new_diag >  0 && old_diag >  0 && o[i].t_right_boundary - new_diag - o[i].s_left_boundary >= OVL_MIN_INTERSECTION) ||
          (new_diag <= 0 && old_diag <= 0 && o[i].s_right_boundary + new_diag - o[i].t_left_boundary >= OVL_MIN_INTERSECTION)) {
        if (new_diag < o[i].min_diag) o[i].min_diag = new_diag;
        if (new_diag > o[i].max_diag) o[i].max_diag = new_diag;
        if (s_lo < o[i].s_left_boundary)  o[i].s_left_boundary  = s_lo;
        if (s_hi > o[i].s_right_boundary) o[i].s_right_boundary = s_hi;
        if (t_lo < o[i].t_left_boundary)  o[i].t_left_boundary  = t_lo;
        if (t_hi > o[i].t_right_boundary) o[i].t_right_boundary = t_hi;
        if (qual < o[i].quality) {
          o[i].s_lo = s_lo; o[i].s_hi = s_hi; o[i].t_lo = t_lo; o[i].t_hi = t_hi;
          o[i].quality = qual; o[i].delta_ct = delta_ct;
        }
        return;
      }
    }
  }
  if (ct >= OVL_MAX_DISTINCT_OLAPS) return;
  OvlOlap &n = o[ct];
  n.s_lo = n.s_left_boundary  = s_lo;  n.s_hi = n.s_right_boundary = s_hi;
  n.t_lo = n.t_left_boundary  = t_lo;  n.t_hi = n.t_right_boundary = t_hi;
  n.quality = qual; n.delta_ct = delta_ct;
  n.min_diag = n.max_diag = t_lo - s_lo;
  ct++;
}

__device__ void combine_into_one(OvlOlap *o, int ct, int *deleted) {            // :42-96
  int best = 0;
  int min_diag = o[0].min_diag, max_diag = o[0].max_diag;
  int slb = o[0].s_left_boundary, srb = o[0].s_right_boundary, tlb = o[0].t_left_boundary, trb = o[0].t_right_boundary;
  for (int i = 1; i < ct; i++) {
    int leni = 1 + min(o[i].s_hi - o[i].s_lo, o[i].t_hi - o[i].t_lo);
    int lenb = 1 + min(o[best].s_hi - o[best].s_lo, o[best].t_hi - o[best].t_lo);
    if (o[i].quality < o[best].quality || (o[i].quality == o[best].quality && leni > lenb)) best = i;
    min_diag = min(min_diag, o[i].min_diag); max_diag = max(max_diag, o[i].max_diag);
    slb = min(slb, o[i].s_left_boundary); srb = max(srb, o[i].s_right_boundary);
    tlb = min(tlb, o[i].t_left_boundary); trb = max(trb, o[i].t_right_boundary);
  }
  o[best].min_diag = min_diag; o[best].max_diag = max_diag;
  o[best].s_left_boundary = slb; o[best].s_right_boundary = srb; o[best].t_left_boundary = tlb; o[best].t_right_boundary = trb;
  for (int i = 0; i < ct; i++) deleted[i] = (i != best);
}

__device__ void merge_intersecting(OvlOlap *p, int ct, int *deleted) {          // :108-162
  for (int i = 0; i < ct - 1; i++)
    for (int j = i + 1; j < ct; j++) {
      if (deleted[i] || deleted[j]) continue;
      int lo_diag = p[i].min_diag, hi_diag = p[i].max_diag;
      if ((lo_diag <= 0 && p[j].min_diag > 0) || (lo_diag > 0 && p[j].min_diag <= 0)) continue;
      if ((lo_diag >= 0 && p[j].t_right_boundary - lo_diag - p[j].s_left_boundary >= OVL_MIN_INTERSECTION) ||
          (lo_diag <= 0 && p[j].s_right_boundary + lo_diag - p[j].t_left_boundary >= OVL_MIN_INTERSECTION) ||
          (hi_diag >= 0 && p[j].t_right_boundary - hi_diag - p[j].s_left_boundary >= OVL_MIN_INTERSECTION) ||
          (hi_diag <= 0 && p[j].s_right_boundary + hi_diag - p[j].t_left_boundary >= OVL_MIN_INTERSECTION)) {
        int keep, disc;
        if (p[i].quality < p[j].quality) { keep = i; disc = j; deleted[j] = 1; }
        else                             { keep = j; disc = i; deleted[i] = 1; }
        p[keep].min_diag = min(p[keep].min_diag, p[disc].min_diag);
        p[keep].max_diag = max(p[keep].max_diag, p[disc].max_diag);
        p[keep].s_left_boundary  = min(p[keep].s_left_boundary,  p[disc].s_left_boundary);
        p[keep].s_right_boundary = max(p[keep].s_right_boundary, p[disc].s_right_boundary);
        p[keep].t_left_boundary  = min(p[keep].t_left_boundary,  p[disc].t_left_boundary);
        p[keep].t_right_boundary = max(p[keep].t_right_boundary, p[disc].t_right_boundary);
      }
    }
}

__device__ void choose_best_partial(OvlOlap *o, int ct, int *deleted) {         // :291-311
  int best = 0;
  double mbest = __dmul_rn(__dsub_rn(1.0, o[0].quality), (double)(2 + o[0].s_hi - o[0].s_lo + o[0].t_hi - o[0].t_lo));
  for (int i = 1; i < ct; i++) {
    double mb = __dmul_rn(__dsub_rn(1.0, o[i].quality), (double)(2 + o[i].s_hi - o[i].s_lo + o[i].t_hi - o[i].t_lo));
    if (mbest < mb || (mbest == mb && o[i].quality < o[best].quality)) best = i;
  }
  for (int i = 0; i < ct; i++) deleted[i] = (i != best);
}

//  CTA prologue: thread 0 publishes the job parameters, lane 0 of every warp binds the warp's scratch.
__device__ __forceinline__ void bind_warp_ctl(const DevParams &P, const ExtScratch &X, int gwarp) {
  if (threadIdx.x == 0) *reinterpret_cast<DevParams *>(ext_sm + EXT_SM_PARAMS) = P;
  if ((threadIdx.x & 31) == 0 && gwarp < X.n_warps) {
    WarpCtl &C = sh_ctl();
    C.gring_cap = X.gring_cap;
    C.gring0 = X.gring + (size_t)gwarp * 2 * X.gring_cap;
    C.gring1 = C.gring0 + X.gring_cap;
    C.arena = X.arena + (size_t)gwarp * X.arena_cap;  C.arena_cap = X.arena_cap;
    const size_t st = (size_t)X.emax + 2;
    C.row_meta = X.row_meta + gwarp * st;
    C.path = X.path + gwarp * st;  C.ival = X.ival + gwarp * st;  C.ikc = X.ikc + gwarp * st;
    C.ldelta = X.ldelta + gwarp * st;  C.rdelta = X.rdelta + gwarp * st;
    C.emax = X.emax;
    C.distinct_ct = 0;
    C.cells = C.calls = C.c_with = C.c_without = C.c_multi = C.c_total = C.c_cont = C.c_dove = 0;
  }
  __syncthreads();
}

//  Persistent kernel: warps pull pairs from a global cursor (pairs differ wildly in cost).
template <int MIN_BLOCKS, int ILP>
__global__ void __launch_bounds__(EXT_THREADS, MIN_BLOCKS)
k_extend_pairs(DevParams P_, ExtScratch X, const PairRec *__restrict__ pairs, const uint32_t *__restrict__ order, uint64_t n_pairs,
               const int32_t *__restrict__ seed_start, const int32_t *__restrict__ seed_off, const int32_t *__restrict__ seed_len,
               uint8_t *seed_alive,
               const uint64_t *__restrict__ rfwd, const uint64_t *__restrict__ rrc, const uint64_t *__restrict__ rwoff,
               const uint32_t *__restrict__ rlen, uint32_t ref_first_id,
               const uint64_t *__restrict__ hfwd, const uint64_t *__restrict__ hrc, const uint64_t *__restrict__ hwoff,
               const uint32_t *__restrict__ hlen, uint32_t hash_first_id,
               ovlb_record *records, uint64_t rec_cap, unsigned long long *work, unsigned long long *counters) {
  int lane_ = threadIdx.x & 31;
#ifndef OVL_EXT_NO_LANE_PIN
  //  keep the lane number in a register: at 64 registers per thread the compiler otherwise re-reads SR_TID.X (S2R, a
  //  slow special-register read) and masks it again at two dozen places inside the row loop (ncu: 3.3 % of the
  //  kernel's instructions were attributed to this line)
  asm volatile("" : "+r"(lane_));
#endif
  int wib_ = threadIdx.x >> 5;
#ifndef OVL_EXT_NO_LANE_PIN
  asm volatile("" : "+r"(wib_));
#endif
  const int lane = lane_, wib = wib_;
  const int gwarp = blockIdx.x * EXT_WARPS + wib;
  bind_warp_ctl(P_, X, gwarp);
  if (gwarp >= X.n_warps) return;
  const DevParams &P = sh_params();
  WarpCtl &C = sh_ctl();
  unsigned long long *err_flags = &counters[CT_ERR_FLAGS];
  unsigned long long t_begin;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));

  while (true) {
    unsigned long long pi = 0;
    if (lane == 0) pi = atomicAdd(&work[1], 1ull);
    pi = __shfl_sync(FULL, pi, 0);
    if (pi >= n_pairs) break;
    const PairRec pr = pairs[order ? (unsigned long long)order[pi] : pi];
    if (pr.n_seeds == 0) continue;

    __syncwarp();
    if (lane == 0) {
      const uint64_t rw = rwoff[pr.ref_idx], hw = hwoff[pr.hash_idx];
      C.s_len = (int)rlen[pr.ref_idx];
      C.s_fwd = (pr.dir ? rrc : rfwd) + rw;
      C.s_rc  = (pr.dir ? rfwd : rrc) + rw;
      C.t_len = (int)hlen[pr.hash_idx];
      C.t_fwd = hfwd + hw;
      C.t_rc  = hrc + hw;
      C.s_id = ref_first_id + pr.ref_idx; C.t_id = hash_first_id + pr.hash_idx;
      C.seed_begin = pr.seed_begin; C.n_seeds = pr.n_seeds; C.dir = pr.dir; C.consistent = pr.consistent;
      C.distinct_ct = 0;
    }
    __syncwarp();
    const int64_t sb = pr.seed_begin;
    const int ns = pr.n_seeds;
    int remaining = ns;

    while (remaining > 0) {
      //  longest seed, first in list order on ties (Process_String_Overlaps.C:424-431)
      unsigned long long best = 0;
      for (int i = lane; i < ns; i += 32)
        if (seed_alive[sb + i]) {
          unsigned long long key = ((unsigned long long)(uint32_t)seed_len[sb + i] << 32) | (uint32_t)(0x7fffffff - i);
          if (key > best) best = key;
        }
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, best, o); if (x > best) best = x; }
      const int li = 0x7fffffff - (int)(uint32_t)best;
      const int m_start = seed_start[sb + li], m_offset = seed_off[sb + li], m_len = seed_len[sb + li];

      int s_lo, s_hi, t_lo, t_hi, errors, ld_len;
      const int kind = warp_extend_alignment<ILP>(m_start, m_offset, m_len, s_lo, s_hi, t_lo, t_hi, errors, ld_len, err_flags, lane, wib);

      const bool usable = (kind == OVL_DOVETAIL) || P.partial;
      if (lane == 0 && usable && 1 + s_hi - s_lo >= P.min_olap_len && 1 + t_hi - t_lo >= P.min_olap_len) {
        int olap_len = 1 + min(s_hi - s_lo, t_hi - t_lo);
        double quality = __ddiv_rn((double)errors, (double)olap_len);
        if (errors <= (int)ceil(__dmul_rn((double)olap_len, P.erate)))
          add_overlap(P, s_lo, s_hi, t_lo, t_hi, quality, ld_len, C.distinct, C.distinct_ct);
      }

      if (C.consistent) break;                        // all remaining seeds are dropped (:473-474)

      //  remove the seed just used and every seed lying on the alignment (:476-490)
      const int32_t *ldelta = C.ldelta;
      int removed = 0;
      for (int i0 = 0; i0 < ns; i0 += 32) {
        const int i = i0 + lane;
        bool rm = false;
        if (i < ns && seed_alive[sb + i]) {
          if (i == li) rm = true;
          else if (usable) {
            int st = seed_start[sb + i], ln = seed_len[sb + i];
            if (s_lo - OVL_SHIFT_SLACK <= st && st + ln <= (s_hi + 1) + OVL_SHIFT_SLACK - 1 &&
                lies_on_alignment(ldelta, ld_len, st, seed_off[sb + i], s_lo, t_lo))
              rm = true;
          }
          if (rm) seed_alive[sb + i] = 0;
        }
        removed += __popc(__ballot_sync(FULL, rm));
      }
      remaining -= removed;
      __syncwarp();
    }

    //  Combine / merge / choose and output: warp-uniform scalar work, done by lane 0
    if (lane == 0) {
      const int distinct_ct = C.distinct_ct;
      OvlOlap *distinct = C.distinct;
      int outputs = 0;
      if (distinct_ct > 0) {
        int deleted[OVL_MAX_DISTINCT_OLAPS] = {0, 0, 0};
        if (P.partial) { if (P.unique) choose_best_partial(distinct, distinct_ct, deleted); }
        else           { if (P.unique) combine_into_one(distinct, distinct_ct, deleted); else merge_intersecting(distinct, distinct_ct, deleted); }
        for (int i = 0; i < distinct_ct; i++)
          if (!deleted[i]) {
            uint32_t a, b; uint64_t w0, w1;
            if (P.partial) {
              ovl_output_partial(C.s_id, C.t_id, C.dir, distinct[i], C.s_len, C.t_len, &a, &b, &w0, &w1);
            } else {
              int cont = ovl_output_overlap(C.s_id, C.s_len, C.dir, C.t_id, C.t_len, distinct[i], &a, &b, &w0, &w1);
              if (cont) C.c_cont++; else C.c_dove++;
            }
            C.c_total++;
            unsigned long long ri = atomicAdd(&work[2], 1ull);
            if (ri < rec_cap) { ovlb_record r; r.a_iid = a; r.b_iid = b; r.dat0 = w0; r.dat1 = w1; records[ri] = r; }
            else atomicOr(err_flags, 2ull);
            outputs++;
          }
      }
      if (outputs == 0) C.c_without++;
      else { C.c_with++; if (outputs > 1) C.c_multi++; }
    }
  }

  if (lane == 0) {
    if (C.c_without) atomicAdd(&counters[CT_HITS_WITHOUT], C.c_without);
    if (C.c_with)    atomicAdd(&counters[CT_HITS_WITH], C.c_with);
    if (C.c_multi)   atomicAdd(&counters[CT_MULTI], C.c_multi);
    if (C.c_total)   atomicAdd(&counters[CT_TOTAL], C.c_total);
    if (C.c_cont)    atomicAdd(&counters[CT_CONTAINED], C.c_cont);
    if (C.c_dove)    atomicAdd(&counters[CT_DOVETAIL], C.c_dove);
    if (C.cells)     atomicAdd(&counters[CT_DP_CELLS], C.cells);
    if (C.calls)     atomicAdd(&counters[CT_EXT_CALLS], C.calls);
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    atomicAdd(&counters[CT_EXT_BUSY], t_end - t_begin);
  }
}

//  Debug tap: one warp per explicit seed; out7 = s_lo, s_hi, t_lo, t_hi, errors, kind, delta_ct.
__global__ void __launch_bounds__(EXT_THREADS)
k_debug_extend(DevParams P_, ExtScratch X, uint32_t n, const uint32_t *__restrict__ ref_index, const int32_t *__restrict__ dir,
               const uint32_t *__restrict__ hash_index, const int32_t *__restrict__ m_start, const int32_t *__restrict__ m_offset,
               const int32_t *__restrict__ m_len,
               const uint64_t *__restrict__ rfwd, const uint64_t *__restrict__ rrc, const uint64_t *__restrict__ rwoff, const uint32_t *__restrict__ rlen,
               const uint64_t *__restrict__ hfwd, const uint64_t *__restrict__ hrc, const uint64_t *__restrict__ hwoff, const uint32_t *__restrict__ hlen,
               int32_t *out7, int32_t *deltas, uint32_t delta_stride, unsigned long long *work, unsigned long long *counters) {
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * EXT_WARPS + (threadIdx.x >> 5);
  bind_warp_ctl(P_, X, gwarp);
  if (gwarp >= X.n_warps) return;
  WarpCtl &C = sh_ctl();
  while (true) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(&work[1], 1ull);
    i = __shfl_sync(FULL, i, 0);
    if (i >= n) break;
    const uint32_t ri = ref_index[i], hi = hash_index[i];
    const int dr = dir[i];
    __syncwarp();
    if (lane == 0) {
      C.s_len = (int)rlen[ri]; C.s_fwd = (dr ? rrc : rfwd) + rwoff[ri]; C.s_rc = (dr ? rfwd : rrc) + rwoff[ri];
      C.t_len = (int)hlen[hi]; C.t_fwd = hfwd + hwoff[hi]; C.t_rc = hrc + hwoff[hi];
    }
    __syncwarp();
    int s_lo, s_hi, t_lo, t_hi, errors, ld_len;
    int kind = warp_extend_alignment<2>(m_start[i], m_offset[i], m_len[i], s_lo, s_hi, t_lo, t_hi, errors, ld_len,
                                     &counters[CT_ERR_FLAGS], lane, (int)(threadIdx.x >> 5));
    if (lane == 0) {
      int32_t *o = out7 + 7 * i;
      o[0] = s_lo; o[1] = s_hi; o[2] = t_lo; o[3] = t_hi; o[4] = errors; o[5] = kind; o[6] = ld_len;
    }
    const int32_t *ldelta = C.ldelta;
    if (deltas)
      for (int j = lane; j < ld_len && j < (int)delta_stride; j += 32) deltas[(size_t)i * delta_stride + j] = ldelta[j];
    __syncwarp();
  }
  if (lane == 0) {
    if (C.cells) atomicAdd(&counters[CT_DP_CELLS], C.cells);
    if (C.calls) atomicAdd(&counters[CT_EXT_CALLS], C.calls);
  }
}

// ------------------------------------------------------------------------------------------------
//  host side
// ------------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ovl_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return OVLB_ERR_CUDA; } } while (0)

static void free_ext(ExtScratch &X) {
  void *ptrs[] = { X.arena, X.row_meta, X.gring, X.path, X.ival, X.ikc, X.ldelta, X.rdelta };
  for (void *p : ptrs) if (p) cudaFree(p);
  X = ExtScratch();
}

//  Size the per-warp scratch for the longest read currently on the device.
int ovl_prepare_ext_scratch(ovlb_ctx *c) {
  uint32_t maxlen = c->hash.max_len > c->ref.max_len ? c->hash.max_len : c->ref.max_len;
  if (maxlen < 64) maxlen = 64;
  int emax = (int)ceil((double)maxlen * c->P.max_erate) + 2;          // Error_Limit <= Error_Bound[len] <= this
  if ((uint32_t)emax + 1 > c->P.n_edit_match_limit) emax = (int)c->P.n_edit_match_limit - 1;
  if (c->ext.arena && c->ext.emax >= emax) return OVLB_OK;
  free_ext(c->ext);
  ExtScratch X;
  X.emax = emax;
  //  worst-case code words of one extension: sum_e ceil((2e+1)/32) + slack
  uint64_t worst = 0;
  for (int e = 1; e <= emax; e++) worst += (uint64_t)(2 * e + 1 + 31) / 32;
  X.arena_cap = worst + 64;
  uint32_t gcap = 64; while (gcap < (uint32_t)(2 * emax + 16)) gcap <<= 1;
  X.gring_cap = gcap;
  const uint64_t per_warp = X.arena_cap * 8 + (uint64_t)gcap * 8 + (uint64_t)(emax + 2) * (4 + 4 + 1 + 4 + 4 + 4 + 4);
  int want_warps = c->sm_count * 32;                                    // up to 4 CTAs of 8 warps per SM
  uint64_t budget = c->mem_budget / 4;
  if (per_warp * want_warps > budget) want_warps = (int)(budget / per_warp);
  want_warps = (want_warps / EXT_WARPS) * EXT_WARPS;
  if (want_warps < EXT_WARPS) { ovl_set_error("not enough device memory for the extension scratch (reads too long for the budget)"); return OVLB_ERR_CAPACITY; }
  X.n_warps = want_warps;
  const size_t st = (size_t)emax + 2;
  CK(cudaMalloc((void **)&X.arena, (size_t)want_warps * X.arena_cap * sizeof(uint2)));
  CK(cudaMalloc((void **)&X.gring, ((size_t)want_warps * 2 * gcap + SRING_PAD) * 4));   // + slack: masked lanes read past the last ring
  CK(cudaMalloc((void **)&X.row_meta, want_warps * st * 8));
  CK(cudaMalloc((void **)&X.path, want_warps * st));
  CK(cudaMalloc((void **)&X.ival, want_warps * st * 4));
  CK(cudaMalloc((void **)&X.ikc, want_warps * st * 4));
  CK(cudaMalloc((void **)&X.ldelta, want_warps * st * 4));
  CK(cudaMalloc((void **)&X.rdelta, want_warps * st * 4));
  c->ext = X;
  const int smem = EXT_SM_BYTES;
  CK(cudaFuncSetAttribute((k_extend_pairs<3, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute((k_extend_pairs<4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute((k_extend_pairs<3, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute((k_extend_pairs<4, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_debug_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return OVLB_OK;
}

int ovl_extend_pairs(ovlb_ctx *c) {
  if (c->n_pairs == 0) return OVLB_OK;
  int rc = ovl_prepare_ext_scratch(c);
  if (rc) return rc;
  //  records: at most 1 (unique) or MAX_DISTINCT_OLAPS per pair
  uint64_t need = c->n_pairs * (c->P.unique_per_pair ? 1 : OVL_MAX_DISTINCT_OLAPS) + 16;
  if (need > c->rec_cap) {
    if (c->d_records) cudaFree(c->d_records); c->d_records = nullptr;
    uint64_t want = need * 5 / 4;
    CK(cudaMalloc((void **)&c->d_records, want * sizeof(ovlb_record)));
    c->rec_cap = want;
  }
  const int smem = EXT_SM_BYTES;
  int blocks = c->ext.n_warps / EXT_WARPS;
  uint64_t need_blocks = (c->n_pairs + EXT_WARPS - 1) / EXT_WARPS;
  if ((uint64_t)blocks > need_blocks) blocks = (int)need_blocks;
  static const int min_blocks = [] { const char *ev = getenv("OVLB_EXT_BLOCKS"); const int v = ev ? atoi(ev) : 4; return (v < 3 || v > 4) ? 4 : v; }();
  //  two 32-cell groups in flight per iteration pay on wide bands (noisy reads: +2 % at --maxerate 0.06) and cost on
  //  HiFi-like reads, whose rows are 3-25 cells wide (14.5 -> 16.2 ms per C2 tile): chosen by the error rate of the job
  static const int ilp_env = [] { const char *ev = getenv("OVLB_EXT_ILP"); return ev ? atoi(ev) : 0; }();
  const int ilp = ilp_env ? (ilp_env == 1 ? 1 : 2) : (c->P.max_erate >= 0.025 ? 2 : 1);
  if ((uint64_t)c->sm_count * min_blocks < (uint64_t)blocks) blocks = c->sm_count * min_blocks;
#define EXT_LAUNCH(MB, IL) k_extend_pairs<MB, IL><<<blocks, EXT_THREADS, smem, c->stream>>>( \
      c->dp, c->ext, c->pairs, c->pair_order, c->n_pairs, c->seed_start, c->seed_off, c->seed_len, c->seed_alive, \
      c->ref.fwd, c->ref.rc, c->ref.woff, c->ref.len, c->ref.first_id, \
      c->hash.fwd, c->hash.rc, c->hash.woff, c->hash.len, c->hash.first_id, \
      c->d_records, c->rec_cap, c->d_work, c->d_counters->v)
  if (min_blocks == 4) { if (ilp == 2) EXT_LAUNCH(4, 2); else EXT_LAUNCH(4, 1); }
  else                 { if (ilp == 2) EXT_LAUNCH(3, 2); else EXT_LAUNCH(3, 1); }
#undef EXT_LAUNCH
  c->launches++;
  c->ext_warps_launched = (uint64_t)blocks * EXT_WARPS;
  CK(cudaGetLastError());
  return OVLB_OK;
}

int ovl_debug_extend(ovlb_ctx *c, uint32_t n, const uint32_t *ref_index, const int32_t *dir, const uint32_t *hash_index,
                     const int32_t *seed_start, const int32_t *seed_offset, const int32_t *seed_len,
                     int32_t *out7, int32_t *deltas, uint32_t delta_stride) {
  int rc = ovl_prepare_ext_scratch(c);
  if (rc) return rc;
  uint32_t *d_u = nullptr; int32_t *d_i = nullptr, *d_out = nullptr, *d_del = nullptr;
  CK(cudaMalloc((void **)&d_u, (size_t)n * 2 * 4));
  CK(cudaMalloc((void **)&d_i, (size_t)n * 4 * 4));
  CK(cudaMalloc((void **)&d_out, (size_t)n * 7 * 4));
  if (deltas) { CK(cudaMalloc((void **)&d_del, (size_t)n * delta_stride * 4)); CK(cudaMemsetAsync(d_del, 0, (size_t)n * delta_stride * 4, c->stream)); }
  CK(cudaMemcpyAsync(d_u, ref_index, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_u + n, hash_index, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_i, dir, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_i + n, seed_start, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_i + 2 * n, seed_offset, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_i + 3 * n, seed_len, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->d_work, 0, 64, c->stream));
  const int smem = EXT_SM_BYTES;
  int blocks = c->ext.n_warps / EXT_WARPS;
  uint64_t need_blocks = ((uint64_t)n + EXT_WARPS - 1) / EXT_WARPS;
  if ((uint64_t)blocks > need_blocks) blocks = (int)need_blocks;
  k_debug_extend<<<blocks, EXT_THREADS, smem, c->stream>>>(
      c->dp, c->ext, n, d_u, d_i, d_u + n, d_i + n, d_i + 2 * n, d_i + 3 * n,
      c->ref.fwd, c->ref.rc, c->ref.woff, c->ref.len, c->hash.fwd, c->hash.rc, c->hash.woff, c->hash.len,
      d_out, d_del, delta_stride, c->d_work, c->d_counters->v);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpyAsync(out7, d_out, (size_t)n * 7 * 4, cudaMemcpyDeviceToHost, c->stream));
  if (deltas) CK(cudaMemcpyAsync(deltas, d_del, (size_t)n * delta_stride * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_u); cudaFree(d_i); cudaFree(d_out); if (d_del) cudaFree(d_del);
  return OVLB_OK;
}
