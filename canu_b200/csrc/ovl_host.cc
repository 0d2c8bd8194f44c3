//  ovl_host.cc -- host-only (no CUDA) helpers of the C ABI: parameter derivation and read packing.
//
//  Follows overlapInCore.C:380-411,454-459 (flag fix-ups), liboverlap/prefixEditDistance.C:23-107
//  (MAX_ERRORS, slope, branch value) and liboverlap/Binomial_Bound.C:36-188 (Edit_Match_Limit).
//  All of it is FP64 libm arithmetic whose results feed integer tables; it must be evaluated in the
//  reference's operation order (and without FMA contraction: compiled with -ffp-contract=off).
#include "../../include/ovlb200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

void ovl_set_error(const std::string &msg);

namespace {

//  Smallest n >= start with P[at least e errors in n trials at rate p] > 1e-4  (Binomial_Bound.C:36-99).
int binomial_bound(int e, double p, int start) {
  const double kProbBound = 1e-4, kNormalZ = 3.62;
  const double q = 1.0 - p;
  if (start < e) start = e;
  for (int n = start; n < (int)OVLB_MAX_READLEN; n++) {
    if (n <= 35) {
      double sum = 0.0, pp = 1.0, qq = pow(q, n);
      int coeff = 1, ct = 0;
      for (int k = 0; k < e && 1.0 - sum > kProbBound; k++) {
        sum += coeff * pp * qq;
        coeff *= n - ct;
        coeff /= ++ct;
        pp *= p;
        qq /= q;
      }
      if (1.0 - sum > kProbBound) return n;
    } else {
      double z = (e - 0.5 - n * p) / sqrt(n * p * q);
      if (z <= kNormalZ) return n;
      double sum = 0.0, mu = 1.0, fact = 1.0, pc = exp(-n * p);
      for (int k = 0; k < e; k++) {
        sum += mu * pc / fact;
        mu *= n * p;
        fact *= k + 1;
      }
      if (1.0 - sum > kProbBound) return n;
    }
  }
  return (int)OVLB_MAX_READLEN;
}

//  Initialize_Match_Limit (Binomial_Bound.C:104-188), AS_MAX_READLEN_BITS == 21 slope constants.
void match_limit_table(int32_t *ml, double erate, int32_t max_errors) {
  int32_t e = 0, s = 1;
  const int32_t exact = max_errors < 2000 ? max_errors : 2000;
  while (e <= 1) ml[e++] = 0;
  for (; e < exact; e++) {
    s = binomial_bound(e - 1, erate, s);
    ml[e] = s - 1;
  }
  const double slope = 0.982064188397525 / erate + 0.067835741959926;
  double v = ml[e - 1] + slope;
  for (; e < max_errors; e++) {
    ml[e] = (int32_t)ceil(v);
    v += slope;
  }
}

}  // namespace

struct ovlb_reads_owner {
  std::vector<uint8_t>  packed;
  std::vector<uint64_t> boff;
  std::vector<uint32_t> len, n_read, n_pos;
  ovlb_reads view;
};

extern "C" {

double ovlb_parse_erate(const char *text) { return (double)strtof(text, nullptr); }

int ovlb_params_init(ovlb_params *p, uint32_t kmer_len, double max_erate, double align_noise,
                     int partial, int unique_per_pair, int min_olap_len, int no_hopeless, int min_kmers,
                     uint32_t max_read_len) {
  if (!p) { ovl_set_error("ovlb_params_init: null"); return OVLB_ERR_ARG; }
  if (kmer_len < 2 || kmer_len > 30) { ovl_set_error("ovlb_params_init: kmer_len must be in 2..30 (k-mer + 3 class bits + 1 sentinel bit must fit 64 bits)"); return OVLB_ERR_ARG; }
  if (!(max_erate > 0.0) || max_erate >= 1.0) { ovl_set_error("ovlb_params_init: max_erate must be in (0,1)"); return OVLB_ERR_ARG; }
  if (align_noise == 0.0) align_noise = 1.0;
  memset(p, 0, sizeof(*p));
  p->kmer_len = kmer_len;
  p->partial = partial ? 1 : 0;
  p->unique_per_pair = unique_per_pair ? 1 : 0;
  p->min_olap_len = min_olap_len;
  p->use_hopeless_check = no_hopeless ? 0 : 1;
  if (max_erate > 0.06) p->use_hopeless_check = 0;
  p->minkmers_exp_factor = exp(-1.0 * (double)kmer_len * max_erate);
  p->filter_by_kmer_count = 0;
  if (min_kmers) {
    //  int(floor(exp(-K*erate) * (Min_Olap_Len - Kmer_Len + 1))), the subtraction done in uint64 (overlapInCore.C:411)
    uint64_t span = (uint64_t)(int64_t)min_olap_len - (uint64_t)kmer_len + 1;
    p->filter_by_kmer_count = (uint64_t)(int)floor(p->minkmers_exp_factor * (double)span);
  }
  p->max_erate = max_erate;
  p->branch_match_value = max_erate / (1 + max_erate);
  p->min_branch_tail_slope = (max_erate > 0.06) ? 1.0 : 0.20;
  const uint32_t max_errors = 1 + (uint32_t)(int)ceil(max_erate * OVLB_MAX_READLEN);
  int32_t *ml = (int32_t *)calloc((size_t)max_errors + 1, sizeof(int32_t));
  if (!ml) { ovl_set_error("ovlb_params_init: out of memory"); return OVLB_ERR_ARG; }
  match_limit_table(ml, max_erate * align_noise, (int32_t)max_errors);
  p->edit_match_limit = ml;
  p->n_edit_match_limit = max_errors;
  p->max_read_len = max_read_len ? max_read_len : OVLB_MAX_READLEN;
  p->device_mem_budget = 0;
  return OVLB_OK;
}

void ovlb_params_free(ovlb_params *p) {
  if (p && p->edit_match_limit) { free((void *)p->edit_match_limit); p->edit_match_limit = nullptr; }
}

int ovlb_pack_reads(const char *bases, const uint64_t *offsets, const uint32_t *lens, uint32_t n_reads,
                    uint32_t first_read_id, uint32_t min_len, ovlb_reads_owner **out) {
  if (!out || (n_reads && (!bases || !offsets || !lens))) { ovl_set_error("ovlb_pack_reads: null argument"); return OVLB_ERR_ARG; }
  ovlb_reads_owner *o = new ovlb_reads_owner();
  o->boff.resize(n_reads); o->len.resize(n_reads);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_reads; i++) {
    uint32_t L = lens[i] < min_len ? 0 : lens[i];
    o->boff[i] = total; o->len[i] = L;
    total += ((uint64_t)L + 3) / 4;
  }
  o->packed.assign(total + 8, 0);
  for (uint32_t i = 0; i < n_reads; i++) {
    const uint32_t L = o->len[i];
    const char *s = bases + offsets[i];
    uint8_t *dst = o->packed.data() + o->boff[i];
    for (uint32_t j = 0; j < L; j++) {
      unsigned code;
      switch (s[j]) {
        case 'A': case 'a': code = 0; break;
        case 'C': case 'c': code = 1; break;
        case 'G': case 'g': code = 2; break;
        case 'T': case 't': code = 3; break;
        case 'N': case 'n': code = 0; o->n_read.push_back(i); o->n_pos.push_back(j); break;
        default:
          delete o;
          ovl_set_error("ovlb_pack_reads: read " + std::to_string(first_read_id + i) + " has a base that is not ACGTN at position " + std::to_string(j));
          return OVLB_ERR_ARG;
      }
      dst[j >> 2] |= (uint8_t)(code << (6 - 2 * (j & 3)));
    }
  }
  o->view.packed = o->packed.data();
  o->view.packed_bytes = total;
  o->view.byte_offset = o->boff.data();
  o->view.len = o->len.data();
  o->view.n_reads = n_reads;
  o->view.first_read_id = first_read_id;
  o->view.n_read = o->n_read.data();
  o->view.n_pos = o->n_pos.data();
  o->view.n_n = o->n_read.size();
  *out = o;
  return OVLB_OK;
}

const ovlb_reads *ovlb_reads_view(const ovlb_reads_owner *o) { return o ? &o->view : nullptr; }
void ovlb_reads_free(ovlb_reads_owner *o) { delete o; }

int ovlb_kmer_keys(const char *kmer, uint32_t kmer_len, uint64_t *fwd_key, uint64_t *rc_key) {
  if (!kmer || !fwd_key || !rc_key || kmer_len < 2 || kmer_len > 30) { ovl_set_error("ovlb_kmer_keys: bad argument (kmer_len must be in 2..30)"); return OVLB_ERR_ARG; }
  uint64_t f = 0, r = 0;
  for (uint32_t j = 0; j < kmer_len; j++) {
    uint64_t code;
    switch (kmer[j]) {
      case 'A': case 'a': code = 0; break;
      case 'C': case 'c': code = 1; break;
      case 'G': case 'g': code = 2; break;
      case 'T': case 't': code = 3; break;
      default: ovl_set_error("ovlb_kmer_keys: non-ACGT base in skip k-mer"); return OVLB_ERR_ARG;
    }
    f |= code << (2 * j);
    r |= (3 - code) << (2 * (kmer_len - 1 - j));
  }
  *fwd_key = f; *rc_key = r;
  return OVLB_OK;
}

//  overlapInCorePartition's partitionLength() (overlapInCorePartition.C:127-257), default library case
//  (no -H/-R restriction): hash blocks of at least hash_block_len bytes (one per base plus one per read),
//  each crossed with ref blocks of at least ref_block_len bases; a ref block never extends past the last
//  read of its hash block, because only refID < hashID pairs are computed.
//
//  strict_reference = 1 reproduces the reference loop bounds exactly (its `while (hashBeg < hashMax)` /
//  `while (refBeg < refMax)` never start a block on the last read); 0 covers every read.
int ovlb_plan_tiles(const uint32_t *read_len, uint32_t n_reads, uint32_t min_olap_len,
                    uint64_t hash_block_len, uint64_t ref_block_len,
                    uint32_t hash_min, uint32_t hash_max, uint32_t ref_min, uint32_t ref_max,
                    int strict_reference, ovlb_tile *out, uint64_t out_cap, uint64_t *n_out) {
  if (!read_len || !n_out) { ovl_set_error("ovlb_plan_tiles: null argument"); return OVLB_ERR_ARG; }
  if (hash_block_len == 0 || ref_block_len == 0) { ovl_set_error("ovlb_plan_tiles: block lengths must be positive"); return OVLB_ERR_ARG; }
  hash_min = std::max(hash_min, 1u); ref_min = std::max(ref_min, 1u);
  hash_max = std::min(hash_max, n_reads); ref_max = std::min(ref_max, n_reads);
  //  a read counts for a block only if it is long enough to be overlapped; a hashed read also costs its terminator
  auto usable = [&](uint32_t id) -> uint64_t { return read_len[id] >= min_olap_len ? (uint64_t)read_len[id] : 0; };
  auto hashed = [&](uint32_t id) -> uint64_t { return read_len[id] >= min_olap_len ? (uint64_t)read_len[id] + 1 : 0; };
  //  the reference's loops only START a block on a read strictly below the range's last one (so its last read can be
  //  left out of the grid); with strict_reference == 0 every read of the range is covered
  const uint32_t start_slack = strict_reference ? 0u : 1u;
  uint64_t all_hashed = 0;
  for (uint32_t id = hash_min; id <= hash_max; id++) all_hashed += hashed(id);

  uint64_t made = 0;
  for (uint32_t h_first = hash_min; h_first < hash_max + start_slack; ) {
    //  grow the hash block read by read until it holds hash_block_len bases (+ terminators) or the range ends
    uint32_t h_last = h_first - 1, n_hashed = 0;
    uint64_t h_size = 0;
    while (true) {
      h_last++;
      const uint64_t add = hashed(h_last);
      h_size += add; n_hashed += add != 0;
      if (h_size >= hash_block_len || h_last >= hash_max) break;
    }
    //  cross it with ref blocks; a ref block never reaches past the hash block's last read (refID < hashID)
    uint32_t r_last = strict_reference ? 0u : ref_min - 1;
    for (uint32_t r_first = ref_min; r_first < ref_max + start_slack && r_first < h_last; r_first = r_last + 1) {
      uint64_t r_size = 0;
      while (true) {
        r_last++;
        r_size += usable(r_last);
        if (r_size >= ref_block_len || r_last >= ref_max) break;
      }
      r_last = std::min(r_last, std::min(ref_max, h_last));
      if (out && made < out_cap) {
        ovlb_tile &t = out[made];
        t.hash_bgn = h_first; t.hash_end = h_last; t.ref_bgn = r_first; t.ref_end = r_last;
        t.hash_bases = h_size; t.ref_bases = 0;
        for (uint32_t id = r_first; id <= r_last; id++) t.ref_bases += usable(id);     // bases actually inside the clamped block
        //  cost model: index build ~ hash bases; lookup ~ 2 orientations of the ref bases; extension ~ ref bases
        //  x the share of all hashable reads that sit in this hash block (= share of each read's overlaps found here)
        t.cost = (double)h_size + (double)t.ref_bases * (2.0 + 8.0 * (double)h_size / (double)(all_hashed ? all_hashed : 1));
        t.has_hash_reads = n_hashed != 0;
      }
      made++;
    }
    h_first = h_last + 1;
  }
  *n_out = made;
  if (out && made > out_cap) { ovl_set_error("ovlb_plan_tiles: output buffer too small"); return OVLB_ERR_CAPACITY; }
  return OVLB_OK;
}

//  Hash-block size from a context's memory budget (the executable's re-blocking; include/ovlb200.h states the model).
//  The run-side terms follow the allocations they stand for: run buffers mem_budget / 12 (ovl_seed_ref_batch), extension
//  scratch per warp = from-code arena (sum over rows of ceil((2e + 1) / 32) words of 8 bytes) + HBM rings + per-row arrays
//  for 148 x 32 warps, at most mem_budget / 4 (ovl_prepare_ext_scratch).
uint64_t ovlb_hash_block_bases(uint64_t budget_bytes, uint32_t max_read_len, double max_erate, uint64_t ref_batch_bases) {
  const double em = max_erate * (double)max_read_len + 64;              // rows of the longest extension
  const double per_warp = (em * em / 32 + em) * 8 + em * 64 + 4096;
  uint64_t run_side = budget_bytes / 12;
  run_side += (uint64_t)std::min<double>((double)budget_bytes / 4, 1.3 * per_warp * 148 * 32);
  run_side += 2 * 3 * ref_batch_bases + (3ull << 30);                   // two ref slots (dp4 both strands, groups, hit words), records, slack
  const uint64_t block_side = budget_bytes > run_side ? (uint64_t)((double)(budget_bytes - run_side) * 0.9) : 0;
  return std::min<uint64_t>(1500000000ull, std::max<uint64_t>(block_side / 150, 1000000ull));
}

//  Cost-balanced cut of one hash block's ref range ("when #hash blocks < #GPUs replicate the hash block on
//  every GPU and split the REF RANGE N ways").  Only refID < hashID pairs are computed (Find_Overlaps.C:279,320), so a ref
//  read meets only the hash reads behind it: the work of ref read r is ~ len_r x (hash bases with ID > r), a triangle,
//  and equal-BASE ref blocks (what partitionLength() cuts) are unequal work.  This cuts [ref_bgn, min(ref_end,
//  hash_end - 1)] into n_parts contiguous tiles of equal estimated work, so that each GPU runs ONE launch per hash block
//  (an extension launch cannot end before its slowest pair: few large launches beat many small ones).
int ovlb_plan_balanced(const uint32_t *read_len, uint32_t n_reads, uint32_t min_olap_len,
                       uint32_t hash_bgn, uint32_t hash_end, uint32_t ref_bgn, uint32_t ref_end,
                       uint32_t n_parts, double lookup_weight, ovlb_tile *out, uint64_t out_cap, uint64_t *n_out) {
  if (!read_len || !n_out || n_parts == 0 || !(lookup_weight >= 0.0)) { ovl_set_error("ovlb_plan_balanced: bad argument"); return OVLB_ERR_ARG; }
  if (hash_bgn < 1) hash_bgn = 1;
  if (ref_bgn < 1) ref_bgn = 1;
  if (hash_end > n_reads) hash_end = n_reads;
  if (ref_end > n_reads) ref_end = n_reads;
  if (hash_end > 0 && ref_end > hash_end - 1) ref_end = hash_end - 1;
  *n_out = 0;
  if (hash_end < hash_bgn || ref_end < ref_bgn) return OVLB_OK;
  auto usable = [&](uint32_t id) -> uint64_t { return read_len[id] >= min_olap_len ? read_len[id] : 0; };
  uint64_t hashBases = 0;
  for (uint32_t id = hash_bgn; id <= hash_end; id++) hashBases += usable(id) ? usable(id) + 1 : 0;
  //  behind[r] = hash bases with ID > r, accumulated from the top
  std::vector<double> w(ref_end - ref_bgn + 1);
  {
    uint64_t behind = 0;
    uint32_t h = hash_end;
    for (uint32_t r = ref_end; ; r--) {
      while (h > r && h >= hash_bgn) { behind += usable(h); h--; }
      //  lookup ~ the read's bases (both orientations); extension ~ read bases x share of the hash bases it can pair with
      const double share = hashBases ? (double)behind / (double)hashBases : 0.0;
      w[r - ref_bgn] = (double)usable(r) * (lookup_weight + share);
      if (r == ref_bgn) break;
    }
  }
  double total = 0; for (double x : w) total += x;
  uint64_t n = 0;
  uint32_t beg = ref_bgn;
  double acc = 0;
  for (uint32_t part = 0; part < n_parts && beg <= ref_end; part++) {
    const double target = total * (double)(part + 1) / (double)n_parts;
    uint32_t end = beg;
    acc += w[end - ref_bgn];
    while (end < ref_end && (part + 1 == n_parts || acc + 0.5 * w[end + 1 - ref_bgn] < target)) { end++; acc += w[end - ref_bgn]; }
    if (out) {
      if (n >= out_cap) { ovl_set_error("ovlb_plan_balanced: output buffer too small"); return OVLB_ERR_CAPACITY; }
      ovlb_tile &t = out[n];
      t.hash_bgn = hash_bgn; t.hash_end = hash_end; t.ref_bgn = beg; t.ref_end = end;
      t.hash_bases = hashBases; t.ref_bases = 0; t.cost = 0;
      for (uint32_t id = beg; id <= end; id++) { t.ref_bases += usable(id); t.cost += w[id - ref_bgn]; }
      t.has_hash_reads = hashBases != 0;
    }
    n++;
    beg = end + 1;
  }
  *n_out = n;
  return OVLB_OK;
}

//  Longest-processing-time-first assignment of tiles to workers: tiles sorted by cost
//  descending, each given to the currently least-loaded worker.  Deterministic (ties: lower index first).
int ovlb_assign_tiles(const ovlb_tile *tiles, uint64_t n_tiles, uint32_t n_workers, uint32_t *owner) {
  if ((n_tiles && (!tiles || !owner)) || n_workers == 0) { ovl_set_error("ovlb_assign_tiles: bad argument"); return OVLB_ERR_ARG; }
  std::vector<uint64_t> order(n_tiles);
  for (uint64_t i = 0; i < n_tiles; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return tiles[a].cost > tiles[b].cost; });
  std::vector<double> load(n_workers, 0.0);
  for (uint64_t k = 0; k < n_tiles; k++) {
    uint32_t best = 0;
    for (uint32_t w = 1; w < n_workers; w++) if (load[w] < load[best]) best = w;
    owner[order[k]] = best;
    load[best] += tiles[order[k]].cost;
  }
  return OVLB_OK;
}

}  // extern "C"
