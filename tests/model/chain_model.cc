//  TEST INFRASTRUCTURE.  Host build (g++) of the product's run-level seed-list construction
//  (canu_b200/csrc/ovl_common.cuh: ovl_chain_simulate) fed by a straightforward CPU enumeration of
//  exact k-mer hits, so the kernel-side algorithm can be checked against the oracle's
//  Match_Node lists without a GPU.  Not shipped, not linked into the product library.
#include "../../canu_b200/csrc/ovl_common.cuh"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

struct PairOut { uint32_t ref_id, hash_id; int32_t dir, consistent, diag_ct, diag_bgn, diag_end, n_seeds; int64_t seed_begin; };
struct SeedOut { int32_t start, offset, len; };

static std::vector<PairOut> g_pairs;
static std::vector<SeedOut> g_seeds;

static inline int code_of(char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
  return -1;
}
static inline char comp(char c) {
  switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; }
  return 'N';
}

extern "C" {

//  reads: ASCII, ids = index+1.  Hash range [hb,he], ref range [rb,re].  Returns number of pairs.
int64_t chain_model_run(const char *bases, const uint64_t *offs, const uint32_t *lens, uint32_t n, int K, int min_len,
                        uint32_t hb, uint32_t he, uint32_t rb, uint32_t re) {
  g_pairs.clear(); g_seeds.clear();
  std::unordered_map<uint64_t, std::vector<uint64_t>> index;   // kmer -> (read<<32 | pos), insertion order
  const uint64_t mask = (1ull << (2 * K)) - 1;
  for (uint32_t id = hb; id <= he; id++) {
    const uint32_t L = lens[id - 1];
    if ((int)L < min_len || (int)L < K) continue;
    const char *s = bases + offs[id - 1];
    uint64_t key = 0; int valid = 0;
    for (uint32_t i = 0; i < L; i++) {
      int c = code_of(s[i]);
      key = (key >> 2);
      if (c < 0) { valid = 0; } else { key |= (uint64_t)c << (2 * (K - 1)); valid++; }
      if (c < 0) key = 0;
      if (valid >= K) index[key & mask].push_back(((uint64_t)id << 32) | (i - K + 1));
    }
  }
  std::vector<char> buf;
  for (uint32_t id = rb; id <= re; id++) {
    const uint32_t L = lens[id - 1];
    if ((int)L < min_len || (int)L < K) continue;
    for (int dir = 0; dir < 2; dir++) {
      buf.assign(bases + offs[id - 1], bases + offs[id - 1] + L);
      if (dir) { std::reverse(buf.begin(), buf.end()); for (auto &c : buf) c = comp(c); }
      //  hits per hash read: (ref pos, hash pos)
      std::map<uint32_t, std::vector<std::pair<int, int>>> hits;
      uint64_t key = 0; int valid = 0;
      for (uint32_t i = 0; i < L; i++) {
        int c = code_of(buf[i]);
        key = (key >> 2);
        if (c < 0) { valid = 0; key = 0; } else { key |= (uint64_t)c << (2 * (K - 1)); valid++; }
        if (valid >= K) {
          auto it = index.find(key & mask);
          if (it != index.end())
            for (uint64_t e : it->second) {
              uint32_t h = (uint32_t)(e >> 32);
              if (id < h) hits[h].push_back({(int)(i - K + 1), (int)(uint32_t)e});
            }
        }
      }
      for (auto &kv : hits) {
        auto &v = kv.second;
        //  maximal diagonal runs
        std::sort(v.begin(), v.end(), [](const std::pair<int,int>&a, const std::pair<int,int>&b){
          int da = a.second - a.first, db = b.second - b.first; if (da != db) return da < db; return a.first < b.first; });
        std::vector<OvlRun> runs;
        for (size_t i = 0; i < v.size(); ) {
          size_t j = i + 1;
          while (j < v.size() && v[j].second - v[j].first == v[i].second - v[i].first && v[j].first == v[j - 1].first + 1) j++;
          runs.push_back({v[i].first, v[i].second, (int)(j - i)});
          i = j;
        }
        std::sort(runs.begin(), runs.end(), [](const OvlRun &a, const OvlRun &b){ if (a.start != b.start) return a.start < b.start; return a.q < b.q; });
        int nr = (int)runs.size();
        std::vector<int32_t> nxt(nr), hts(nr), act(nr), order(nr);
        int consistent = 1;
        if (nr == 1) order[0] = 0; else consistent = ovl_chain_simulate(runs.data(), nr, K, nxt.data(), hts.data(), act.data(), order.data());
        PairOut p; p.ref_id = id; p.hash_id = kv.first; p.dir = dir; p.consistent = consistent;
        p.diag_ct = 0; p.diag_end = 0; p.diag_bgn = runs[0].start;
        for (auto &r : runs) { p.diag_ct += r.len; p.diag_end = std::max(p.diag_end, r.start + r.len - 1); }
        p.n_seeds = nr; p.seed_begin = (int64_t)g_seeds.size();
        for (int k = 0; k < nr; k++) { const OvlRun &r = runs[order[k]]; g_seeds.push_back({r.start, r.q, K + r.len - 1}); }
        g_pairs.push_back(p);
      }
    }
  }
  return (int64_t)g_pairs.size();
}

const PairOut *chain_model_pairs() { return g_pairs.data(); }
const SeedOut *chain_model_seeds() { return g_seeds.data(); }
int64_t chain_model_num_seeds() { return (int64_t)g_seeds.size(); }

}
