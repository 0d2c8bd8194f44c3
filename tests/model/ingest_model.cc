//  TEST HARNESS: compiles the product's host+device ingest logic (canu_b200/csrc/ovl_common.cuh, ovl_ingest_twin) for the
//  CPU and applies it the way ovl_ingest.cu does (twin + filter per record, then a full-key sort), so that the code the
//  kernel runs is checked against the reference-minted goldens without a GPU.
#include "../../canu_b200/csrc/ovl_common.cuh"

#include <algorithm>
#include <cstdint>
#include <vector>

struct Rec { uint32_t a, b; uint64_t w0, w1; };

extern "C" int64_t ingest_model(const Rec *in, int64_t n, uint32_t max_evalue, Rec *out) {
  std::vector<Rec> v;
  v.reserve(2 * n);
  for (int64_t i = 0; i < n; i++) {
    Rec f = in[i], r;
    r.a = f.b; r.b = f.a;
    if (ovl_ingest_twin(in[i].w0, in[i].w1, max_evalue, &f.w0, &r.w0, &r.w1)) { v.push_back(f); v.push_back(r); }
  }
  std::sort(v.begin(), v.end(), [](const Rec &x, const Rec &y) {
    if (x.a != y.a) return x.a < y.a;
    if (x.b != y.b) return x.b < y.b;
    if (x.w0 != y.w0) return x.w0 < y.w0;
    return x.w1 < y.w1;
  });
  std::copy(v.begin(), v.end(), out);
  return (int64_t)v.size();
}
