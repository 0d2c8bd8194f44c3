//  freqmers.cc -- `ovlFrequentMers`: the frequent-k-mer list Canu passes to `overlapInCore -k <file>`, computed on the GPU.
//
//  Replaces, for the ovl overlapper, the three meryl steps of the pipeline (src/pipelines/canu/Meryl.pm:529-533 `meryl
//  count`, :603-607 `greater-than 1 ... union-sum`, :663-671 `print at-least distinct=D at-least threshold=T`): one pass
//  over the sqStore, canonical k-mer counting on the device (ovlb_kmer_census), and the same text file on the way out --
//  one "kmer<TAB>count" line per frequent k-mer (overlapInCore reads the first word of every line and adds both
//  orientations, overlapInCore-Build_Hash_Index.C:186-257).  Line order is unspecified (meryl's own order depends on its
//  file prefix bits).  No CPU fallback.
//
//      ovlFrequentMers -k 22 [-distinct 0.9999] [-threshold N] [-slices S] [--gpu n] -o asm.ms22.dump <seqStore>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "../../include/ovlb200.h"
#include "pack.h"
#include "sqstore.h"

using namespace ovlhost;

int main(int argc, char **argv) {
  uint32_t K = 0; double distinct = -1.0; uint64_t threshold = 0; int gpu = 0; uint32_t slices = 0;
  const char *out = nullptr, *store = nullptr;
  int err = 0;
  for (int a = 1; a < argc; a++) {
    auto need = [&](const char *f) -> const char * { if (a + 1 >= argc) { fprintf(stderr, "option %s needs a value\n", f); err++; return "0"; } return argv[++a]; };
    if      (!strcmp(argv[a], "-k"))         K = (uint32_t)strtoul(need("-k"), nullptr, 10);
    else if (!strcmp(argv[a], "-distinct"))  distinct = strtod(need("-distinct"), nullptr);
    else if (!strcmp(argv[a], "-threshold")) threshold = strtoull(need("-threshold"), nullptr, 10);
    else if (!strcmp(argv[a], "-slices"))    slices = (uint32_t)strtoul(need("-slices"), nullptr, 10);
    else if (!strcmp(argv[a], "--gpu"))      gpu = atoi(need("--gpu"));
    else if (!strcmp(argv[a], "-o"))         out = need("-o");
    else if (argv[a][0] == '-' && argv[a][1]) { fprintf(stderr, "Unknown option '%s'\n", argv[a]); err++; }
    else if (!store)                         store = argv[a];
    else { fprintf(stderr, "Unknown option '%s'\n", argv[a]); err++; }
  }
  if (K < 2 || K > 30) { fprintf(stderr, "-k must be in 2..30\n"); err++; }
  if (distinct > 1.0) { fprintf(stderr, "-distinct must be a fraction in 0..1\n"); err++; }
  if (distinct < 0 && threshold == 0) { fprintf(stderr, "one of -distinct / -threshold is needed\n"); err++; }
  if (err || !out || !store) {
    fprintf(stderr, "usage: %s -k K [-distinct D] [-threshold T] [-slices S] [--gpu n] -o out.dump <seqStore>\n", argv[0]);
    return 1;
  }
  std::string e;
  SqStore S;
  if (!S.open(store, e)) { fprintf(stderr, "sqStore()--  failed to open '%s' for read-only access: %s.\n", store, e.c_str()); return 1; }
  const uint32_t N = S.lastReadID();
  uint32_t maxLen = 64;
  for (uint32_t id = 1; id <= N; id++) maxLen = std::max(maxLen, S.readLength(id));
  ovlb_params P;
  if (ovlb_params_init(&P, K, 0.06, 1.0, 0, 1, 0, 0, 0, maxLen)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  ovlb_ctx *ctx = nullptr;
  if (ovlb_create(gpu, &P, &ctx)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  Packed R;
  if (!pack_range(S, 1, N, 0, UINT32_MAX, K, R, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }
  if (ovlb_load_hash_reads(ctx, &R.view)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  //  slices: scratch is ~30 bytes per base and slice; keep it under a third of the device
  uint32_t sb = 0;
  if (slices) { while ((1u << sb) < slices) sb++; }
  else { uint64_t fr = 0, tot = 0; ovlb_device_memory(gpu, &fr, &tot); while (sb < 8 && (R.bases >> sb) * 30 > fr / 3) sb++; }
  std::vector<uint64_t> keys(1 << 20); std::vector<uint32_t> cnts(1 << 20);
  uint64_t n = 0, st[4] = {0, 0, 0, 0};
  int rc;
  while ((rc = ovlb_kmer_census(ctx, sb, distinct, threshold, keys.data(), cnts.data(), keys.size(), &n, st)) == OVLB_ERR_CAPACITY && n >= keys.size()) {
    keys.resize(keys.size() * 4); cnts.resize(keys.size());
  }
  if (rc) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  FILE *F = fopen(out, "w");
  if (!F) { fprintf(stderr, "ERROR: cannot write '%s'\n", out); return 1; }
  std::string km(K, 'A');
  for (uint64_t i = 0; i < n; i++) {
    for (uint32_t j = 0; j < K; j++) km[j] = "ACGT"[(keys[i] >> (2 * j)) & 3];
    fprintf(F, "%s\t%u\n", km.c_str(), cnts[i]);
  }
  fclose(F);
  fprintf(stderr, "ovlFrequentMers: %u reads, %lu bases; %lu distinct %u-mers of count >= 2 (%lu occurrences), %lu singletons; threshold %lu -> %lu k-mers written to '%s'\n",
          N, (unsigned long)R.bases, (unsigned long)st[0], K, (unsigned long)st[1], (unsigned long)st[2], (unsigned long)st[3], (unsigned long)n, out);
  ovlb_params_free(&P);
  fflush(stdout); fflush(stderr);
  _exit(0);
}
