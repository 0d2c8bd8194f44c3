//  storebuild.cc -- `ovlStoreBuild`: overlapper output (.ovb files) -> Canu overlap store, the step right after the ovl
//  path.  Replaces the sequential `ovStoreBuild` (stores/ovStoreBuild.C:176-263: load every
//  overlap, ovStoreFilter::filterOverlap -> mirrored twin + error-rate filter, std::sort, ovStoreWriter) -- and, for a
//  job that fits the device, the bucketizer / sorter / indexer trio of the parallel build -- by one read of the .ovb
//  files, ovlb_ingest_records on the GPU (mirror + filter + sort), and a streaming write of the store files.
//
//      ovlStoreBuild -O asm.ovlStore -S asm.seqStore [-e maxErate] [--gpu n] a.ovb b.ovb ...
//
//  The store is readable by the reference's ovStoreDump / ovStore class (parity: tests/test_store_build.py).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "../../include/ovlb200.h"
#include "ovfile.h"
#include "ovstore.h"
#include "sqstore.h"

using namespace ovlhost;

static bool read_ovb(const char *fn, std::vector<ovlb_record> &recs, std::string &err) {
  FILE *f = fopen(fn, "rb");
  if (!f) { err = std::string("cannot open ") + fn; return false; }
  std::vector<uint8_t> comp, raw;
  while (true) {
    uint64_t cl;
    if (fread(&cl, 8, 1, f) != 1) break;
    comp.resize(cl);
    if (fread(comp.data(), 1, cl, f) != cl) { err = "short read"; fclose(f); return false; }
    if (!snappy_uncompress(comp.data(), cl, raw) || raw.size() % 24) { err = std::string("bad block in ") + fn; fclose(f); return false; }
    for (size_t p = 0; p < raw.size(); p += 24) {
      uint32_t w[6]; memcpy(w, &raw[p], 24);
      ovlb_record r; r.a_iid = w[0]; r.b_iid = w[1];
      r.dat0 = ((uint64_t)w[2] << 32) | w[3]; r.dat1 = ((uint64_t)w[4] << 32) | w[5];
      recs.push_back(r);
    }
  }
  fclose(f);
  return true;
}

int main(int argc, char **argv) {
  const char *outp = nullptr, *seqp = nullptr; double erate = 1.0; int gpu = 0;
  std::vector<const char *> inputs;
  int err = 0;
  for (int a = 1; a < argc; a++) {
    auto need = [&](const char *f) -> const char * { if (a + 1 >= argc) { fprintf(stderr, "option %s needs a value\n", f); err++; return "0"; } return argv[++a]; };
    if      (!strcmp(argv[a], "-O"))    outp = need("-O");
    else if (!strcmp(argv[a], "-S"))    seqp = need("-S");
    else if (!strcmp(argv[a], "-e"))    erate = strtod(need("-e"), nullptr);
    else if (!strcmp(argv[a], "--gpu")) gpu = atoi(need("--gpu"));
    else if (argv[a][0] == '-' && argv[a][1]) { fprintf(stderr, "Unknown option '%s'\n", argv[a]); err++; }
    else inputs.push_back(argv[a]);
  }
  if (err || !outp || !seqp || inputs.empty()) {
    fprintf(stderr, "usage: %s -O asm.ovlStore -S asm.seqStore [-e maxErate] [--gpu n] *.ovb\n", argv[0]);
    return 1;
  }
  std::string e;
  SqStore S;
  if (!S.open(seqp, e)) { fprintf(stderr, "sqStore()--  failed to open '%s' for read-only access: %s.\n", seqp, e.c_str()); return 1; }
  const uint32_t maxID = S.lastReadID();
  std::vector<ovlb_record> in;
  for (const char *fn : inputs) if (!read_ovb(fn, in, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }
  ovlb_params P;
  if (ovlb_params_init(&P, 22, 0.06, 1.0, 0, 1, 0, 0, 0, 1024)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  ovlb_ctx *ctx = nullptr;
  if (ovlb_create(gpu, &P, &ctx)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  const uint32_t maxEvalue = erate < 65535 / 100000.0 ? (uint32_t)(100000.0 * erate + 0.5) : 65535u;   // AS_OVS_encodeEvalue (ovOverlap.H:31-35)
  std::vector<ovlb_record> out(2 * in.size() + 1);
  uint64_t n = 0;
  if (ovlb_ingest_records(ctx, in.data(), in.size(), maxEvalue, maxID, out.data(), out.size(), &n)) { fprintf(stderr, "ERROR: %s\n", ovlb_last_error()); return 1; }
  if (!write_ovstore(outp, maxID, out.data(), n, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }
  fprintf(stderr, "Created ovStore '%s' with %lu overlaps (from %lu in %zu file(s)) for reads up to %u.\n",
          outp, (unsigned long)n, (unsigned long)in.size(), inputs.size(), maxID);
  ovlb_params_free(&P);
  fflush(stdout); fflush(stderr);
  _exit(0);
}
