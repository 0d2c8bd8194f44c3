//  ovltool.cc -- small companion tool for the host-side formats (used by the CPU test-suite):
//    ovltool dump-store <seqStore> [--packed]    reads as the overlapper sees them, FASTA on stdout
//                                                (--packed: through the zero-decode 2-bit path)
//    ovltool dump-ovb <file.ovb>                 records as "a b dat0 dat1" (hex words) on stdout
//    ovltool rewrite-ovb <in.ovb> <out.ovb> <lastReadID>   decode with our snappy reader, write with our writer
//    ovltool cmp-ovb <a.ovb> <b.ovb> [ignoreReadID]        canonical sort + compare, exit 0 iff identical
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ovfile.h"
#include "ovstore.h"
#include "prefetch.h"
#include "sqstore.h"

//  ovltool does not link the CUDA library: nothing it packs is ever page-locked, so Packed never calls these
extern "C" int ovlb_host_register(const void *, size_t) { return 1; }
extern "C" int ovlb_host_unregister(const void *) { return 0; }

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
    if (!snappy_uncompress(comp.data(), cl, raw)) { err = "bad snappy block"; fclose(f); return false; }
    if (raw.size() % 24) { err = "block is not a whole number of records"; fclose(f); return false; }
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
  std::string err;
  if (argc >= 3 && !strcmp(argv[1], "dump-store")) {
    const bool packed = argc >= 4 && !strcmp(argv[3], "--packed");
    SqStore S;
    if (!S.open(argv[2], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::string b;
    for (uint32_t id = 1; id <= S.lastReadID(); id++) {
      const uint32_t L = S.readLength(id);
      if (L == 0) continue;
      bool done = false;
      if (packed) {
        std::vector<uint8_t> pk;
        int r = S.appendPacked2bit(id, pk, err);
        if (r < 0) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        if (r == 1) {
          b.resize(L);
          for (uint32_t j = 0; j < L; j++) b[j] = "ACGT"[(pk[j >> 2] >> (6 - 2 * (j & 3))) & 3];
          done = true;
        }
      }
      if (!done && !S.loadRead(id, b, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
      printf(">read%u len=%u lib=%u\n%s\n", id, L, S.libraryID(id), b.c_str());
    }
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "dump-ovb")) {
    std::vector<ovlb_record> recs;
    if (!read_ovb(argv[2], recs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    for (auto &r : recs) printf("%u %u %016lx %016lx\n", r.a_iid, r.b_iid, (unsigned long)r.dat0, (unsigned long)r.dat1);
    return 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "rewrite-ovb")) {
    std::vector<ovlb_record> recs;
    if (!read_ovb(argv[2], recs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    OvFileWriter W;
    if (!W.open(argv[3], (uint32_t)strtoul(argv[4], nullptr, 10), err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    //  hand the records over in uneven batches to exercise block boundaries
    size_t p = 0, step = 1000;
    while (p < recs.size()) {
      size_t n = std::min(step, recs.size() - p);
      W.submit(std::vector<ovlb_record>(recs.begin() + p, recs.begin() + p + n));
      p += n; step = step * 3 + 7;
    }
    if (!W.close(err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    return 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "pack-ovb")) {                     // flat {u32 a, u32 b, u64 dat0, u64 dat1} records -> .ovb + .oc
    FILE *f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    OvFileWriter W;
    if (!W.open(argv[3], (uint32_t)strtoul(argv[4], nullptr, 10), err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::vector<ovlb_record> buf(1 << 20);
    size_t got;
    while ((got = fread(buf.data(), sizeof(ovlb_record), buf.size(), f)) > 0)
      W.submit(std::vector<ovlb_record>(buf.begin(), buf.begin() + got));
    fclose(f);
    if (!W.close(err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "cmp-ovb")) {
    //  canonical-sort both files (ovOverlap::operator<, stores/ovOverlap.H:263-276) and compare record by record;
    //  optional 4th argument: a read ID whose records are ignored on both sides (the reference's last-ref-read quirk)
    std::vector<ovlb_record> A, B;
    if (!read_ovb(argv[2], A, err) || !read_ovb(argv[3], B, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    const uint32_t ignore = argc >= 5 ? (uint32_t)strtoul(argv[4], nullptr, 10) : 0;
    auto drop = [&](std::vector<ovlb_record> &v) {
      if (!ignore) return;
      size_t k = 0;
      for (auto &r : v) if (r.a_iid != ignore && r.b_iid != ignore) v[k++] = r;
      v.resize(k);
    };
    drop(A); drop(B);
    auto lt = [](const ovlb_record &x, const ovlb_record &y) {
      if (x.a_iid != y.a_iid) return x.a_iid < y.a_iid;
      if (x.b_iid != y.b_iid) return x.b_iid < y.b_iid;
      if (x.dat0 != y.dat0) return x.dat0 < y.dat0;
      return x.dat1 < y.dat1;
    };
    std::sort(A.begin(), A.end(), lt); std::sort(B.begin(), B.end(), lt);
    size_t i = 0, j = 0, onlyA = 0, onlyB = 0, shown = 0;
    while (i < A.size() || j < B.size()) {
      if (j == B.size() || (i < A.size() && lt(A[i], B[j]))) {
        if (shown++ < 10) printf("only-first  %u %u %016lx %016lx\n", A[i].a_iid, A[i].b_iid, (unsigned long)A[i].dat0, (unsigned long)A[i].dat1);
        onlyA++; i++;
      } else if (i == A.size() || lt(B[j], A[i])) {
        if (shown++ < 10) printf("only-second %u %u %016lx %016lx\n", B[j].a_iid, B[j].b_iid, (unsigned long)B[j].dat0, (unsigned long)B[j].dat1);
        onlyB++; j++;
      } else { i++; j++; }
    }
    printf("first %zu second %zu only-first %zu only-second %zu\n", A.size(), B.size(), onlyA, onlyB);
    return (onlyA || onlyB) ? 1 : 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "write-store")) {                  // flat SORTED {u32 a, u32 b, u64 dat0, u64 dat1} records -> ovStore directory
    FILE *f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    std::vector<ovlb_record> recs; std::vector<ovlb_record> buf(1 << 20);
    size_t got;
    while ((got = fread(buf.data(), sizeof(ovlb_record), buf.size(), f)) > 0) recs.insert(recs.end(), buf.begin(), buf.begin() + got);
    fclose(f);
    if (!write_ovstore(argv[3], (uint32_t)strtoul(argv[4], nullptr, 10), recs.data(), recs.size(), err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "lengths")) {                      // read lengths as the overlapper sees them, one per line (ID order)
    SqStore S;
    if (!S.open(argv[2], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    for (uint32_t id = 1; id <= S.lastReadID(); id++) printf("%u\n", S.readLength(id));
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "hash-ovb")) {                     // order-independent digest of a (large) .ovb: count + 2 x 64-bit sums of per-record hashes
    std::vector<ovlb_record> recs;
    if (!read_ovb(argv[2], recs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    uint64_t s1 = 0, s2 = 0;
    auto mix = [](uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; };
    for (auto &r : recs) {
      const uint64_t h = mix(mix(((uint64_t)r.a_iid << 32) | r.b_iid) ^ mix(r.dat0 + 0x9E3779B97F4A7C15ull) ^ mix(r.dat1 * 3 + 1));
      s1 += h; s2 += mix(h);
    }
    printf("%zu:%016lx%016lx\n", recs.size(), (unsigned long)s1, (unsigned long)s2);
    return 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "prefetch-check")) {               // <seqStore> <pieces> <threads>: the multi-threaded prefetcher hands out the same batches, in plan order, as one thread
    SqStore S;
    if (!S.open(argv[2], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const uint32_t N = S.lastReadID(), pieces = std::max(1u, (uint32_t)strtoul(argv[3], nullptr, 10)), threads = std::max(1u, (uint32_t)strtoul(argv[4], nullptr, 10));
    std::vector<PackItem> plan;
    plan.push_back({true, 1, std::max(1u, N / 3)});                    // a hash block, then ref pieces of uneven size, an empty range, another block
    for (uint32_t k = 0, b = 1; k < pieces && b <= N; k++) {
      const uint32_t e = std::min(N, b + (N / pieces) * (1 + k % 3) / 2);
      plan.push_back({false, b, e}); b = e + 1;
      if (k == pieces / 2) { plan.push_back({false, 5, 4}); plan.push_back({true, N / 2 + 1, N}); }
    }
    auto digest = [](const Packed &P) {
      uint64_t h = 1469598103934665603ull;
      auto eat = [&h](const void *p, size_t n) { const uint8_t *b = (const uint8_t *)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
      eat(P.packed.data(), P.packed.size()); eat(P.boff.data(), P.boff.size() * 8); eat(P.len.data(), P.len.size() * 4);
      eat(P.n_read.data(), P.n_read.size() * 4); eat(P.n_pos.data(), P.n_pos.size() * 4);
      eat(P.src_len.data(), P.src_len.size() * 4); eat(P.clear_bgn.data(), P.clear_bgn.size() * 4); eat(&P.bases, 8);
      eat(&P.view.n_reads, 4); eat(&P.view.first_read_id, 4);
      return h;
    };
    auto run = [&](unsigned nt, std::vector<uint64_t> &out) -> bool {
      Prefetcher pf(argv[2], plan, 0, 0xFFFFFFFFu, 0, 0xFFFFFFFFu, 500, nt > 1 ? 4 : 3, nt);
      for (size_t i = 0; i < plan.size(); i++) {
        std::unique_ptr<Packed> p = pf.next(err);
        if (!p) return false;
        if (p->view.n_reads && p->view.first_read_id != plan[i].bgn) { err = "batch out of plan order"; return false; }
        out.push_back(digest(*p));
        pf.recycle(std::move(p));                                        // recycled buffers must not leak a previous batch's contents
      }
      std::string e2;
      if (pf.next(e2)) { err = "prefetcher handed out more than the plan"; return false; }
      return true;
    };
    std::vector<uint64_t> one, many;
    if (!run(1, one) || !run(threads, many)) { fprintf(stderr, "prefetch-check: %s\n", err.c_str()); return 1; }
    uint64_t all = 0;
    for (size_t i = 0; i < one.size(); i++) {
      if (one[i] != many[i]) { fprintf(stderr, "prefetch-check: item %zu differs between 1 and %u threads\n", i, threads); return 1; }
      all = all * 31 + one[i];
    }
    printf("prefetch-check ok: %zu items, %u threads, digest %016lx\n", one.size(), threads, (unsigned long)all);
    return 0;
  }
  fprintf(stderr, "usage: ovltool write-store <sorted.bin> <out.ovlStore> <lastReadID> | lengths <seqStore> | hash-ovb <file.ovb> | dump-store <seqStore> [--packed] | dump-ovb <file.ovb> | rewrite-ovb <in.ovb> <out.ovb> <lastReadID> | pack-ovb <in.bin> <out.ovb> <lastReadID> | cmp-ovb <a.ovb> <b.ovb> [ignoreReadID] | prefetch-check <seqStore> <pieces> <threads>\n");
  return 1;
}
