//  main.cc -- `overlapInCore`, B200 edition: a drop-in for Canu's ovl overlapper executable.
//
//  Same command line, same sqStore input, same .ovb / .oc / .stats outputs as the reference
//  (overlapInCore.C:284-565), so the `overlap.sh` script Canu generates
//  (src/pipelines/canu/OverlapInCore.pm:175-194) can call it unchanged.  The compute path is the
//  CUDA library behind include/ovlb200.h; there is no CPU fallback.
//
//  Flags that only size the reference's CPU data structures are accepted and ignored because the
//  output does not depend on them: --hashbits, --hashload, --hashdatalen, -t.
//  Extra flags of this build: --gpu N (device, default 0), --refbatch BASES, --hashblock BASES.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>

#include "../../include/ovlb200.h"
#include "ovfile.h"
#include "sqstore.h"

using namespace ovlhost;

struct Options {
  bool     partial = false;
  uint32_t bgnHashID = 1, endHashID = UINT32_MAX, minLibToHash = 0, maxLibToHash = UINT32_MAX;
  uint32_t bgnRefID = 1, endRefID = UINT32_MAX, minLibToRef = 0, maxLibToRef = UINT32_MAX;
  uint64_t kmerLen = 0;
  const char *kmerSkipFileName = nullptr;
  bool     unique = true;
  const char *outName = nullptr, *statName = nullptr, *storePath = nullptr;
  int32_t  minOlapLen = 0;
  bool     minKmers = false, noHopeless = false;
  double   maxErate = 0.06, alignNoise = 1.0;
  int      gpu = 0;
  uint64_t refBatchBases = 0, hashBlockBases = 0;
};

static bool file_exists(const char *p) { struct stat st; return stat(p, &st) == 0 && S_ISREG(st.st_mode); }

//  decodeRange (utility/src/datastructures/types-v1.C): "a-b" or "a"
static void decode_range(const char *s, uint32_t &lo, uint32_t &hi) {
  char *end = nullptr;
  unsigned long a = strtoul(s, &end, 10);
  lo = hi = (uint32_t)a;
  if (end && *end == '-') hi = (uint32_t)strtoul(end + 1, nullptr, 10);
}

static void usage(const char *argv0) {
  fprintf(stderr, "USAGE:  %s [options] <seqStorePath>\n\n", argv0);
  fprintf(stderr, "-partial    do partial overlaps\n");
  fprintf(stderr, "-h <range>  to specify fragments to put in hash table\n");
  fprintf(stderr, "-H <range>  libraries to put in the hash table\n");
  fprintf(stderr, "-r <range>  specify old fragments to overlap\n");
  fprintf(stderr, "-R <range>  libraries to overlap\n");
  fprintf(stderr, "-k          if a number, the length of a kmer, otherwise\n");
  fprintf(stderr, "            the filename containing a list of kmers to ignore in\n");
  fprintf(stderr, "            the hash table\n");
  fprintf(stderr, "-m          allow multiple overlaps per oriented fragment pair\n");
  fprintf(stderr, "-u          allow only 1 overlap per oriented fragment pair\n");
  fprintf(stderr, "-o          specify output file name\n");
  fprintf(stderr, "-s          specify statistics file name\n");
  fprintf(stderr, "-t <n>      accepted for compatibility (the GPU build ignores it)\n");
  fprintf(stderr, "-z          skip the hopeless check (also skipped at > 0.06)\n\n");
  fprintf(stderr, "--maxerate <n>     only output overlaps with fraction <n> or less error (e.g., 0.06 == 6%%)\n");
  fprintf(stderr, "--minlength <n>    only output overlaps of <n> or more bases\n");
  fprintf(stderr, "--minkmers         filter candidate pairs by k-mer count\n\n");
  fprintf(stderr, "--hashbits n / --hashdatalen n / --hashload f   accepted and ignored (output does not depend on them)\n\n");
  fprintf(stderr, "--gpu n            CUDA device to use (default 0)\n");
  fprintf(stderr, "--refbatch n       bases of reference reads per device batch\n");
  fprintf(stderr, "--hashblock n      bases of hash reads per device-resident index\n\n");
}

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

//  Pack reads bgn..end (inclusive) of the store into the wire format.  Reads outside the library
//  filter or shorter than `min_len` keep their slot with length 0 (Build_Hash_Index.C:504-526,
//  Process_Overlaps.C:55-60).
struct Packed {
  std::vector<uint8_t> packed; std::vector<uint64_t> boff; std::vector<uint32_t> len, n_read, n_pos;
  ovlb_reads view; uint64_t bases = 0;
};

static bool pack_range(SqStore &S, uint32_t bgn, uint32_t end, uint32_t min_lib, uint32_t max_lib, uint32_t min_len,
                       Packed &P, std::string &err) {
  const uint32_t n = end >= bgn ? end - bgn + 1 : 0;
  P.packed.clear(); P.boff.assign(n, 0); P.len.assign(n, 0); P.n_read.clear(); P.n_pos.clear(); P.bases = 0;
  std::string bases;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t id = bgn + i;
    P.boff[i] = P.packed.size();
    const uint32_t lib = S.libraryID(id);
    const uint32_t L = S.readLength(id);
    if (lib < min_lib || lib > max_lib || L < min_len) continue;
    int fast = S.appendPacked2bit(id, P.packed, err);
    if (fast < 0) return false;
    if (fast == 0) {
      if (!S.loadRead(id, bases, err)) return false;
      if (bases.size() != L) { err = "read " + std::to_string(id) + ": decoded length differs from metadata"; return false; }
      const size_t b0 = P.packed.size();
      P.packed.resize(b0 + (L + 3) / 4, 0);
      for (uint32_t j = 0; j < L; j++) {
        unsigned code;
        switch (bases[j]) {
          case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
          case 'N': code = 0; P.n_read.push_back(i); P.n_pos.push_back(j); break;
          default: err = "read " + std::to_string(id) + " contains a base that is not ACGTN; this build does not support IUPAC codes"; return false;
        }
        P.packed[b0 + (j >> 2)] |= (uint8_t)(code << (6 - 2 * (j & 3)));
      }
    }
    P.len[i] = L;
    P.bases += L;
  }
  P.packed.resize(P.packed.size() + 8, 0);
  P.view.packed = P.packed.data(); P.view.packed_bytes = P.packed.size() - 8;
  P.view.byte_offset = P.boff.data(); P.view.len = P.len.data(); P.view.n_reads = n; P.view.first_read_id = bgn;
  P.view.n_read = P.n_read.data(); P.view.n_pos = P.n_pos.data(); P.view.n_n = P.n_read.size();
  return true;
}

//  Mark_Skip_Kmers' file format (Build_Hash_Index.C:186-257): one k-mer per line, first whitespace-
//  delimited word; a line starting with '>' is a header and the NEXT line holds the k-mer.
static int load_skip_kmers(const char *fn, uint32_t K, std::vector<uint64_t> &keys) {
  FILE *F = fopen(fn, "r");
  if (!F) FAIL("Failed to open skip k-mer file '%s'", fn);
  char line[1024]; int lineNum = 0, kmerNum = 0;
  while (fgets(line, 1024, F)) {
    lineNum++;
    if (line[0] == '>') { if (!fgets(line, 1024, F)) break; lineNum++; }
    for (int i = 0; line[i]; i++) if (line[i] == ' ' || line[i] == '\t' || line[i] == '\n' || line[i] == '\r') { line[i] = 0; break; }
    size_t len = strlen(line);
    if (len != K) { fprintf(stderr, "Short kmer skip kmer '%s' at line %d, expecting length %d got length %d.\n", line, lineNum, (int)K, (int)len); fclose(F); return 1; }
    uint64_t f, r;
    if (ovlb_kmer_keys(line, K, &f, &r)) { fprintf(stderr, "Bad skip kmer '%s' at line %d: %s\n", line, lineNum, ovlb_last_error()); fclose(F); return 1; }
    keys.push_back(f); keys.push_back(r);
    kmerNum++;
  }
  fclose(F);
  fprintf(stderr, "\nRead %d kmers to mark to skip\n\n", kmerNum);
  return 0;
}

int main(int argc, char **argv) {
  Options G;
  int err = 0;
  for (int arg = 1; arg < argc; arg++) {
    auto need = [&](const char *f) -> const char * { if (arg + 1 >= argc) { fprintf(stderr, "option %s needs a value\n", f); err++; return "0"; } return argv[++arg]; };
    const char *a = argv[arg];
    if      (!strcmp(a, "-partial"))      G.partial = true;
    else if (!strcmp(a, "-h"))            decode_range(need(a), G.bgnHashID, G.endHashID);
    else if (!strcmp(a, "-H"))            decode_range(need(a), G.minLibToHash, G.maxLibToHash);
    else if (!strcmp(a, "-r"))            decode_range(need(a), G.bgnRefID, G.endRefID);
    else if (!strcmp(a, "-R"))            decode_range(need(a), G.minLibToRef, G.maxLibToRef);
    else if (!strcmp(a, "-k")) {
      const char *v = need(a);
      if (file_exists(v)) G.kmerSkipFileName = v;                      // overlapInCore.C:309-315
      else                G.kmerLen = strtoul(v, nullptr, 10);
    }
    else if (!strcmp(a, "-l")) {
      long v = strtol(need(a), nullptr, 10);
      if (v >= 1) { fprintf(stderr, "ERROR: -l (per-read overlap limit) is not supported by the GPU build; Canu never passes it.\n"); err++; }
    }
    else if (!strcmp(a, "-m"))            G.unique = false;
    else if (!strcmp(a, "-u"))            G.unique = true;
    else if (!strcmp(a, "--hashbits"))    need(a);
    else if (!strcmp(a, "--hashdatalen")) need(a);
    else if (!strcmp(a, "--hashload"))    need(a);
    else if (!strcmp(a, "-o"))            G.outName = need(a);
    else if (!strcmp(a, "-s"))            G.statName = need(a);
    else if (!strcmp(a, "-t"))            need(a);
    else if (!strcmp(a, "--minlength"))   G.minOlapLen = (int32_t)strtol(need(a), nullptr, 10);
    else if (!strcmp(a, "--minkmers"))    G.minKmers = true;
    else if (!strcmp(a, "--maxerate"))    G.maxErate = ovlb_parse_erate(need(a));     // strtof, overlapInCore.C:380
    else if (!strcmp(a, "--alignnoise"))  G.alignNoise = ovlb_parse_erate(need(a));
    else if (!strcmp(a, "-z"))            G.noHopeless = true;
    else if (!strcmp(a, "--gpu"))         G.gpu = atoi(need(a));
    else if (!strcmp(a, "--refbatch"))    G.refBatchBases = strtoull(need(a), nullptr, 10);
    else if (!strcmp(a, "--hashblock"))   G.hashBlockBases = strtoull(need(a), nullptr, 10);
    else if (!strcmp(a, "--version"))     { printf("overlapInCore (canu_b200, B200-native ovl) for canu v2.3\n"); return 0; }
    else if (G.storePath == nullptr)      G.storePath = a;
    else { fprintf(stderr, "Unknown option '%s'\n", a); err++; }
  }
  if (G.kmerLen == 0)       { fprintf(stderr, "* No kmer length supplied; -k needed!\n"); err++; }
  if (G.outName == nullptr) { fprintf(stderr, "ERROR:  No output file name specified\n"); err++; }
  if (err || G.storePath == nullptr) { usage(argv[0]); return 1; }
  if (G.kmerLen > 31) FAIL("ERROR: k-mer length %lu is too large (max 31)", (unsigned long)G.kmerLen);

  auto t_start = std::chrono::steady_clock::now();

  SqStore store;
  std::string e;
  if (!store.open(G.storePath, e)) FAIL("sqStore()--  failed to open '%s' for read-only access: %s.", G.storePath, e.c_str());
  const uint32_t N = store.lastReadID();
  if (G.bgnHashID < 1) G.bgnHashID = 1;
  if (G.endHashID > N) G.endHashID = N;
  if (G.bgnRefID < 1)  G.bgnRefID = 1;
  if (G.endRefID > N)  G.endRefID = N;

  uint32_t maxLen = 64;
  for (uint32_t id = std::min(G.bgnHashID, G.bgnRefID); id <= std::max(G.endHashID, G.endRefID) && id <= N; id++)
    maxLen = std::max(maxLen, store.readLength(id));

  ovlb_params P;
  if (ovlb_params_init(&P, (uint32_t)G.kmerLen, G.maxErate, G.alignNoise, G.partial, G.unique, G.minOlapLen, G.noHopeless, G.minKmers, maxLen))
    FAIL("ERROR: %s", ovlb_last_error());

  if (ovlb_device_count() == 0) FAIL("ERROR: no CUDA device found; this overlapInCore has no CPU path.");
  ovlb_ctx *ctx = nullptr;
  if (ovlb_create(G.gpu, &P, &ctx)) FAIL("ERROR: %s", ovlb_last_error());

  std::vector<uint64_t> skip;
  if (G.kmerSkipFileName && load_skip_kmers(G.kmerSkipFileName, (uint32_t)G.kmerLen, skip)) return 1;

  OvFileWriter out;
  if (!out.open(G.outName, N, e)) FAIL("ERROR: %s", e.c_str());

  //  A read must be at least --minlength long to be hashed or searched, and at least K long to hold a k-mer.
  const uint32_t minLen = (uint32_t)std::max<int64_t>(G.minOlapLen, (int64_t)G.kmerLen);
  const uint64_t hashBlock = G.hashBlockBases ? G.hashBlockBases : 1500000000ull;   // index footprint ~45 B/base
  uint64_t refBatch = G.refBatchBases ? G.refBatchBases : 256000000ull;

  fprintf(stderr, "overlapInCore (B200): store '%s' has %u reads (version flags 0x%x); hash %u-%u ref %u-%u\n",
          G.storePath, N, store.version(), G.bgnHashID, G.endHashID, G.bgnRefID, G.endRefID);

  Packed HB, RB;
  std::vector<ovlb_record> recs;
  uint32_t hb = G.bgnHashID;
  while (hb <= G.endHashID && G.bgnHashID <= G.endHashID) {
    //  choose the hash block [hb, he] by bases
    uint32_t he = hb; uint64_t bases = 0;
    while (he <= G.endHashID) {
      uint32_t L = store.readLength(he);
      if (bases > 0 && bases + L > hashBlock) break;
      bases += L; he++;
    }
    he--;
    if (!pack_range(store, hb, he, G.minLibToHash, G.maxLibToHash, minLen, HB, e)) FAIL("ERROR: %s", e.c_str());
    fprintf(stderr, "Build_Hash_Index from %u to %u (%lu bases)\n", hb, he, (unsigned long)HB.bases);
    if (ovlb_load_hash_reads(ctx, &HB.view)) FAIL("ERROR: %s", ovlb_last_error());
    if (!skip.empty() && ovlb_mark_skip_kmers(ctx, skip.data(), skip.size())) FAIL("ERROR: %s", ovlb_last_error());
    if (ovlb_build_index(ctx)) FAIL("ERROR: %s", ovlb_last_error());

    //  only ref reads with ID below the last hash read can produce pairs (refID < hashID)
    const uint32_t re = std::min(G.endRefID, he > 0 ? he - 1 : 0);
    uint32_t rb = G.bgnRefID;
    while (rb <= re) {
      uint32_t r2 = rb; uint64_t rbases = 0;
      while (r2 <= re && r2 - rb < 200000) {
        uint32_t L = store.readLength(r2);
        if (rbases > 0 && rbases + L > refBatch) break;
        rbases += L; r2++;
      }
      r2--;
      if (!pack_range(store, rb, r2, G.minLibToRef, G.maxLibToRef, minLen, RB, e)) FAIL("ERROR: %s", e.c_str());
      uint64_t n = 0;
      int rc = ovlb_stage_ref_batch(ctx, &RB.view);
      if (!rc) rc = ovlb_run_staged(ctx, &n);
      if (rc == OVLB_ERR_CAPACITY && r2 > rb) {                       // seed buffers overflowed: halve the batch
        refBatch = std::max<uint64_t>(rbases / 2, 1000000);
        fprintf(stderr, "ref batch %u-%u too large for the device buffers (%s); retrying with %lu bases per batch\n",
                rb, r2, ovlb_last_error(), (unsigned long)refBatch);
        continue;
      }
      if (rc) FAIL("ERROR: %s", ovlb_last_error());
      recs.resize(n);
      if (ovlb_fetch_records(ctx, recs.data(), recs.size(), &n)) FAIL("ERROR: %s", ovlb_last_error());
      fprintf(stderr, "Processed reads %u-%u (%lu bases): %lu overlaps\n", rb, r2, (unsigned long)rbases, (unsigned long)n);
      out.submit(std::move(recs));
      recs = std::vector<ovlb_record>();
      rb = r2 + 1;
    }
    hb = he + 1;
  }

  ovlb_counters C;
  if (ovlb_get_counters(ctx, &C)) FAIL("ERROR: %s", ovlb_last_error());
  if (!out.close(e)) FAIL("ERROR: %s", e.c_str());
  ovlb_destroy(ctx);
  ovlb_params_free(&P);

  FILE *stats = stderr;
  if (G.statName) {
    stats = fopen(G.statName, "w");
    if (!stats) { fprintf(stderr, "WARNING: failed to open '%s' for writing\n", G.statName); stats = stderr; }
  }
  //  overlapInCore.C:550-558, verbatim
  fprintf(stats, " Kmer hits without olaps = %ld\n", (long)C.kmer_hits_without_olap);
  fprintf(stats, "    Kmer hits with olaps = %ld\n", (long)C.kmer_hits_with_olap);
  fprintf(stats, "  Multiple overlaps/pair = %ld\n", (long)C.multi_overlap);
  fprintf(stats, " Total overlaps produced = %ld\n", (long)C.total_overlaps);
  fprintf(stats, "      Contained overlaps = %ld\n", (long)C.contained);
  fprintf(stats, "       Dovetail overlaps = %ld\n", (long)C.dovetail);
  fprintf(stats, "Rejected by short window = %ld\n", 0L);
  fprintf(stats, " Rejected by long window = %ld\n", 0L);
  if (stats != stderr) fclose(stats);

  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  fprintf(stderr, "%lu overlaps, %lu candidate pairs, %lu DP cells in %.2f s\nBye.\n",
          (unsigned long)out.numOverlaps(), (unsigned long)C.pairs, (unsigned long)C.dp_cells, secs);
  return 0;
}
