//  main.cc -- `overlapInCore`, B200 edition: a drop-in for Canu's ovl overlapper executable.
//
//  Same command line, same sqStore input, same .ovb / .oc / .stats outputs as the reference
//  (overlapInCore.C:284-565), so the `overlap.sh` script Canu generates
//  (src/pipelines/canu/OverlapInCore.pm:175-194) can call it unchanged.  The compute path is the
//  CUDA library behind include/ovlb200.h; there is no CPU fallback.
//
//  Flags that only size the reference's CPU data structures are accepted and ignored because the
//  output does not depend on them: --hashbits, --hashload, --hashdatalen, -t.
//  Extra flags of this build: --gpu N (device, default 0), --gpus a,b,..|all, --streams N (contexts per device,
//  default 1; measured: two contexts per B200 do NOT pay -- their persistent kernels serialise and the tiles get smaller), --refbatch BASES, --hashblock BASES.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <memory>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/ovlb200.h"
#include "ovfile.h"
#include "pack.h"
#include "prefetch.h"
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
  std::vector<int> gpus;           // --gpus a,b,..  (or "all")
  uint64_t refBatchBases = 0, hashBlockBases = 0;
  int      streams = 1;            // --streams: worker contexts per device
  std::vector<int> visibleRemap;   // physical device of each visible index when we set CUDA_VISIBLE_DEVICES ourselves
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
  fprintf(stderr, "--gpus a,b,..|all  spread the job's hash-block x ref-block tiles over several CUDA devices\n");
  fprintf(stderr, "--streams n        worker contexts per device (default 1)\n");
  fprintf(stderr, "--refbatch n       bases of reference reads per device batch\n");
  fprintf(stderr, "--hashblock n      bases of hash reads per device-resident index\n\n");
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Phase { double create = 0, pack_hash = 0, index = 0, pack_ref = 0, stage = 0, run = 0, fetch = 0, submit = 0, destroy = 0, total = 0; };

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

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

//  Page-locked record buffers cycling between the workers (device -> host copies land here at full speed) and the
//  writer threads (which hand a buffer back as soon as they have consumed its records).
class RecPool {
 public:
  struct Buf { std::vector<ovlb_record> v; const void *reg = nullptr; size_t reg_bytes = 0; };
  explicit RecPool(size_t n) { for (size_t i = 0; i < n; i++) free_.push_back(new Buf()); total_ = n; }
  ~RecPool() { drain(); for (Buf *b : free_) { if (b->reg) ovlb_host_unregister(b->reg); delete b; } }
  Buf *acquire(uint64_t n_records) {
    Buf *b;
    { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [this] { return !free_.empty(); }); b = free_.back(); free_.pop_back(); }
    if (b->v.size() < n_records + 1) {
      if (b->reg) { ovlb_host_unregister(b->reg); b->reg = nullptr; }
      b->v.resize(std::max<size_t>((size_t)(n_records * 3 / 2) + 1024, 1u << 20));
      if (ovlb_host_register(b->v.data(), b->v.size() * sizeof(ovlb_record)) == 0) { b->reg = b->v.data(); b->reg_bytes = b->v.size() * sizeof(ovlb_record); }
    }
    return b;
  }
  void release(Buf *b) { { std::lock_guard<std::mutex> lk(mu_); free_.push_back(b); } cv_.notify_one(); }
  void drain() { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [this] { return free_.size() == total_; }); }
 private:
  std::mutex mu_; std::condition_variable cv_; std::vector<Buf *> free_; size_t total_ = 0;
};

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
    else if (!strcmp(a, "--readsperbatch"))  need(a);                    // reference work-distribution knobs (overlapInCore.C:352-356): no effect on output
    else if (!strcmp(a, "--readsperthread")) need(a);
    else if (!strcmp(a, "-o"))            G.outName = need(a);
    else if (!strcmp(a, "-s"))            G.statName = need(a);
    else if (!strcmp(a, "-t"))            need(a);
    else if (!strcmp(a, "--minlength"))   G.minOlapLen = (int32_t)strtol(need(a), nullptr, 10);
    else if (!strcmp(a, "--minkmers"))    G.minKmers = true;
    else if (!strcmp(a, "--maxerate"))    G.maxErate = ovlb_parse_erate(need(a));     // strtof, overlapInCore.C:380
    else if (!strcmp(a, "--alignnoise"))  G.alignNoise = ovlb_parse_erate(need(a));
    else if (!strcmp(a, "-z"))            G.noHopeless = true;
    else if (!strcmp(a, "--gpu"))         G.gpu = atoi(need(a));
    else if (!strcmp(a, "--gpus")) {
      const char *v = need(a);
      G.gpus.clear();
      if (!strcmp(v, "all")) G.gpus.push_back(-1);
      else for (const char *q = v; *q; ) { G.gpus.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
    }
    else if (!strcmp(a, "--streams"))     G.streams = std::max(1, std::min(4, atoi(need(a))));
    else if (!strcmp(a, "--refbatch"))    G.refBatchBases = strtoull(need(a), nullptr, 10);
    else if (!strcmp(a, "--hashblock"))   G.hashBlockBases = strtoull(need(a), nullptr, 10);
    else if (!strcmp(a, "--version"))     { printf("overlapInCore (canu_b200, B200-native ovl) for canu v2.3\n"); return 0; }
    else if (a[0] == '-' && a[1] != 0)    { fprintf(stderr, "Unknown option '%s'\n", a); err++; }
    else if (G.storePath == nullptr)      G.storePath = a;
    else { fprintf(stderr, "Unknown option '%s'\n", a); err++; }
  }
  if (G.kmerLen == 0)       { fprintf(stderr, "* No kmer length supplied; -k needed!\n"); err++; }
  if (G.outName == nullptr) { fprintf(stderr, "ERROR:  No output file name specified\n"); err++; }
  if (err || G.storePath == nullptr) { usage(argv[0]); return 1; }
  if (G.kmerLen < 2 || G.kmerLen > 30) FAIL("ERROR: k-mer length %lu is out of range (2..30)", (unsigned long)G.kmerLen);

  auto t_start = std::chrono::steady_clock::now();

  const double t_begin = now_s();
  //  CUDA initialisation costs seconds on a multi-GPU node (the driver brings up every visible device), so (a) when the
  //  devices are named explicitly only those are made visible, and (b) it runs on its own thread while the store is
  //  opened and the host tables are built.
  if (!G.gpus.empty() ? G.gpus[0] >= 0 : true) {
    if (getenv("CUDA_VISIBLE_DEVICES") == nullptr) {
      std::vector<int> want = G.gpus.empty() ? std::vector<int>{G.gpu} : G.gpus;
      std::vector<int> uniq(want); std::sort(uniq.begin(), uniq.end()); uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
      bool sane = true; for (int d : uniq) if (d < 0 || d > 1023) sane = false;
      if (sane) {
        std::string vis;
        for (size_t i = 0; i < uniq.size(); i++) { if (i) vis += ","; vis += std::to_string(uniq[i]); }
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
        for (int &d : want) d = (int)(std::lower_bound(uniq.begin(), uniq.end(), d) - uniq.begin());
        G.gpus = want; G.gpu = want[0];
        G.visibleRemap = uniq;
      }
    }
  }
  int ndev_async = 0;
  std::thread cuda_init([&ndev_async] { ndev_async = ovlb_device_count(); });
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } cuda_init_joiner{cuda_init};
  SqStore store;
  std::string e;
  if (!store.open(G.storePath, e)) FAIL("sqStore()--  failed to open '%s' for read-only access: %s.", G.storePath, e.c_str());
  const uint32_t N = store.lastReadID();
  if (G.bgnHashID < 1) G.bgnHashID = 1;
  if (G.endHashID > N) G.endHashID = N;
  if (G.bgnRefID < 1)  G.bgnRefID = 1;
  if (G.endRefID > N)  G.endRefID = N;

  const double t_store = now_s();
  //  A read must be at least --minlength long to be hashed or searched, and at least K long to hold a k-mer.
  const uint32_t minLen = (uint32_t)std::max<int64_t>(G.minOlapLen, (int64_t)G.kmerLen);
  uint32_t maxLen = 64;
  std::vector<uint32_t> readLen(N + 2, 0);
  uint64_t refBasesTotal = 0;
  for (uint32_t id = 1; id <= N; id++) {
    readLen[id] = store.readLength(id);
    if (id >= std::min(G.bgnHashID, G.bgnRefID) && id <= std::max(G.endHashID, G.endRefID)) maxLen = std::max(maxLen, readLen[id]);
    if (id >= G.bgnRefID && id <= G.endRefID && readLen[id] >= minLen) refBasesTotal += readLen[id];
  }

  ovlb_params P;
  if (ovlb_params_init(&P, (uint32_t)G.kmerLen, G.maxErate, G.alignNoise, G.partial, G.unique, G.minOlapLen, G.noHopeless, G.minKmers, maxLen))
    FAIL("ERROR: %s", ovlb_last_error());

  const double t_params = now_s();
  std::vector<uint64_t> skip;
  if (G.kmerSkipFileName && load_skip_kmers(G.kmerSkipFileName, (uint32_t)G.kmerLen, skip)) return 1;

  cuda_init.join();
  const int ndev = ndev_async;
  if (ndev == 0) FAIL("ERROR: no CUDA device found; this overlapInCore has no CPU path.");
  if (!G.visibleRemap.empty() && (size_t)ndev < G.visibleRemap.size()) FAIL("ERROR: --gpu/--gpus names a CUDA device that does not exist (%d of the %zu named are visible)", ndev, G.visibleRemap.size());
  if (G.gpus.empty()) G.gpus.push_back(G.gpu);
  if (G.gpus.size() == 1 && G.gpus[0] < 0) { G.gpus.clear(); for (int d = 0; d < ndev; d++) G.gpus.push_back(d); }
  for (int d : G.gpus) if (d < 0 || d >= ndev) FAIL("ERROR: --gpu/--gpus names device %d but only %d CUDA device(s) exist", d, ndev);
  const double t_cuda = now_s();
  //  --streams contexts per device, each with its own stream, buffers and share of the device's memory: while one
  //  packs reads, copies records or sits in the tail of its extension kernel, the other's kernels fill the SMs.
  //  (A device listed twice in --gpus already asks for that explicitly and is left alone.)
  std::vector<int> sharers;                                             // contexts on the device of worker wi
  {
    std::vector<int> uniq(G.gpus); std::sort(uniq.begin(), uniq.end());
    const bool listed_twice = std::adjacent_find(uniq.begin(), uniq.end()) != uniq.end();
    const int per = listed_twice ? 1 : G.streams;
    std::vector<int> all;
    for (int d : G.gpus) for (int k = 0; k < per; k++) all.push_back(d);
    G.gpus = all;
    for (int d : G.gpus) sharers.push_back((int)std::count(G.gpus.begin(), G.gpus.end(), d));
  }
  //  free memory of a device, asked once (by the first worker that gets there, on its own thread: creating the
  //  device's primary context is the slow part and must not be serialised over the devices)
  std::vector<uint64_t> devFree(ndev, 0);
  std::vector<std::mutex> devMu(ndev);
  auto phys = [&G](int d) { return (size_t)d < G.visibleRemap.size() ? G.visibleRemap[d] : d; };
  const uint32_t W = (uint32_t)G.gpus.size();

  OvFileWriter out;
  //  compression threads of the writer: HiFi-like jobs produce ~0.6 GB/s of records per GPU (26 MB per 45 ms tile)
  const unsigned writerThreads = (unsigned)std::max<size_t>(2, std::min<size_t>(16, 2 * G.gpus.size()));
  if (!out.open(G.outName, N, e, writerThreads)) FAIL("ERROR: %s", e.c_str());

  //  Re-block the job inside the process (the output does not depend on blocking): hash blocks
  //  sized for HBM, ref batches sized for the device seed buffers -- and small enough that every GPU gets several.
  //  Hash block from the memory a context gets: ovlb_build_index needs, per hash base, 1 B of dp4 reads (both
  //  orientations) + ~40 B of tuple scratch (16 B partition records, 12 - 24 B of bucket tuples, occurrence) and, per
  //  DISTINCT k-mer, ~100 B (slot, scratch slot + first position, 32 B of table buckets) -- and in a large job most
  //  k-mers of a block are distinct (a block covers the genome less than once), so the model is 150 B per base.
  //  Everything else is taken off the budget first: the seed-run buffers (1/12 of it, ovl_index.cu), the extension
  //  scratch (from the longest read and the error rate, at most 1/4, ovl_extend.cu), two ref slots and the records.
  //  Bigger blocks mean fewer tiles: every ref window is looked up once per hash block, and where most lookups miss
  //  (a human-size job) a tile costs the same whatever the block holds.
  uint64_t minBudget = ~0ull;
  for (uint32_t wi = 0; wi < (uint32_t)G.gpus.size(); wi++) {
    const int d = G.gpus[wi];
    uint64_t tot = 0;                                                   // total, not free: asking for free memory would create the
    if (ovlb_device_total_memory(d, &tot)) FAIL("ERROR: %s", ovlb_last_error());   // device's context here, serially over the GPUs
    minBudget = std::min<uint64_t>(minBudget, (uint64_t)((double)tot * 0.95 * 0.8 / sharers[wi]));
  }
  const uint64_t refBatchDefault = 256000000ull;
  const uint64_t hashBlock = G.hashBlockBases ? G.hashBlockBases
                                              : ovlb_hash_block_bases(minBudget, maxLen, G.maxErate, G.refBatchBases ? G.refBatchBases : refBatchDefault);
  uint64_t refBatch = G.refBatchBases ? G.refBatchBases : refBatchDefault;
  //  Several workers: every one should get a few tiles of each hash block so that longest-first assignment can balance
  //  them, but not many small ones -- an extension launch cannot end before its slowest pair does (0.2 - 1 s on noisy
  //  reads, tools/scale_run.py), so a tile should hold tens of pairs per resident warp.
  if (W > 1 && !G.refBatchBases) refBatch = std::max<uint64_t>(std::min<uint64_t>(refBatch, refBasesTotal / (2ull * W) + 1), 16000000ull);

  std::vector<ovlb_tile> tiles;
  {
    uint64_t nt = 0;
    //  blocks are cut on read lengths with --minlength 0 so that short reads still belong to some block
    if (G.bgnHashID <= G.endHashID && G.bgnRefID <= G.endRefID) {
      if (ovlb_plan_tiles(readLen.data(), N, 0, hashBlock, refBatch, G.bgnHashID, G.endHashID, G.bgnRefID, G.endRefID, 0, nullptr, 0, &nt))
        FAIL("ERROR: %s", ovlb_last_error());
      tiles.resize(nt);
      if (nt && ovlb_plan_tiles(readLen.data(), N, 0, hashBlock, refBatch, G.bgnHashID, G.endHashID, G.bgnRefID, G.endRefID, 0, tiles.data(), nt, &nt))
        FAIL("ERROR: %s", ovlb_last_error());
    }
  }
  //  Fewer hash blocks than twice the workers (small and medium jobs; always the case when the job fits one block): every
  //  worker indexes every hash block and takes ONE cost-balanced part of its ref range -- an extension
  //  launch cannot end before its slowest pair, so few large launches beat the many small tiles LPT would need, and the
  //  triangular refID < hashID rule makes equal-base tiles unequal work (ovlb_plan_balanced).  A part larger than the
  //  device's ref batch is cut further, all pieces staying with the same worker.
  std::vector<int32_t> fixedOwner;
  if (W > 1 && !tiles.empty()) {
    std::vector<std::pair<uint32_t, uint32_t>> blocks;
    for (const ovlb_tile &T : tiles) if (blocks.empty() || blocks.back().first != T.hash_bgn) blocks.push_back({T.hash_bgn, T.hash_end});
    //  noisy reads: the index build is < 1 % of a tile, so balanced parts win until there are many blocks per GPU (C3 on two
    //  GPUs with 9 whole blocks dealt out: 197.7 s against 176.9 s of extension); HiFi-like reads: the build is a quarter
    const size_t wholeBlocksFrom = (G.maxErate >= 0.03 ? 16 : 2) * (size_t)W;
    if (blocks.size() < wholeBlocksFrom) {
      const double lookupWeight = G.maxErate >= 0.03 ? 0.01 : 0.6;
      const uint64_t maxPiece = G.refBatchBases ? G.refBatchBases : 256000000ull;
      std::vector<ovlb_tile> bal;
      for (auto &hb : blocks) {
        std::vector<ovlb_tile> parts(W);
        uint64_t np2 = 0;
        if (ovlb_plan_balanced(readLen.data(), N, 0, hb.first, hb.second, G.bgnRefID, G.endRefID, W, lookupWeight, parts.data(), W, &np2)) FAIL("ERROR: %s", ovlb_last_error());
        for (uint64_t pi = 0; pi < np2; pi++) {
          const ovlb_tile &Pt = parts[pi];
          const uint32_t pieces = (uint32_t)std::max<uint64_t>(1, (Pt.ref_bases + maxPiece - 1) / maxPiece);
          const uint64_t per = Pt.ref_bases / pieces + 1;
          uint32_t b = Pt.ref_bgn;
          while (b <= Pt.ref_end) {
            uint64_t acc = 0; uint32_t e2 = b;
            while (e2 < Pt.ref_end && acc + readLen[e2] < per) { acc += readLen[e2]; e2++; }
            ovlb_tile Q = Pt; Q.ref_bgn = b; Q.ref_end = e2; Q.ref_bases = 0;
            for (uint32_t id = b; id <= e2; id++) Q.ref_bases += readLen[id];
            Q.cost = Pt.cost * (double)Q.ref_bases / (double)std::max<uint64_t>(Pt.ref_bases, 1);
            bal.push_back(Q); fixedOwner.push_back((int32_t)pi);
            b = e2 + 1;
          }
        }
      }
      tiles.swap(bal);
    }
  }
  //  The device keys a seed run by (ref read:18 bits | hash read:24 bits): cut ref ranges of more than 200 000 reads up
  //  front (short-read stores), and refuse hash blocks of more than 2^24 - 1 reads with a usable message.
  {
    std::vector<ovlb_tile> cut;
    std::vector<int32_t> cutOwner;
    for (size_t ti = 0; ti < tiles.size(); ti++) {
      const ovlb_tile &T = tiles[ti];
      const int32_t fo = fixedOwner.empty() ? -1 : fixedOwner[ti];
      if (T.hash_end - T.hash_bgn + 1 >= (1u << 24))
        FAIL("ERROR: hash block %u-%u holds more than 16777215 reads; use a smaller --hashblock", T.hash_bgn, T.hash_end);
      const uint32_t maxRef = 200000;
      if (T.ref_end - T.ref_bgn + 1 <= maxRef) { cut.push_back(T); cutOwner.push_back(fo); continue; }
      const uint32_t pieces = (T.ref_end - T.ref_bgn + maxRef) / maxRef;
      for (uint32_t b = T.ref_bgn; b <= T.ref_end; ) {
        const uint32_t e2 = (uint32_t)std::min<uint64_t>(T.ref_end, (uint64_t)b + maxRef - 1);
        ovlb_tile P2 = T; P2.ref_bgn = b; P2.ref_end = e2;
        P2.ref_bases = 0; for (uint32_t id = b; id <= e2; id++) P2.ref_bases += readLen[id];
        P2.cost = T.cost / pieces;
        cut.push_back(P2); cutOwner.push_back(fo);
        if (e2 == T.ref_end) break;
        b = e2 + 1;
      }
    }
    tiles.swap(cut);
    if (!fixedOwner.empty()) fixedOwner.swap(cutOwner);
  }
  //  owner of every tile: whole hash blocks per GPU when there are plenty of them (no index is built twice),
  //  else tile by tile (the hash block is then indexed on every GPU that got one of its tiles)
  std::vector<uint32_t> owner(tiles.size(), 0);
  if (!fixedOwner.empty()) {
    for (size_t i = 0; i < tiles.size(); i++) owner[i] = (uint32_t)fixedOwner[i];
  } else if (W > 1 && !tiles.empty()) {
    std::vector<ovlb_tile> blocks;                                     // one pseudo-tile per hash block, cost summed
    std::vector<size_t> blockOf(tiles.size());
    for (size_t i = 0; i < tiles.size(); i++) {
      if (blocks.empty() || blocks.back().hash_bgn != tiles[i].hash_bgn) { blocks.push_back(tiles[i]); blocks.back().cost = 0; }
      blocks.back().cost += tiles[i].cost;
      blockOf[i] = blocks.size() - 1;
    }
    if (blocks.size() >= 2 * (size_t)W) {
      std::vector<uint32_t> bo(blocks.size());
      if (ovlb_assign_tiles(blocks.data(), blocks.size(), W, bo.data())) FAIL("ERROR: %s", ovlb_last_error());
      for (size_t i = 0; i < tiles.size(); i++) owner[i] = bo[blockOf[i]];
    } else if (ovlb_assign_tiles(tiles.data(), tiles.size(), W, owner.data())) FAIL("ERROR: %s", ovlb_last_error());
  }

  fprintf(stderr, "overlapInCore (B200): store '%s' has %u reads (version flags 0x%x); hash %u-%u ref %u-%u; %zu tile(s) on %u context(s) of %zu GPU(s)\n",
          G.storePath, N, store.version(), G.bgnHashID, G.endHashID, G.bgnRefID, G.endRefID, tiles.size(), W, std::set<int>(G.gpus.begin(), G.gpus.end()).size());

  //  One worker thread per GPU; tiles share nothing, records go to the one writer thread, counters are summed.
  std::vector<ovlb_counters> counters(W);
  std::vector<Phase> phase(W);
  const double t_setup_done = now_s();
  std::vector<std::string> werr(W);
  std::mutex log_mu;
  auto worker = [&](uint32_t wi) {
    memset(&counters[wi], 0, sizeof(ovlb_counters));
    std::string err;
    SqStore st;                                                         // the reader is stateful: one per thread
    if (!st.open(G.storePath, err)) { werr[wi] = err; return; }
    ovlb_ctx *ctx = nullptr;
    Phase &ph = phase[wi];
    const double t_worker = now_s();
    //  what this worker will read, in order: the hash block whenever it changes, then the tile's ref range
    std::vector<PackItem> plan;
    {
      uint32_t hb = 0, he = 0;
      for (size_t ti = 0; ti < tiles.size(); ti++) {
        if (owner[ti] != wi) continue;
        const ovlb_tile &T = tiles[ti];
        if (T.hash_bgn != hb || T.hash_end != he) { plan.push_back({true, T.hash_bgn, T.hash_end}); hb = T.hash_bgn; he = T.hash_end; }
        const uint32_t re = std::min(T.ref_end, T.hash_end > 0 ? T.hash_end - 1 : 0);
        if (T.ref_bgn <= re) plan.push_back({false, T.ref_bgn, re});
      }
    }
    //  HiFi-like jobs run through a ref batch in ~12 ms: three packers per GPU, one more batch in flight
    const unsigned packers = G.maxErate < 0.03 ? 3 : 1;
    Prefetcher pf(G.storePath, plan, G.minLibToHash, G.maxLibToHash, G.minLibToRef, G.maxLibToRef, minLen, packers > 1 ? 4 : 3, packers);
    double t0 = now_s();
    ovlb_params Pw = P;
    {
      const int d = G.gpus[wi];
      std::lock_guard<std::mutex> lk(devMu[d]);
      if (devFree[d] == 0) { uint64_t tot = 0; if (ovlb_device_memory(d, &devFree[d], &tot)) { werr[wi] = ovlb_last_error(); return; } }
      Pw.device_mem_budget = (uint64_t)((double)devFree[d] * 0.8 / sharers[wi]);
    }
    if (ovlb_create(G.gpus[wi], &Pw, &ctx)) { werr[wi] = ovlb_last_error(); return; }
    ph.create += now_s() - t0;
    std::unique_ptr<Packed> HBp;
    uint32_t curHb = 0, curHe = 0;
    //  Ref batches of the tiles of the current hash block, pipelined: while batch i runs, batch i+1 is
    //  already packed (prefetch thread), page-locked and being uploaded + encoded on the copy stream into the context's
    //  second ref slot; its records come back into a page-locked buffer that a writer thread hands back when it has
    //  consumed them.  A batch that overflows the device's seed buffers is cut in two and re-queued.
    struct Work { uint32_t rb, re; bool planned; };
    struct Staged { std::unique_ptr<Packed> own; uint32_t rb = 0, re = 0; };
    std::deque<Work> q;
    auto pin = [&](Packed &P) {                                           // page-lock the batch (once per buffer: recycled buffers keep their storage)
      //  only the packed bases (the per-read arrays are a few hundred KB and go first, see ovl_upload_reads): every
      //  cudaHostRegister costs tens of milliseconds whatever the size (measured: 25 registrations = 0.7 s)
      const void *ptrs[5] = { P.packed.data(), nullptr, nullptr, nullptr, nullptr };
      const size_t bytes[5] = { P.packed.capacity(), 0, 0, 0, 0 };
      for (int k = 0; k < 1; k++) {
        if (P.pinned[k].first == ptrs[k] && P.pinned[k].second == bytes[k]) continue;
        if (P.pinned[k].first) ovlb_host_unregister(P.pinned[k].first);
        P.pinned[k] = {nullptr, 0};
        if (ptrs[k] && bytes[k] && ovlb_host_register(ptrs[k], bytes[k]) == 0) P.pinned[k] = {ptrs[k], bytes[k]};
      }
    };
    auto take = [&](Staged &S, std::string &err) -> bool {                // next batch of the queue, packed and pinned
      const Work w = q.front(); q.pop_front();
      double t1 = now_s();
      //  page-locking pays where uploads are a visible share of a tile (HiFi-like reads: 33 ms tiles); a noisy tile runs for
      //  seconds and a registration costs tens of milliseconds (C3 on 4 GPUs: 3.1 s of `stage` for nothing)
      static const bool pinBatches = G.maxErate < 0.03;
      if (w.planned) { S.own = pf.next(err); if (!S.own) return false; }
      else { S.own = pf.spare(); if (!pack_range(st, w.rb, w.re, G.minLibToRef, G.maxLibToRef, minLen, *S.own, err)) return false; }
      S.rb = w.rb; S.re = w.re;
      ph.pack_ref += now_s() - t1; t1 = now_s();
      if (pinBatches) pin(*S.own);
      ph.stage += now_s() - t1;
      return true;
    };
    RecPool pool(3);
    for (size_t ti = 0; ti < tiles.size() && werr[wi].empty(); ti++) {
      if (owner[ti] != wi) continue;
      const ovlb_tile &T = tiles[ti];
      if (T.hash_bgn != curHb || T.hash_end != curHe) {
        t0 = now_s();
        pf.recycle(std::move(HBp));
        HBp = pf.next(err);
        if (!HBp) { werr[wi] = err; break; }
        Packed &HB = *HBp;
        ph.pack_hash += now_s() - t0; t0 = now_s();
        { std::lock_guard<std::mutex> lk(log_mu); fprintf(stderr, "[gpu %d] Build_Hash_Index from %u to %u (%lu bases)\n", phys(G.gpus[wi]), T.hash_bgn, T.hash_end, (unsigned long)HB.bases); }
        if (ovlb_load_hash_reads(ctx, &HB.view) ||
            (!skip.empty() && ovlb_mark_skip_kmers(ctx, skip.data(), skip.size())) ||
            ovlb_build_index(ctx)) { werr[wi] = ovlb_last_error(); break; }
        ph.index += now_s() - t0;
        curHb = T.hash_bgn; curHe = T.hash_end;
      }
      //  all of this worker's tiles of this hash block, in plan order (the prefetcher packs them in that order);
      //  only ref reads with ID below the last hash read can produce pairs (refID < hashID)
      q.clear();
      size_t tj = ti;
      for (; tj < tiles.size(); tj++) {
        if (owner[tj] != wi) continue;
        if (tiles[tj].hash_bgn != curHb || tiles[tj].hash_end != curHe) break;
        const uint32_t re = std::min(tiles[tj].ref_end, tiles[tj].hash_end > 0 ? tiles[tj].hash_end - 1 : 0);
        if (tiles[tj].ref_bgn <= re) q.push_back({tiles[tj].ref_bgn, re, true});
        ti = tj;
      }
      Staged A, B;
      bool haveA = false;
      if (!q.empty()) {
        if (!take(A, err)) { werr[wi] = err; break; }
        t0 = now_s();
        if (ovlb_stage_ref_batch(ctx, &A.own->view)) { werr[wi] = ovlb_last_error(); break; }
        ph.stage += now_s() - t0;
        haveA = true;
      }
      while (haveA && werr[wi].empty()) {
        bool haveB = false;
        if (!q.empty()) {
          if (!take(B, err)) { werr[wi] = err; break; }
          t0 = now_s();
          if (ovlb_stage_next_ref_batch(ctx, &B.own->view)) { werr[wi] = ovlb_last_error(); break; }
          ph.stage += now_s() - t0;
          haveB = true;
        }
        t0 = now_s();
        uint64_t n = 0;
        int rc = ovlb_run_staged(ctx, &n);
        ph.run += now_s() - t0; t0 = now_s();
        if (rc == OVLB_ERR_CAPACITY && A.re > A.rb) {                     // seed buffers overflowed: halve the batch, re-queue
          const uint32_t mid = A.rb + (A.re - A.rb) / 2;
          q.push_front({mid + 1, A.re, false}); q.push_front({A.rb, mid, false});
        } else if (rc) { werr[wi] = ovlb_last_error(); break; }
        else {
          RecPool::Buf *rb = pool.acquire(n);
          if (ovlb_fetch_records(ctx, rb->v.data(), rb->v.size(), &n)) { werr[wi] = ovlb_last_error(); pool.release(rb); break; }
          ph.fetch += now_s() - t0;
          { std::lock_guard<std::mutex> lk(log_mu); fprintf(stderr, "[gpu %d] Processed reads %u-%u against %u-%u (%lu bases): %lu overlaps\n", phys(G.gpus[wi]), A.rb, A.re, curHb, curHe, (unsigned long)A.own->bases, (unsigned long)n); }
          t0 = now_s();
          out.submit(rb->v.data(), n, [&pool, rb] { pool.release(rb); });
          ph.submit += now_s() - t0;
        }
        if (ovlb_advance_staged(ctx)) { werr[wi] = ovlb_last_error(); break; }
        pf.recycle(std::move(A.own));
        if (haveB) { A = std::move(B); }
        else if (!q.empty()) {                                            // only halves of an overflowed batch are left
          if (!take(A, err)) { werr[wi] = err; break; }
          t0 = now_s();
          if (ovlb_stage_ref_batch(ctx, &A.own->view)) { werr[wi] = ovlb_last_error(); break; }
          ph.stage += now_s() - t0;
        } else haveA = false;
      }
    }
    pool.drain();
    if (werr[wi].empty() && ovlb_get_counters(ctx, &counters[wi])) werr[wi] = ovlb_last_error();
    //  The context is NOT destroyed: the process _exit()s right after the outputs are closed, and freeing tens of GB
    //  buffer by buffer (every cudaFree synchronises the device) cost 0.2 - 2 s per worker for nothing.
    (void)ctx;
    ph.total = now_s() - t_worker;
  };
  if (W == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t wi = 0; wi < W; wi++) th.emplace_back(worker, wi);
    for (auto &t : th) t.join();
  }
  for (uint32_t wi = 0; wi < W; wi++) if (!werr[wi].empty()) FAIL("ERROR: [gpu %d] %s", phys(G.gpus[wi]), werr[wi].c_str());

  ovlb_counters C;
  memset(&C, 0, sizeof(C));
  for (uint32_t wi = 0; wi < W; wi++) {
    const uint64_t *src = reinterpret_cast<const uint64_t *>(&counters[wi]);
    uint64_t *dst = reinterpret_cast<uint64_t *>(&C);
    for (size_t k = 0; k < sizeof(ovlb_counters) / 8; k++) dst[k] += src[k];
  }
  const double t_workers_done = now_s();
  if (!out.close(e)) FAIL("ERROR: %s", e.c_str());
  ovlb_params_free(&P);
  const double t_closed = now_s();
  fprintf(stderr, "phases (s): setup %.2f (store %.2f, lengths+tables %.2f, wait for cuda init %.2f, plan+open %.2f) | workers %.2f | writer drain %.2f\n",
          t_setup_done - t_begin, t_store - t_begin, t_params - t_store, t_cuda - t_params, t_setup_done - t_cuda,
          t_workers_done - t_setup_done, t_closed - t_workers_done);
  for (uint32_t wi = 0; wi < W; wi++)
    fprintf(stderr, "  [gpu %d] create %.2f  pack-hash %.2f  load+index %.2f  pack-ref %.2f  stage %.2f  run %.2f  fetch %.2f  submit %.2f  destroy %.2f  (worker total %.2f)\n", phys(G.gpus[wi]),
            phase[wi].create, phase[wi].pack_hash, phase[wi].index, phase[wi].pack_ref, phase[wi].stage, phase[wi].run, phase[wi].fetch,
            phase[wi].submit, phase[wi].destroy, phase[wi].total);

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
  //  every output is closed; skip the CUDA runtime's process-exit teardown (hundreds of ms per device)
  fflush(stdout); fflush(stderr);
  _exit(0);
}
