#include "sqstore.h"
#include <unistd.h>
#include <sys/mman.h>
#include <fcntl.h>

#include <cstring>
#include <sys/stat.h>

namespace ovlhost {

static bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

SqStore::~SqStore() { for (BlobMap &m : blob_maps_) if (m.p) munmap(const_cast<uint8_t *>(m.p), m.n); }

bool SqStore::readFile(const std::string &name, std::vector<uint8_t> &out, std::string &err) const {
  std::string p = path_ + "/" + name;
  FILE *f = fopen(p.c_str(), "rb");
  if (!f) { err = "cannot open '" + p + "'"; return false; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize((size_t)sz);
  if (sz > 0 && fread(out.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); err = "short read on '" + p + "'"; return false; }
  fclose(f);
  return true;
}

template <typename T>
static bool load_array(const std::vector<uint8_t> &raw, size_t n, std::vector<T> &out) {
  if (raw.size() < n * sizeof(T)) return false;
  out.resize(n);
  memcpy(out.data(), raw.data(), n * sizeof(T));
  return true;
}

bool SqStore::open(const std::string &path, std::string &err) {
  path_ = path;
  struct stat st;
  if (stat(path.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) { err = "sqStore '" + path + "' doesn't exist"; return false; }

  std::vector<uint8_t> info;
  if (!readFile("info", info, err)) return false;
  if (info.size() < 8 || memcmp(info.data(), "MAGC", 4) != 0) { err = "sqStore info file is not in the IFF (version 9) format"; return false; }
  bool have_nrea = false;
  for (size_t p = 0; p + 8 <= info.size(); ) {
    char tag[5] = {0}; memcpy(tag, &info[p], 4);
    uint32_t len; memcpy(&len, &info[p + 4], 4);
    const uint8_t *d = &info[p + 8];
    if (p + 8 + len > info.size()) { err = "truncated sqStore info file"; return false; }
    auto u32 = [&]() { uint32_t v = 0; memcpy(&v, d, len < 4 ? len : 4); return v; };
    if      (!strcmp(tag, "NLIB")) num_libs_ = u32();
    else if (!strcmp(tag, "NREA")) { num_reads_ = u32(); have_nrea = true; }
    else if (!strcmp(tag, "NBLO")) num_blobs_ = u32();
    else if (!strcmp(tag, "MRLB")) { if (u32() != 21) { err = "sqStore was built with AS_MAX_READLEN_BITS != 21"; return false; } }
    else if (!strcmp(tag, "READ")) memcpy(reads_by_version_, d, len < sizeof(reads_by_version_) ? len : sizeof(reads_by_version_));
    p += 8 + (size_t)len;
  }
  if (!have_nrea) { err = "sqStore info file has no NREA object"; return false; }

  const size_t n = (size_t)num_reads_ + 1;
  std::vector<uint8_t> raw;
  if (!readFile("reads", raw, err)) return false;
  if (!load_array(raw, 2 * n, meta_)) { err = "sqStore 'reads' file is too short"; return false; }
  struct { const char *name; std::vector<SqReadSeq> *dst; } seqs[] = {
    { "reads-rawu", &rawu_ }, { "reads-rawc", &rawc_ }, { "reads-coru", &coru_ }, { "reads-corc", &corc_ } };
  for (auto &s : seqs) {
    if (!readFile(s.name, raw, err)) return false;
    if (!load_array(raw, n, *s.dst)) { err = std::string("sqStore '") + s.name + "' file is too short"; return false; }
  }

  //  default version: the most recent of raw < raw-trimmed < corrected < corrected-trimmed that has reads,
  //  plus 'compressed' when the store is flagged for homopolymer compression (sqStoreConstructor.C:140-166)
  uint32_t mr = 0;
  if (reads_by_version_[SQ_RAW] > 0)                        mr = SQ_RAW;
  if (reads_by_version_[SQ_RAW | SQ_TRIMMED] > 0)           mr = SQ_RAW | SQ_TRIMMED;
  if (reads_by_version_[SQ_CORRECTED] > 0)                  mr = SQ_CORRECTED;
  if (reads_by_version_[SQ_CORRECTED | SQ_TRIMMED] > 0)     mr = SQ_CORRECTED | SQ_TRIMMED;
  if (mr == 0 && num_reads_ > 0) { err = "sqStore has no raw or corrected reads"; return false; }
  which_ = mr;
  if (file_exists(path_ + "/homopolymerCompression")) which_ |= SQ_COMPRESSED;

  blob_maps_.assign((size_t)num_blobs_ + 2, BlobMap());
  return true;
}

const SqReadSeq &SqStore::seq(uint32_t id) const {
  const bool cmp = which_ & SQ_COMPRESSED;
  if (which_ & SQ_RAW) return cmp ? rawc_[id] : rawu_[id];
  return cmp ? corc_[id] : coru_[id];
}

uint32_t SqStore::readLength(uint32_t id) const {
  if (id == 0 || id > num_reads_) return 0;
  const SqReadSeq &s = seq(id);
  if (which_ & SQ_TRIMMED)
    return (!s.trimmed() || !s.valid() || s.ignoreT()) ? 0 : (s.clearEnd() - s.clearBgn());
  return (!s.valid() || s.ignoreU()) ? 0 : s.length();
}

//  Locate the encoded sequence chunk (raw or corrected) of read `id` inside its BLOB.
bool SqStore::fetchChunk(uint32_t id, const uint8_t *&chunk, uint32_t &chunk_len, char &enc, std::string &err) {
  const uint64_t m1 = meta_[2 * (size_t)id + 1];
  const uint32_t segm = (uint32_t)((m1 >> 8) & 0xffff);
  const uint64_t byte = m1 >> 24;
  //  The blob files are mapped, not read: a read's chunk is handed out as a pointer into the mapping, so packing a
  //  2-bit read is one memcpy (stdio cost 4.8 us per read: a seek, two freads and an intermediate copy).
  if (segm >= blob_maps_.size()) blob_maps_.resize(segm + 1);
  if (!blob_maps_[segm].p) {
    char name[64]; snprintf(name, sizeof(name), "/blobs.%04u", segm);
    const std::string fn = path_ + name;
    const int fd = ::open(fn.c_str(), O_RDONLY);
    struct stat sb;
    if (fd < 0 || fstat(fd, &sb) != 0) { if (fd >= 0) ::close(fd); err = "cannot open '" + fn + "'"; return false; }
    void *m = sb.st_size ? mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
    ::close(fd);
    if (m == MAP_FAILED) { err = "cannot map '" + fn + "'"; return false; }
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    blob_maps_[segm].p = static_cast<const uint8_t *>(m); blob_maps_[segm].n = (size_t)sb.st_size;
  }
  const BlobMap &M = blob_maps_[segm];
  if (byte + 8 > M.n || memcmp(M.p + byte, "BLOB", 4) != 0) {
    err = "read " + std::to_string(id) + ": no BLOB at segment " + std::to_string(segm) + " byte " + std::to_string(byte);
    return false;
  }
  uint32_t blen; memcpy(&blen, M.p + byte + 4, 4);
  if (byte + 8 + (uint64_t)blen > M.n) { err = "read " + std::to_string(id) + ": truncated BLOB"; return false; }
  const uint8_t *blob = M.p + byte + 8;
  const char want = (which_ & SQ_RAW) ? 'R' : 'C';
  for (uint32_t p = 0; p + 8 <= blen; ) {
    const uint8_t *t = blob + p;
    uint32_t clen; memcpy(&clen, t + 4, 4);
    if (t[1] == 'S' && t[2] == 'Q' && t[3] == (uint8_t)want && (t[0] == '2' || t[0] == '3' || t[0] == 'U')) {
      chunk = t + 8; chunk_len = clen; enc = (char)t[0];
      if (p + 8 + (uint64_t)clen > blen) { err = "read " + std::to_string(id) + ": chunk overruns BLOB"; return false; }
      return true;
    }
    p += 8 + clen;
  }
  err = "read " + std::to_string(id) + ": no sequence chunk in BLOB";
  return false;
}

uint32_t homopolyCompress(const std::string &in, std::string &out) {
  out.clear();
  if (in.empty()) return 0;
  out.push_back(in[0]);
  for (size_t i = 1; i < in.size(); i++)
    if ((in[i] | 0x20) != (out.back() | 0x20)) out.push_back(in[i]);
  return (uint32_t)out.size();
}

bool SqStore::loadRead(uint32_t id, std::string &bases, std::string &err) {
  bases.clear();
  if (readLength(id) == 0) return true;
  const uint8_t *chunk; uint32_t clen; char enc;
  if (!fetchChunk(id, chunk, clen, enc, err)) return false;
  const uint32_t ulen = ((which_ & SQ_RAW) ? rawu_[id] : coru_[id]).length();   // stored (uncompressed, untrimmed) length
  std::string full;
  full.resize(ulen);
  static const char D[5] = { 'A', 'C', 'G', 'T', 'N' };
  if (enc == '2') {
    if ((uint64_t)clen * 4 < ulen) { err = "read " + std::to_string(id) + ": 2-bit chunk too short"; return false; }
    for (uint32_t i = 0; i < ulen; i++) full[i] = D[(chunk[i >> 2] >> (6 - 2 * (i & 3))) & 3];
  } else if (enc == '3') {
    if ((uint64_t)clen * 3 < ulen) { err = "read " + std::to_string(id) + ": 3-bit chunk too short"; return false; }
    for (uint32_t i = 0; i < ulen; i += 3) {
      uint8_t b = chunk[i / 3];
      uint8_t c1 = b / 25; b -= c1 * 25;
      uint8_t c2 = b / 5;  b -= c2 * 5;
      uint8_t c3 = b;
      if (c1 > 4 || c2 > 4 || c3 > 4) { err = "read " + std::to_string(id) + ": bad 3-bit code"; return false; }
      full[i] = D[c1];
      if (i + 1 < ulen) full[i + 1] = D[c2];
      if (i + 2 < ulen) full[i + 2] = D[c3];
    }
  } else {
    if (clen < ulen) { err = "read " + std::to_string(id) + ": 8-bit chunk too short"; return false; }
    memcpy(&full[0], chunk, ulen);
  }
  if (which_ & SQ_COMPRESSED) { std::string c; homopolyCompress(full, c); full.swap(c); }
  if (which_ & SQ_TRIMMED) {
    const SqReadSeq &s = seq(id);
    if (s.clearEnd() > full.size() || s.clearBgn() > s.clearEnd()) { err = "read " + std::to_string(id) + ": clear range outside the read"; return false; }
    bases = full.substr(s.clearBgn(), s.clearEnd() - s.clearBgn());
  } else {
    bases.swap(full);
  }
  for (char &ch : bases) if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
  return true;
}

int SqStore::appendPacked2bit(uint32_t id, std::vector<uint8_t> &packed, std::string &err) {
  if (which_ & SQ_COMPRESSED) return 0;
  const uint32_t len = readLength(id);
  if (len == 0) return 1;
  uint32_t bgn = 0;
  if (which_ & SQ_TRIMMED) bgn = seq(id).clearBgn();
  if (bgn & 3u) return 0;
  const uint8_t *chunk; uint32_t clen; char enc;
  if (!fetchChunk(id, chunk, clen, enc, err)) return -1;
  if (enc != '2') return 0;
  const uint32_t b0 = bgn >> 2, nb = (len + 3) >> 2;
  if ((uint64_t)b0 + nb > clen) { err = "read " + std::to_string(id) + ": 2-bit chunk too short"; return -1; }
  packed.insert(packed.end(), chunk + b0, chunk + b0 + nb);
  return 1;
}

int SqStore::appendRaw2bit(uint32_t id, std::vector<uint8_t> &packed, uint32_t &src_len, uint32_t &clear_bgn, std::string &err) {
  if (readLength(id) == 0) return 0;
  const uint8_t *chunk; uint32_t clen; char enc;
  if (!fetchChunk(id, chunk, clen, enc, err)) return -1;
  if (enc != '2') return 0;
  const uint32_t ulen = ((which_ & SQ_RAW) ? rawu_[id] : coru_[id]).length();
  if ((uint64_t)clen * 4 < ulen) { err = "read " + std::to_string(id) + ": 2-bit chunk too short"; return -1; }
  src_len = ulen;
  clear_bgn = (which_ & SQ_TRIMMED) ? seq(id).clearBgn() : 0;
  packed.insert(packed.end(), chunk, chunk + ((ulen + 3) >> 2));
  return 1;
}

}  // namespace ovlhost
