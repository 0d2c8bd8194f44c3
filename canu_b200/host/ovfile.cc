#include "ovfile.h"

#include <algorithm>
#include <cstring>
#include <sys/stat.h>

namespace ovlhost {

static const size_t kBlockWords = 262080;          // (1 MiB / (480*4)) * 480, ovStoreFile.C:118-126

// ---------------------------------------------------------------------------------------------
//  snappy raw format
// ---------------------------------------------------------------------------------------------
size_t snappy_max_compressed(size_t n) { return 32 + n + n / 6; }

static inline uint32_t load32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

static uint8_t *emit_literal(uint8_t *op, const uint8_t *lit, size_t len) {
  if (len == 0) return op;
  size_t n = len - 1;
  if (n < 60) {
    *op++ = (uint8_t)(n << 2);
  } else {
    uint8_t *base = op++;
    int count = 0;
    while (n > 0) { *op++ = (uint8_t)(n & 0xff); n >>= 8; count++; }
    *base = (uint8_t)((59 + count) << 2);
  }
  memcpy(op, lit, len);
  return op + len;
}

static uint8_t *emit_copy_upto64(uint8_t *op, size_t offset, size_t len) {
  if (len < 12 && offset < 2048) {                 // 1-byte-offset copy: len 4..11
    *op++ = (uint8_t)(1 | ((len - 4) << 2) | ((offset >> 8) << 5));
    *op++ = (uint8_t)(offset & 0xff);
  } else {                                         // 2-byte-offset copy: len 1..64
    *op++ = (uint8_t)(2 | ((len - 1) << 2));
    *op++ = (uint8_t)(offset & 0xff);
    *op++ = (uint8_t)(offset >> 8);
  }
  return op;
}

static uint8_t *emit_copy(uint8_t *op, size_t offset, size_t len) {
  while (len >= 68) { op = emit_copy_upto64(op, offset, 64); len -= 64; }
  if (len > 64)     { op = emit_copy_upto64(op, offset, 60); len -= 60; }
  return emit_copy_upto64(op, offset, len);
}

size_t snappy_compress(const uint8_t *in, size_t n, uint8_t *out) {
  uint8_t *op = out;
  for (size_t v = n; ; ) {                         // preamble: uncompressed length as a varint
    uint8_t b = (uint8_t)(v & 0x7f); v >>= 7;
    if (v) { *op++ = b | 0x80; } else { *op++ = b; break; }
  }
  const size_t kFrag = 65536;                      // offsets stay below 64 KiB
  std::vector<uint16_t> table(1 << 14);
  for (size_t f0 = 0; f0 < n; f0 += kFrag) {
    const uint8_t *base = in + f0;
    const size_t flen = (n - f0 < kFrag) ? n - f0 : kFrag;
    std::fill(table.begin(), table.end(), 0);
    size_t ip = 0, lit = 0;
    if (flen >= 15) {
      const size_t limit = flen - 4;
      while (ip < limit) {
        uint32_t h = (load32(base + ip) * 0x1e35a7bdu) >> 18;
        size_t cand = table[h];
        table[h] = (uint16_t)ip;
        if (cand < ip && load32(base + cand) == load32(base + ip)) {
          size_t m = 4;
          while (ip + m < flen && base[cand + m] == base[ip + m]) m++;
          op = emit_literal(op, base + lit, ip - lit);
          op = emit_copy(op, ip - cand, m);
          ip += m;
          lit = ip;
        } else {
          ip++;
        }
      }
    }
    op = emit_literal(op, base + lit, flen - lit);
  }
  return (size_t)(op - out);
}

bool snappy_uncompress(const uint8_t *in, size_t n, std::vector<uint8_t> &out) {
  size_t ip = 0, ulen = 0; int shift = 0;
  while (true) {
    if (ip >= n) return false;
    uint8_t b = in[ip++];
    ulen |= (size_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  out.clear(); out.reserve(ulen);
  while (ip < n) {
    uint8_t tag = in[ip++];
    size_t len, off;
    switch (tag & 3) {
      case 0: {
        len = (tag >> 2) + 1;
        if (len > 60) { size_t nb = len - 60; if (ip + nb > n) return false; len = 0; for (size_t i = 0; i < nb; i++) len |= (size_t)in[ip + i] << (8 * i); len += 1; ip += nb; }
        if (ip + len > n) return false;
        out.insert(out.end(), in + ip, in + ip + len); ip += len;
        continue;
      }
      case 1: if (ip + 1 > n) return false; len = ((tag >> 2) & 7) + 4; off = ((size_t)(tag >> 5) << 8) | in[ip]; ip += 1; break;
      case 2: if (ip + 2 > n) return false; len = (tag >> 2) + 1; off = in[ip] | ((size_t)in[ip + 1] << 8); ip += 2; break;
      default: if (ip + 4 > n) return false; len = (tag >> 2) + 1; off = load32(in + ip); ip += 4; break;
    }
    if (off == 0 || off > out.size()) return false;
    size_t s = out.size() - off;
    for (size_t i = 0; i < len; i++) out.push_back(out[s + i]);
  }
  return out.size() == ulen;
}

// ---------------------------------------------------------------------------------------------
//  writer: a pool of threads, each turns whole batches into blocks (transpose to the 6-word record, snappy, one
//  fwrite under the file lock).  Blocks are independent and record order is unspecified (the reference's own order
//  depends on its thread interleaving), so nothing has to be re-ordered; the per-read counts are atomic adds.
// ---------------------------------------------------------------------------------------------
OvFileWriter::~OvFileWriter() { std::string e; if (file_ || !th_.empty()) close(e); }

bool OvFileWriter::open(const std::string &name, uint32_t last_read_id, std::string &err, unsigned n_threads) {
  name_ = name;
  //  findBaseFileName: strip everything from the first '.' after the last '/', unless `name` is a directory
  std::string prefix = name;
  struct stat st;
  if (!(stat(name.c_str(), &st) == 0 && S_ISDIR(st.st_mode))) {
    size_t slash = prefix.rfind('/');
    size_t dot = prefix.find('.', slash == std::string::npos ? 0 : slash);
    if (dot != std::string::npos) prefix.resize(dot);
  }
  oc_name_ = prefix + ".oc";
  file_ = fopen(name.c_str(), "wb");
  if (!file_) { err = "cannot open '" + name + "' for writing"; return false; }
  setvbuf(file_, nullptr, _IOFBF, 1 << 22);
  opr_.assign((size_t)last_read_id + 1, 0);
  if (n_threads == 0) n_threads = 1;
  for (unsigned t = 0; t < n_threads; t++) th_.emplace_back(&OvFileWriter::run, this);
  return true;
}

void OvFileWriter::submit(std::vector<ovlb_record> &&batch) {
  if (batch.empty()) return;
  auto *v = new std::vector<ovlb_record>(std::move(batch));
  submit(v->data(), v->size(), [v] { delete v; });
}

void OvFileWriter::submit(const ovlb_record *recs, size_t n, std::function<void()> done) {
  if (n == 0) { if (done) done(); return; }
  { std::lock_guard<std::mutex> lk(mu_); q_.push_back(Batch{recs, n, std::move(done)}); }
  cv_.notify_one();
}

void OvFileWriter::run() {
  std::vector<uint32_t> block; block.reserve(kBlockWords);
  std::vector<uint8_t> comp;
  while (true) {
    Batch b;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return done_ || !q_.empty(); });
      if (q_.empty()) break;
      b = std::move(q_.front());
      q_.pop_front();
    }
    const size_t per_block = kBlockWords / 6;
    for (size_t i0 = 0; i0 < b.n; i0 += per_block) {
      const size_t m = std::min(per_block, b.n - i0);
      block.clear();
      for (size_t i = i0; i < i0 + m; i++) {
        const ovlb_record &r = b.recs[i];
        block.push_back(r.a_iid);
        block.push_back(r.b_iid);
        block.push_back((uint32_t)(r.dat0 >> 32)); block.push_back((uint32_t)r.dat0);
        block.push_back((uint32_t)(r.dat1 >> 32)); block.push_back((uint32_t)r.dat1);
        if (r.a_iid < opr_.size()) __atomic_fetch_add(&opr_[r.a_iid], 1u, __ATOMIC_RELAXED);
        if (r.b_iid < opr_.size()) __atomic_fetch_add(&opr_[r.b_iid], 1u, __ATOMIC_RELAXED);
      }
      if (i0 + m == b.n && b.done) { b.done(); b.done = nullptr; }       // records consumed: the caller's buffer is free again
      const size_t nbytes = block.size() * 4;
      comp.resize(snappy_max_compressed(nbytes));
      const uint64_t cl = snappy_compress((const uint8_t *)block.data(), nbytes, comp.data());
      std::lock_guard<std::mutex> lk(file_mu_);
      if (!failed_ && (fwrite(&cl, 8, 1, file_) != 1 || fwrite(comp.data(), 1, cl, file_) != cl)) { failed_ = true; err_ = "write to '" + name_ + "' failed"; }
      n_olaps_ += m;
    }
  }
}

bool OvFileWriter::close(std::string &err) {
  if (!th_.empty()) {
    { std::lock_guard<std::mutex> lk(mu_); done_ = true; }
    cv_.notify_all();
    for (auto &t : th_) t.join();
    th_.clear();
  }
  bool ok = !failed_;
  if (file_) { if (fclose(file_) != 0) ok = false; file_ = nullptr; }
  if (!ok) { err = err_.empty() ? "closing '" + name_ + "' failed" : err_; return false; }
  FILE *oc = fopen(oc_name_.c_str(), "wb");
  if (!oc) { err = "cannot open '" + oc_name_ + "' for writing"; return false; }
  uint32_t opr_max = (uint32_t)opr_.size();
  bool w = fwrite(&n_olaps_, 8, 1, oc) == 1 && fwrite(&opr_max, 4, 1, oc) == 1 &&
           (opr_max == 0 || fwrite(opr_.data(), 4, opr_max, oc) == opr_max);
  if (fclose(oc) != 0 || !w) { err = "write to '" + oc_name_ + "' failed"; return false; }
  return true;
}

}  // namespace ovlhost
