//  ovfile.h -- writer of the overlapper output files: <name>.ovb (snappy-framed overlap records)
//  and <prefix>.oc (overlaps per read).  Asynchronous: records are handed over in batches and a
//  writer thread compresses and writes them while the GPU works on the next batch.
//
//  Format followed (reference paths under /root/reference/src/stores):
//    ovStoreFile.C:98-283     ovFileFullWrite: record = 6 x uint32 (a_iid, b_iid, hi32(dat0), lo32(dat0),
//                             hi32(dat1), lo32(dat1)); blocks of at most 262,080 words, each written as
//                             [uint64 compressed length][snappy raw block]
//    ovStoreFile.H:50-115     ovFileOCW: uint64 nOlaps, uint32 oprMax (= lastReadID+1), uint32 opr[oprMax],
//                             opr[a]++ and opr[b]++ per record; written even when there are no overlaps
//    utility/src/files/accessing-v1.C:71-88   prefix = name up to the first '.' after the last '/'
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ovlb200.h"

namespace ovlhost {

//  Snappy raw-format compressor (format_description.txt of the snappy project; the reference vendors
//  snappy 1.x under stores/libsnappy).  Any valid stream is acceptable to the reader; bytes need not
//  match the reference's compressor.  Returns the compressed size.
size_t snappy_max_compressed(size_t n);
size_t snappy_compress(const uint8_t *in, size_t n, uint8_t *out);
bool   snappy_uncompress(const uint8_t *in, size_t n, std::vector<uint8_t> &out);   // for self-checks / tests

class OvFileWriter {
 public:
  OvFileWriter() {}
  ~OvFileWriter();
  bool open(const std::string &name, uint32_t last_read_id, std::string &err, unsigned n_threads = 1);
  void submit(std::vector<ovlb_record> &&batch);       // takes ownership; returns immediately
  //  the same for a caller-owned (e.g. page-locked, recycled) buffer: `done` is called from a writer thread as soon as
  //  the records have been consumed
  void submit(const ovlb_record *recs, size_t n, std::function<void()> done);
  bool close(std::string &err);                         // drains, writes the .oc file
  uint64_t numOverlaps() const { return n_olaps_; }

 private:
  void run();
  struct Batch { const ovlb_record *recs = nullptr; size_t n = 0; std::function<void()> done; };

  std::string name_, oc_name_;
  FILE *file_ = nullptr;
  std::vector<std::thread> th_;
  std::mutex mu_, file_mu_;
  std::condition_variable cv_;
  std::deque<Batch> q_;
  bool done_ = false, failed_ = false;
  std::string err_;
  std::vector<uint32_t> opr_;
  uint64_t n_olaps_ = 0;
};

}  // namespace ovlhost
