//  prefetch.h -- packs the read ranges a worker is going to need, in plan order, on threads of their own and a few
//  items ahead of the consumer (used by the overlapInCore executable; `ovltool prefetch-check` is its self-test).
#pragma once

#include <algorithm>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pack.h"
#include "sqstore.h"

namespace ovlhost {

//  Packs the read ranges a worker is going to need, in order, on a thread of its own and a few items ahead: reading
//  and packing a batch from the sqStore costs about as much host time as the GPU needs for it on HiFi-like reads, and
//  the first batches are packed while the CUDA context is still being created.
struct PackItem { bool is_hash; uint32_t bgn, end; };
class Prefetcher {
 public:
  //  `n_threads` packers work on consecutive items of the plan at once (each with its own store handle); items are
  //  handed out in plan order, at most `depth` of them packed or being packed ahead of the consumer.  One packer keeps
  //  up with a noisy job; a HiFi-like job consumes a 256 Mbase batch every ~12 ms per GPU and one thread copies the stored
  //  blobs at ~2.5 GB/s (C5 fraction: 0.37 s of a 0.8 s run spent waiting for it).
  Prefetcher(const char *store_path, std::vector<PackItem> plan, uint32_t minLibH, uint32_t maxLibH, uint32_t minLibR, uint32_t maxLibR,
             uint32_t min_len, size_t depth, unsigned n_threads = 1)
      : path_(store_path), plan_(std::move(plan)), lib_{minLibH, maxLibH, minLibR, maxLibR}, min_len_(min_len), depth_(std::max<size_t>(depth, n_threads)) {
    for (unsigned t = 0; t < std::max(1u, n_threads); t++) th_.emplace_back([this] { run(); });
  }
  ~Prefetcher() { { std::lock_guard<std::mutex> lk(mu_); stop_ = true; } cv_.notify_all(); for (auto &t : th_) if (t.joinable()) t.join(); }
  //  hand a consumed batch back: its buffers (already faulted in, already big enough) are reused for a later item --
  //  first-touch page faults on fresh vectors cost 9x the packing itself (2.3 vs 20 Gbases/s measured)
  void recycle(std::unique_ptr<Packed> p) { if (!p) return; std::lock_guard<std::mutex> lk(mu_); free_.push_back(std::move(p)); }
  //  a recycled (or new) batch object for a caller that packs by itself (halves of an overflowed batch)
  std::unique_ptr<Packed> spare() {
    { std::lock_guard<std::mutex> lk(mu_); if (!free_.empty()) { auto p = std::move(free_.back()); free_.pop_back(); return p; } }
    return std::unique_ptr<Packed>(new Packed());
  }
  //  next item of the plan; nullptr + err on failure
  std::unique_ptr<Packed> next(std::string &err) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return ready_.count(head_) || !err_.empty() || head_ >= plan_.size(); });
    auto it = ready_.find(head_);
    if (it == ready_.end()) { err = err_.empty() ? "prefetcher: plan exhausted" : err_; return nullptr; }
    std::unique_ptr<Packed> p = std::move(it->second); ready_.erase(it); head_++;
    lk.unlock(); cv_.notify_all();
    return p;
  }

 private:
  void run() {
    SqStore st; std::string err;
    if (!st.open(path_.c_str(), err)) { fail(err); return; }
    for (;;) {
      size_t i;
      std::unique_ptr<Packed> p;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !err_.empty() || claim_ >= plan_.size() || claim_ < head_ + depth_; });
        if (stop_ || !err_.empty() || claim_ >= plan_.size()) return;
        i = claim_++;
        if (!free_.empty()) { p = std::move(free_.back()); free_.pop_back(); }
      }
      if (!p) p.reset(new Packed());
      const PackItem &it = plan_[i];
      const bool ok = it.is_hash ? pack_range(st, it.bgn, it.end, lib_[0], lib_[1], min_len_, *p, err)
                                 : pack_range(st, it.bgn, it.end, lib_[2], lib_[3], min_len_, *p, err);
      if (!ok) { fail(err); return; }
      { std::lock_guard<std::mutex> lk(mu_); ready_[i] = std::move(p); }
      cv_.notify_all();
    }
  }
  void fail(const std::string &e) { { std::lock_guard<std::mutex> lk(mu_); if (err_.empty()) err_ = e; } cv_.notify_all(); }

  std::string path_; std::vector<PackItem> plan_; uint32_t lib_[4]; uint32_t min_len_; size_t depth_;
  std::vector<std::thread> th_; std::mutex mu_; std::condition_variable cv_;
  std::map<size_t, std::unique_ptr<Packed>> ready_;
  size_t head_ = 0, claim_ = 0;                                          // next item the consumer takes / the packers start
  std::vector<std::unique_ptr<Packed>> free_;
  bool stop_ = false; std::string err_;
};

}  // namespace ovlhost
