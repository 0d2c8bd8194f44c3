//  pack.h -- sqStore reads -> the ovlb_reads wire format (shared by the overlapInCore and ovlFrequentMers executables).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "../../include/ovlb200.h"
#include "sqstore.h"

namespace ovlhost {

//  Pack reads bgn..end (inclusive) of the store into the wire format.  Reads outside the library
//  filter or shorter than `min_len` keep their slot with length 0 (Build_Hash_Index.C:504-526,
//  Process_Overlaps.C:55-60).
struct Packed {
  std::vector<uint8_t> packed; std::vector<uint64_t> boff; std::vector<uint32_t> len, n_read, n_pos, src_len, clear_bgn;
  ovlb_reads view; uint64_t bases = 0;
  std::pair<const void *, size_t> pinned[5] = {};                          // page-locked ranges (set by the overlapInCore worker)
  //  A page-locked vector must not reallocate behind the registration's back (the freed range could be handed out again
  //  while the driver still maps the old pages): growth goes through here, which drops the registration first.
  template <class V> void ensure_cap(V &v, int slot, size_t need) {
    if (v.capacity() >= need) return;
    if (pinned[slot].first) { ovlb_host_unregister(pinned[slot].first); pinned[slot] = {nullptr, 0}; }
    v.reserve(need + need / 4 + 64);
  }
  ~Packed() { for (auto &p : pinned) if (p.first) ovlb_host_unregister(p.first); }
};

inline bool pack_range(SqStore &S, uint32_t bgn, uint32_t end, uint32_t min_lib, uint32_t max_lib, uint32_t min_len,
                       Packed &P, std::string &err) {
  const uint32_t n = end >= bgn ? end - bgn + 1 : 0;
  P.ensure_cap(P.boff, 1, n); P.ensure_cap(P.len, 2, n); P.ensure_cap(P.src_len, 3, n); P.ensure_cap(P.clear_bgn, 4, n);
  P.packed.clear(); P.boff.assign(n, 0); P.len.assign(n, 0); P.n_read.clear(); P.n_pos.clear(); P.bases = 0;
  P.src_len.assign(n, 0); P.clear_bgn.assign(n, 0);
  bool any_raw = false;
  std::string bases;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t id = bgn + i;
    P.boff[i] = P.packed.size();
    const uint32_t lib = S.libraryID(id);
    const uint32_t L = S.readLength(id);
    if (lib < min_lib || lib > max_lib || L < min_len) continue;
    P.ensure_cap(P.packed, 0, P.packed.size() + (S.storedLength(id) + 3) / 4 + 16);
    int fast = S.appendPacked2bit(id, P.packed, err);
    if (fast < 0) return false;
    if (fast == 0) {                                                      // homopolymer-compressed store, or a clear range that does not start
      uint32_t sl = 0, cb = 0;                                            // on a byte: the blob goes up as stored, the device prepares the read
      fast = S.appendRaw2bit(id, P.packed, sl, cb, err);
      if (fast < 0) return false;
      if (fast == 1) { P.src_len[i] = sl; P.clear_bgn[i] = cb; any_raw = true; }
    }
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
  P.ensure_cap(P.packed, 0, P.packed.size() + 8);
  P.packed.resize(P.packed.size() + 8, 0);
  P.view.packed = P.packed.data(); P.view.packed_bytes = P.packed.size() - 8;
  P.view.byte_offset = P.boff.data(); P.view.len = P.len.data(); P.view.n_reads = n; P.view.first_read_id = bgn;
  P.view.n_read = P.n_read.data(); P.view.n_pos = P.n_pos.data(); P.view.n_n = P.n_read.size();
  P.view.src_len = any_raw ? P.src_len.data() : nullptr; P.view.clear_bgn = any_raw ? P.clear_bgn.data() : nullptr;
  P.view.homopoly_compress = (S.version() & SQ_COMPRESSED) ? 1u : 0u;
  return true;
}


}  // namespace ovlhost
