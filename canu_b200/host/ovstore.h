//  ovstore.h -- writer of Canu's overlap store (ovStore) from records that are already mirrored, filtered and sorted
//  (ovlb_ingest_records): the second half of SURVEY.md 8f row f2.
//
//  On-disk format followed (reference paths under /root/reference/src/stores):
//    ovStore.H:39-133          `info`: ovStoreInfo = {u64 magic "canu:OVS", u64 version 4, u32 readLenInBits 21,
//                              u32 bgnID, u32 endID, u32 maxID, u64 numOlaps}, 40 bytes
//    ovStore.H:140-168         `index`: ovStoreOfft[maxID + 1] = {u16 slice, u16 piece, u32 offset (overlaps into the
//                              piece), u32 numOlaps, u64 overlapID}, 24 bytes each with the compiler's padding
//    ovStoreFile.C:39-47       data files `SSSS-PPP` (slice, piece)
//    ovStoreFile.C:140-160     ovFileNormalWrite: NOT compressed (the store needs random access), records without a_iid:
//                              5 x uint32 = b_iid, hi32(dat0), lo32(dat0), hi32(dat1), lo32(dat1)  (:371-395)
//    ovStoreWriter.C:25-170    sequential writer: one slice, a new piece once the current one holds more than
//                              OVFILE_MAX_OVERLAPS (ovStoreFile.H:27) overlaps and the read changes
//  The `statistics` file (ovStoreHistogram, only read by ovStoreStats-type reports) is not written.
#pragma once

#include <cstdint>
#include <string>

#include "../../include/ovlb200.h"

namespace ovlhost {

//  recs[0..n) sorted by (a_iid, b_iid, dat0, dat1) with every a_iid, b_iid in 1..max_id.  Creates the directory.
bool write_ovstore(const std::string &path, uint32_t max_id, const ovlb_record *recs, uint64_t n, std::string &err);

}  // namespace ovlhost
