//  sqstore.h -- read-only reader of Canu's sqStore (the input side of the ovl path).
//
//  On-disk format followed (reference paths under /root/reference/src):
//    stores/sqStoreInfo.C:266-299          `info`: IFF objects MAGC VERS LSIZ RSIZ "MLB " "LNS " "MRB " MRLB
//                                          NLIB NREA NBLO and arrays READ, BASE (16 x uint64 each)
//    stores/sqStoreConstructor.C:118-200   metadata files + default read version selection
//    stores/sqRead.H:168-331               sqReadSeq (3 x uint32 bit-fields), sqReadMeta (2 x uint64)
//    stores/sqStore.H:397-413              read length rules (0 if invalid / ignored / untrimmed)
//    stores/sqReadData.C:27-100,           BLOB{NAME, 2SQR|3SQR|USQR, 2SQC|3SQC|USQC} chunk walk
//    stores/sqCache.C:94-155,262-318
//    utility/src/files/buffered-v1-writing.C:109-160   IFF chunk = 4-byte tag, uint32 padded length, data
//    utility/src/sequence/sequence-v1.C:203-261,268-474   homopolymer compression, 2-bit / 3-bit codecs
//
//  Nothing is written; the store is never modified.
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ovlhost {

enum : uint32_t { SQ_RAW = 0x1, SQ_CORRECTED = 0x2, SQ_COMPRESSED = 0x4, SQ_TRIMMED = 0x8 };

struct SqReadSeq {                 // one 12-byte record of reads-rawu / -rawc / -coru / -corc
  uint32_t w0, w1, w2;
  bool     valid()    const { return w0 & 1u; }
  uint32_t length()   const { return w0 >> 2; }
  bool     ignoreU()  const { return w1 & 1u; }
  bool     ignoreT()  const { return (w1 & 1u) | ((w1 >> 1) & 1u); }
  uint32_t clearBgn() const { return w1 >> 2; }
  bool     trimmed()  const { return w2 & 1u; }
  uint32_t clearEnd() const { return w2 >> 2; }
};

class SqStore {
 public:
  ~SqStore();
  bool open(const std::string &path, std::string &err);

  uint32_t lastReadID() const { return num_reads_; }
  uint32_t version()    const { return which_; }          // SQ_* flags of the default read version
  uint32_t libraryID(uint32_t id) const { return (uint32_t)((meta_[2 * (size_t)id] >> 30) & 0xfff); }
  uint32_t readLength(uint32_t id) const;                 // 0 if the read is not usable
  uint32_t storedLength(uint32_t id) const {              // bases of the stored (uncompressed, untrimmed) sequence
    return ((which_ & SQ_RAW) ? rawu_[id] : coru_[id]).length();
  }

  //  The read as the overlapper sees it (default version: trimming / homopolymer compression applied),
  //  upper-case ASCII.  Returns false (and sets err) on a corrupt store.
  bool loadRead(uint32_t id, std::string &bases, std::string &err);

  //  Fast path for the common case (2-bit blob, not compressed, clear range starting on a byte):
  //  appends the read's packed bytes (sqStore's own encoding, 4 bases/byte MSB first) to `packed`
  //  without decoding.  Returns 1 if done, 0 if the caller must use loadRead(), -1 on error.
  int appendPacked2bit(uint32_t id, std::vector<uint8_t> &packed, std::string &err);

  //  Second fast path: the read's WHOLE 2-bit blob as stored (src_len bases, before homopolymer compression and
  //  trimming) for the device to prepare (ovlb_reads.src_len / clear_bgn / homopoly_compress).  Returns 1 if done,
  //  0 if the blob is not 2-bit (reads with N are stored 3-bit: the caller decodes those with loadRead()), -1 on error.
  int appendRaw2bit(uint32_t id, std::vector<uint8_t> &packed, uint32_t &src_len, uint32_t &clear_bgn, std::string &err);

 private:
  bool readFile(const std::string &name, std::vector<uint8_t> &out, std::string &err) const;
  bool fetchChunk(uint32_t id, const uint8_t *&chunk, uint32_t &chunk_len, char &enc, std::string &err);
  const SqReadSeq &seq(uint32_t id) const;

  std::string path_;
  uint32_t num_reads_ = 0, num_libs_ = 0, num_blobs_ = 0;
  uint64_t reads_by_version_[16] = {0};
  uint32_t which_ = 0;
  std::vector<uint64_t> meta_;                            // 2 x uint64 per read (index 0 unused)
  std::vector<SqReadSeq> rawu_, rawc_, coru_, corc_;
  struct BlobMap { const uint8_t *p = nullptr; size_t n = 0; };
  std::vector<BlobMap> blob_maps_;                        // blobs.NNNN, memory-mapped read-only on first use
};

uint32_t homopolyCompress(const std::string &in, std::string &out);   // sequence-v1.C:203-261

}  // namespace ovlhost
