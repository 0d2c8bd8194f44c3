#include "ovstore.h"

#include <cstdio>
#include <cstring>
#include <vector>
#include <sys/stat.h>

namespace ovlhost {

namespace {
struct Offt { uint16_t slice, piece; uint32_t offset, numOlaps, pad; uint64_t overlapID; };   // ovStoreOfft, 24 bytes
static_assert(sizeof(Offt) == 24, "ovStoreOfft layout");
struct Info { uint64_t magic, version; uint32_t readLenInBits, bgnID, endID, maxID; uint64_t numOlaps; };   // ovStoreInfo
static_assert(sizeof(Info) == 40, "ovStoreInfo layout");
const uint64_t kMaxPerPiece = 1024ull * 1024 * 1024 / (16 + 4);       // OVFILE_MAX_OVERLAPS
}  // namespace

bool write_ovstore(const std::string &path, uint32_t max_id, const ovlb_record *recs, uint64_t n, std::string &err) {
  if (mkdir(path.c_str(), 0777) != 0) {
    struct stat st;
    if (!(stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode))) { err = "cannot create directory '" + path + "'"; return false; }
  }
  std::vector<Offt> index((size_t)max_id + 1);
  memset(index.data(), 0, index.size() * sizeof(Offt));
  Info info; memset(&info, 0, sizeof(info));
  info.magic = 0x53564f3a756e6163ull; info.version = 4; info.readLenInBits = OVLB_MAX_READLEN_BITS;
  info.bgnID = UINT32_MAX; info.endID = 0; info.maxID = max_id;

  FILE *f = nullptr;
  uint32_t piece = 1; uint64_t in_piece = 0;
  std::vector<uint32_t> buf; buf.reserve(5 << 16);
  auto flush = [&]() -> bool { if (buf.empty()) return true; bool ok = fwrite(buf.data(), 4, buf.size(), f) == buf.size(); buf.clear(); return ok; };
  auto open_piece = [&]() -> bool {
    char name[64]; snprintf(name, sizeof(name), "/%04u-%03u", 1u, piece);
    f = fopen((path + name).c_str(), "wb");
    if (!f) { err = "cannot write '" + path + name + "'"; return false; }
    setvbuf(f, nullptr, _IOFBF, 1 << 22);
    return true;
  };
  for (uint64_t i = 0; i < n; i++) {
    const ovlb_record &r = recs[i];
    if (r.a_iid < 1 || r.a_iid > max_id || r.b_iid < 1 || r.b_iid > max_id) { err = "overlap names a read outside 1..maxID"; if (f) fclose(f); return false; }
    if (i && recs[i - 1].a_iid > r.a_iid) { err = "overlaps are not sorted"; if (f) fclose(f); return false; }
    if (f && in_piece > kMaxPerPiece && r.a_iid > info.endID) {          // ovStoreWriter.C:104-117
      if (!flush() || fclose(f) != 0) { err = "write failed"; return false; }
      f = nullptr; piece++; in_piece = 0;
    }
    if (!f && !open_piece()) return false;
    Offt &o = index[r.a_iid];
    if (o.numOlaps == 0) { o.slice = 1; o.piece = (uint16_t)piece; o.offset = (uint32_t)in_piece; o.overlapID = i; }
    o.numOlaps++;
    buf.push_back(r.b_iid);
    buf.push_back((uint32_t)(r.dat0 >> 32)); buf.push_back((uint32_t)r.dat0);
    buf.push_back((uint32_t)(r.dat1 >> 32)); buf.push_back((uint32_t)r.dat1);
    if (buf.size() >= (5u << 16) && !flush()) { err = "write failed"; fclose(f); return false; }
    in_piece++;
    if (r.a_iid < info.bgnID) info.bgnID = r.a_iid;
    if (r.a_iid > info.endID) info.endID = r.a_iid;
    info.numOlaps++;
  }
  if (f && (!flush() || fclose(f) != 0)) { err = "write failed"; return false; }
  FILE *x = fopen((path + "/index").c_str(), "wb");
  if (!x || fwrite(index.data(), sizeof(Offt), index.size(), x) != index.size() || fclose(x) != 0) { err = "cannot write '" + path + "/index'"; return false; }
  FILE *g = fopen((path + "/info").c_str(), "wb");
  if (!g || fwrite(&info, sizeof(info), 1, g) != 1 || fclose(g) != 0) { err = "cannot write '" + path + "/info'"; return false; }
  return true;
}

}  // namespace ovlhost
