//  TEST INFRASTRUCTURE -- not product code.  OUR driver around the REFERENCE's own store-ingest code, linked against the
//  reference objects built by oracle/build_ref.sh (nothing of the reference is copied here).
//
//  It performs exactly the in-memory phase of `ovStoreBuild` (stores/ovStoreBuild.C:176-263) with the reference's
//  functions: read every overlap of an .ovb with ovFile (ovFileFull), make the mirrored twin and apply the error filter with
//  ovStoreFilter::filterOverlap (-> ovOverlap::swapIDs, stores/ovOverlap.C:215-246), keep what still carries a
//  forUTG/forOBT/forDUP flag, std::sort with ovOverlap::operator< (stores/ovOverlap.H:265-279), and write the records as
//  flat little-endian { u32 a_iid, u32 b_iid, u64 dat[0], u64 dat[1] } -- the golden for ovlb_ingest_records.
//
//  usage: ovsort_ref <seqStore> <in.ovb> <maxErate> <out.bin>
#include "system.H"
#include "sqStore.H"
#include "ovStore.H"

#include <algorithm>
#include <vector>

int
main(int argc, char **argv) {
  if (argc != 5) {
    fprintf(stderr, "usage: %s seqStore in.ovb maxErate out.bin\n", argv[0]);
    return 1;
  }

  sqStore        *seq    = new sqStore(argv[1]);
  ovStoreFilter  *filter = new ovStoreFilter(seq, atof(argv[3]));
  ovFile         *in     = new ovFile(seq, argv[2], ovFileFull);

  std::vector<ovOverlap>  ovls;
  ovOverlap               f, r;

  while (in->readOverlap(&f)) {
    filter->filterOverlap(f, r);

    if (f.dat.ovl.forUTG || f.dat.ovl.forOBT || f.dat.ovl.forDUP)   ovls.push_back(f);
    if (r.dat.ovl.forUTG || r.dat.ovl.forOBT || r.dat.ovl.forDUP)   ovls.push_back(r);
  }

  std::sort(ovls.begin(), ovls.end());

  FILE *out = fopen(argv[4], "wb");
  if (out == NULL) {
    fprintf(stderr, "cannot write '%s'\n", argv[4]);
    return 1;
  }
  for (size_t i = 0; i < ovls.size(); i++) {
    uint32 ids[2] = { ovls[i].a_iid, ovls[i].b_iid };
    uint64 dat[2] = { ovls[i].dat.dat[0], ovls[i].dat.dat[1] };
    fwrite(ids, sizeof(uint32), 2, out);
    fwrite(dat, sizeof(uint64), 2, out);
  }
  fclose(out);

  fprintf(stderr, "ovsort_ref: %lu records (saved utg %lu obt %lu, high-error %lu)\n",
          (unsigned long)ovls.size(), (unsigned long)filter->savedUnitigging(), (unsigned long)filter->savedTrimming(), (unsigned long)filter->filteredErate());

  delete in;
  delete filter;
  delete seq;
  return 0;
}
