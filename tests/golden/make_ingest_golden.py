#!/usr/bin/env python
"""Mint the goldens of the store-ingest step from the REFERENCE's own code.

TEST INFRASTRUCTURE.  Needs /root/reference via oracle/build_ref.sh -> oracle/_ref/bin/{overlapInCore,ovsort_ref}.
`ovsort_ref` is our small driver (oracle/ref_shim/ovsort_ref.C) around the reference's ovFile reader,
ovStoreFilter::filterOverlap (-> ovOverlap::swapIDs) and ovOverlap::operator<, i.e. the in-memory phase of ovStoreBuild.

    python tests/golden/make_ingest_golden.py

Writes under tests/golden/:
    A_partial_small.ovb                 reference `overlapInCore -partial -h 1-120 -r 1-120` on A.seqStore (obt flags)
    ingest_<input>_e<erate>.bin.gz      sorted + mirrored + filtered records, flat {u32 a, u32 b, u64 dat0, u64 dat1}
    ingest.json                         the list of (input .ovb, max error rate, golden file, record count)
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
STORE = os.path.join(HERE, "A.seqStore")


def main():
    tmp = tempfile.mkdtemp(prefix="ingest_golden_")
    small = os.path.join(HERE, "A_partial_small.ovb")
    subprocess.check_call([os.path.join(BIN, "overlapInCore"), "-partial", "-t", "3", "-k", "22", "--hashbits", "22", "--hashload", "0.8",
                           "--maxerate", "0.045", "--minlength", "500", "-h", "1-120", "-r", "1-120",
                           "-o", os.path.join(tmp, "p.ovb"), "-s", os.path.join(tmp, "p.stats"), STORE],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.replace(os.path.join(tmp, "p.ovb"), small)
    cases = []
    for inp, erate in (("A_default.ovb", "1.0"), ("A_default.ovb", "0.02"), ("A_partial_small.ovb", "1.0"), ("A_partial_small.ovb", "0.03")):
        out = os.path.join(tmp, "o.bin")
        subprocess.check_call([os.path.join(BIN, "ovsort_ref"), STORE, os.path.join(HERE, inp), erate, out], stderr=subprocess.DEVNULL)
        data = open(out, "rb").read()
        name = "ingest_%s_e%s.bin.gz" % (inp.replace(".ovb", ""), erate.replace(".", ""))
        with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as f:
            f.write(data)
        cases.append({"input": inp, "max_erate": float(erate), "golden": name, "records": len(data) // 24})
        print(name, len(data) // 24)
    json.dump(cases, open(os.path.join(HERE, "ingest.json"), "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
