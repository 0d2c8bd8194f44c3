#!/usr/bin/env python
"""Mint golden tile plans from the UNMODIFIED reference `overlapInCorePartition` (oracle/_ref/bin, built by
oracle/build_ref.sh) on the committed golden stores.  Run in the build container only (needs oracle/_ref):

    python tests/golden/make_partition_golden.py

Writes tests/golden/partition.json: for every (store, -hl, -rl, -ol) the lines of the .ovlopt file."""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "overlapInCorePartition")

CASES = [("A", 200000, 400000, 500), ("A", 100000, 100000, 500), ("A", 50000, 1000000, 500), ("A", 1000000, 50000, 500),
         ("A", 1000000, 1000000, 500), ("A", 30000, 30000, 4000), ("B", 150000, 90000, 500), ("B", 40000, 40000, 1),
         ("C", 120000, 250000, 500), ("C", 20000, 700000, 500)]


def main():
    out = []
    for store, hl, rl, ol in CASES:
        with tempfile.TemporaryDirectory() as d:
            subprocess.check_call([BIN, "-S", os.path.join(HERE, store + ".seqStore"), "-hl", str(hl), "-rl", str(rl),
                                   "-ol", str(ol), "-o", os.path.join(d, "p")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = [ln.strip() for ln in open(os.path.join(d, "p.ovlopt")) if ln.strip()]
        out.append({"store": store, "hl": hl, "rl": rl, "ol": ol, "ovlopt": lines})
    with open(os.path.join(HERE, "partition.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
