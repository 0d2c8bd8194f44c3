#!/usr/bin/env python
"""Mint the goldens of the skip-list census from the REFERENCE meryl binary.

TEST INFRASTRUCTURE.  Needs oracle/_ref/bin/meryl (oracle/build_ref.sh).  For each golden store and parameter set it runs
exactly what Canu's meryl scripts run (src/pipelines/canu/Meryl.pm:529-533, 603-607, 663-671):

    meryl k=K count <store> output db ; meryl greater-than 1 output db2 union-sum db ; meryl print at-least ... db2

and stores {orientation-free k-mer text: count} plus the `statistics` header numbers in tests/golden/census_<case>.json."""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.census_oracle import text_canonical  # noqa: E402

MERYL = os.path.join(ROOT, "oracle", "_ref", "bin", "meryl")
CASES = [  # name, store, K, filter args of `meryl print`
    ("A_k22_d90", "A", 22, ["at-least", "distinct=0.90"]),
    ("A_k22_t12", "A", 22, ["at-least", "threshold=12"]),
    ("A_k17_d99", "A", 17, ["at-least", "distinct=0.99"]),
    ("B_k22_d9999", "B", 22, ["at-least", "distinct=0.9999"]),
    ("C_k22_d95", "C", 22, ["at-least", "distinct=0.95"]),
]


def main():
    tmp = "/tmp/census_golden"
    for name, store, K, flt in CASES:
        shutil.rmtree(tmp, ignore_errors=True); os.makedirs(tmp)
        st = os.path.join(HERE, store + ".seqStore")
        run = lambda a: subprocess.run([MERYL] + a, cwd=tmp, check=True, capture_output=True)
        run(["k=%d" % K, "threads=2", "memory=2", "count", st, "output", "db.meryl"])
        run(["threads=2", "memory=2", "greater-than", "1", "output", "db2.meryl", "union-sum", "db.meryl"])
        out = run(["threads=1", "memory=2", "print"] + flt + ["db2.meryl"]).stdout.decode()
        stats = run(["statistics", "db2.meryl"]).stdout.decode()
        nums = {}
        for ln in stats.splitlines():
            f = ln.split()
            if len(f) >= 2 and f[0] in ("unique", "distinct", "present") and f[1].isdigit():
                nums[f[0]] = int(f[1])
        kmers = {}
        for ln in out.splitlines():
            k, c = ln.split()
            kmers[text_canonical(k)] = int(c)
        json.dump({"store": store, "K": K, "filter": flt, "statistics": nums, "min_printed_count": min(kmers.values()) if kmers else None,
                   "kmers": kmers}, open(os.path.join(HERE, "census_%s.json" % name), "w"), indent=0, sort_keys=True)
        print(name, len(kmers), nums, "min count", min(kmers.values()) if kmers else None)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
