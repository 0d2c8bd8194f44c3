#!/usr/bin/env python
"""Strong scaling of ONE ovl job over the GPUs of a node, through the drop-in executable and the real file formats.

For each case: make a synthetic read set (BASELINE.json C3-like / C2-like read models), build a sqStore with the
reference's sqStoreCreate, then run `canu_b200/bin/overlapInCore -h 1-N -r 1-N` on the SAME store with --gpus 0 /
0,1 / 0,1,2,3 ... (hash blocks x ref batches are planned inside the process, tiles are assigned to GPUs longest-first,
SURVEY.md 8e; no collective).  Every multi-GPU output must be record-identical to the 1-GPU output after canonical sort
(`ovltool cmp-ovb`), with identical .oc bytes and .stats lines: the tile grid and the merge may not change the result.
Reports wall time of the whole process, candidate pairs/s, DP Gcell/s and the efficiency T1 / (N x TN).

    python tools/scale_run.py [--cases noisy,hifi] [--gpu-counts 1,2,4] [--out profiles/rN_scale.json]

Needs oracle/_ref/bin/sqStoreCreate (built by oracle/build_ref.sh) and GPUs.  Not a bench: bench.py is the bench."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
OURS = os.path.join(ROOT, "canu_b200", "bin")

#  name: (genome bp, coverage, per-read error, --maxerate, length model, sqStoreCreate tech, --hashblock bases)
CASES = {
    "noisy": (8_000_000, 40, 0.03, "0.06", ("uniform", 10000, 20000), "-pacbio", 80_000_000),       # C3 read model, 4 hash blocks
    "hifi": (20_000_000, 50, 0.001, "0.01", ("lognormal", 9.25, 0.3), "-pacbio-hifi", 250_000_000),  # C2 read model, 4 hash blocks
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="noisy,hifi")
    ap.add_argument("--gpu-counts", default="1,2,4")
    ap.add_argument("--out", default="")
    ap.add_argument("--scale", type=float, default=1.0, help="multiply the genome sizes (and hash blocks)")
    ap.add_argument("--streams", default="", help="pass --streams N to the executable")
    args = ap.parse_args()
    from canu_b200 import synth
    report = []
    ok = True
    for name in args.cases.split(","):
        G, cov, err, erate, lm, tech, hb = CASES[name]
        G = int(G * args.scale); hb = int(hb * args.scale)
        wd = tempfile.mkdtemp(prefix="ovlscale_")
        try:
            t0 = time.perf_counter()
            g = synth.make_genome(G, seed=21)
            if lm[0] == "uniform":
                reads = synth.simulate_reads(g, cov, lm[1], lm[2], err, seed=22)
            else:
                reads = synth.simulate_reads(g, cov, 3000, 30000, err, seed=22, lognormal=(lm[1], lm[2]))
            fa, st = os.path.join(wd, "r.fasta"), os.path.join(wd, "r.seqStore")
            synth.write_fasta(fa, reads)
            n = len(reads)
            bases = int(sum(r.size for r in reads))
            del reads, g
            subprocess.check_call([os.path.join(REF, "sqStoreCreate"), "-o", st, "-minlength", "1000", tech, "lib", fa],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.remove(fa)
            prep_s = time.perf_counter() - t0
            base = None
            for ng in [int(x) for x in args.gpu_counts.split(",")]:
                tag = "g%d" % ng
                cmd = [os.path.join(OURS, "overlapInCore"), "-k", "22", "--maxerate", erate, "--minlength", "500",
                       "-h", "1-%d" % n, "-r", "1-%d" % n, "--hashblock", str(hb),
                       "--gpus", ",".join(str(i) for i in range(ng))] + (["--streams", args.streams] if args.streams else []) + [
                       "-o", os.path.join(wd, tag + ".ovb"), "-s", os.path.join(wd, tag + ".stats"), st]
                t1 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True)
                wall = time.perf_counter() - t1
                log = r.stderr.decode()
                assert r.returncode == 0, log[-2000:]
                m = re.search(r"(\d+) overlaps, (\d+) candidate pairs, (\d+) DP cells in ([0-9.]+) s", log)
                ovl, pairs, cells = int(m.group(1)), int(m.group(2)), int(m.group(3))
                tiles = re.search(r"(\d+) tile\(s\) on (\d+) context", log)
                phases = [ln.strip() for ln in log.splitlines() if ln.startswith("phases") or ln.strip().startswith("[gpu") and "create" in ln]
                row = {"case": name, "genome_bp": G, "coverage": cov, "read_error": err, "maxerate": erate, "reads": n, "bases": bases,
                       "hash_block_bases": hb, "gpus": ng, "tiles": int(tiles.group(1)) if tiles else None,
                       "wall_s": round(wall, 2), "overlaps": ovl, "read_pairs": pairs, "dp_cells": cells,
                       "read_pairs_per_s": round(pairs / wall, 1), "gcells_per_s": round(cells / 1e9 / wall, 2), "phases": phases}
                if base is None:
                    base = (tag, wall)
                    row["identical_to_1gpu"] = None
                    row["efficiency"] = 1.0
                else:
                    c = subprocess.run([os.path.join(OURS, "ovltool"), "cmp-ovb", os.path.join(wd, base[0] + ".ovb"), os.path.join(wd, tag + ".ovb")],
                                       capture_output=True)
                    same = (c.returncode == 0 and
                            open(os.path.join(wd, base[0] + ".stats")).read() == open(os.path.join(wd, tag + ".stats")).read() and
                            open(os.path.join(wd, base[0] + ".oc"), "rb").read() == open(os.path.join(wd, tag + ".oc"), "rb").read())
                    row["identical_to_1gpu"] = same
                    row["cmp"] = c.stdout.decode().strip().splitlines()[-1] if c.stdout else c.stderr.decode()[-300:]
                    row["efficiency"] = round(base[1] / (ng * wall), 3)
                    ok = ok and same
                    os.remove(os.path.join(wd, tag + ".ovb"))
                row["prep_s"] = round(prep_s, 1)
                report.append(row)
                print(json.dumps(row), flush=True)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(report, f, indent=1)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
