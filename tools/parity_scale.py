#!/usr/bin/env python
"""Parity at bench scale, through the drop-in boundary: for each case make a synthetic read set, build a sqStore
with the reference's own sqStoreCreate, run the UNMODIFIED reference `overlapInCore -t <cores>` and our
`canu_b200/bin/overlapInCore` on the same store with the same flags, and compare: canonical-sorted .ovb records,
.oc file bytes, .stats lines.  Also times both (wall clock of the whole process, store already on disk).

    python tools/parity_scale.py [--cases S1,S2,...] [--out profiles/rN_parity_scale.json] [--gpus 0]

Needs oracle/_ref/bin (reference binaries, built by oracle/build_ref.sh) and a GPU."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
OURS = os.path.join(ROOT, "canu_b200", "bin")

#  name: (genome bp, coverage, per-read error, --maxerate, length model, extra flags, sqStoreCreate tech)
CASES = {
    "S1": (2_000_000, 30, 0.01, "0.045", ("uniform", 3000, 15000), [], "-pacbio"),          # C1-like
    "S2": (2_000_000, 50, 0.001, "0.01", ("lognormal", 9.25, 0.3), [], "-pacbio-hifi"),     # C2-like (the bench workload)
    "S3": (600_000, 40, 0.03, "0.06", ("uniform", 3000, 20000), [], "-pacbio"),             # C3-like
    "S4a": (200_000, 40, 0.045, "0.09", ("uniform", 3000, 12000), [], "-pacbio"),           # C4 sweep
    "S4b": (200_000, 40, 0.06, "0.12", ("uniform", 3000, 12000), [], "-pacbio"),
    "S4c": (150_000, 40, 0.075, "0.15", ("uniform", 3000, 12000), [], "-pacbio"),
    "S5": (1_000_000, 30, 0.01, "0.045", ("uniform", 3000, 15000), ["-partial"], "-pacbio"),  # obt mode
    # planted 4 kb x 40-copy repeat + its k-mers as the -k skip list (Mark_Skip_Kmers, hopeless check on)
    "S6": (1_000_000, 30, 0.01, "0.045", ("uniform", 3000, 15000), ["SKIP"], "-pacbio"),
    "C1full": (4_600_000, 30, 0.01, "0.045", ("uniform", 3000, 15000), [], "-pacbio"),     # BASELINE configs[0] at full size
    "C2full": (5_000_000, 50, 0.001, "0.01", ("lognormal", 9.25, 0.3), [], "-pacbio-hifi"),  # BASELINE configs[1] at full size
    "S7": (800_000, 30, 0.02, "0.06", ("uniform", 3000, 15000), ["SKIP", "--minkmers"], "-pacbio"),
}


def run(cmd, **kw):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, **kw)
    return time.perf_counter() - t0, r


def pick_threads(n_reads, cores):
    """A -t for which the reference does not drop the last ref read."""
    for t in range(cores, 0, -1):
        per = 1 + (n_reads - 1) // t // 8
        if (n_reads - 1) % per != 0:
            return t
    return 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(k for k in CASES if not k.endswith("full")))
    ap.add_argument("--out", default="")
    ap.add_argument("--gpus", default="0")
    args = ap.parse_args()
    from canu_b200 import synth
    cores = os.cpu_count() or 1
    report = []
    for name in args.cases.split(","):
        G, cov, err, erate, lm, extra, tech = CASES[name]
        wd = tempfile.mkdtemp(prefix="ovlparity_")
        try:
            use_skip = "SKIP" in extra
            extra = [x for x in extra if x != "SKIP"]
            g = synth.make_genome(G, seed=11, repeat_len=4000 if use_skip else 0, repeat_copies=40 if use_skip else 0)
            if lm[0] == "uniform":
                reads = synth.simulate_reads(g, cov, lm[1], lm[2], err, seed=12)
            else:
                reads = synth.simulate_reads(g, cov, 3000, 30000, err, seed=12, lognormal=(lm[1], lm[2]))
            fa, st = os.path.join(wd, "r.fasta"), os.path.join(wd, "r.seqStore")
            synth.write_fasta(fa, reads)
            subprocess.check_call([os.path.join(REF, "sqStoreCreate"), "-o", st, "-minlength", "1000", tech, "lib", fa],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            n = len(reads)
            if use_skip:
                # what `meryl print` would list: every k-mer of the repeat unit (high count), one per line with a count
                step = G // 40
                unit = g[step // 3: step // 3 + 4000].tobytes().decode()
                skipf = os.path.join(wd, "skip.dump")
                with open(skipf, "w") as f:
                    for i in range(0, len(unit) - 22 + 1):
                        f.write("%s\t%d\n" % (unit[i:i + 22], 40 * cov))
                extra = ["-k", skipf] + extra
            t = pick_threads(n, cores)
            common = ["-k", "22", "--hashbits", "23", "--hashload", "0.8", "--hashdatalen", str(10 ** 10), "--maxerate", erate,
                      "--minlength", "500", "-h", "1-%d" % n, "-r", "1-%d" % n] + extra
            t_ref, r = run([os.path.join(REF, "overlapInCore"), "-t", str(t)] + common +
                           ["-o", os.path.join(wd, "ref.ovb"), "-s", os.path.join(wd, "ref.stats"), st])
            assert r.returncode == 0, r.stderr.decode()[-1500:]
            gp = ["--gpus", args.gpus] if "," in args.gpus or args.gpus == "all" else ["--gpu", args.gpus]
            t_our, r = run([os.path.join(OURS, "overlapInCore"), "-t", str(t)] + common + gp +
                           ["-o", os.path.join(wd, "our.ovb"), "-s", os.path.join(wd, "our.stats"), st])
            assert r.returncode == 0, r.stderr.decode()[-1500:]
            _, c = run([os.path.join(OURS, "ovltool"), "cmp-ovb", os.path.join(wd, "ref.ovb"), os.path.join(wd, "our.ovb")])
            cmp_line = c.stdout.decode().strip().splitlines()[-1] if c.stdout else "cmp failed: " + c.stderr.decode()
            stats_ref = open(os.path.join(wd, "ref.stats")).read()
            stats_our = open(os.path.join(wd, "our.stats")).read()
            oc_same = open(os.path.join(wd, "ref.oc"), "rb").read() == open(os.path.join(wd, "our.oc"), "rb").read()
            vals = {k.strip(): int(v) for k, v in (ln.split("=") for ln in stats_ref.splitlines())}
            pairs = vals["Kmer hits without olaps"] + vals["Kmer hits with olaps"]
            row = {"case": name, "genome_bp": G, "coverage": cov, "read_error": err, "maxerate": erate, "flags": extra,
                   "reads": n, "bases": int(sum(r_.size for r_ in reads)), "ref_threads": t, "host_cores": cores,
                   "overlaps": vals["Total overlaps produced"], "read_pairs": pairs,
                   "records_identical": c.returncode == 0, "cmp": cmp_line, "stats_identical": stats_ref == stats_our,
                   "oc_identical": oc_same, "ref_wall_s": round(t_ref, 2), "ours_wall_s": round(t_our, 2),
                   "ref_pairs_per_s": round(pairs / t_ref, 1), "ours_pairs_per_s": round(pairs / t_our, 1),
                   "speedup_wall": round(t_ref / t_our, 1)}
            report.append(row)
            print(json.dumps(row), flush=True)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(report, f, indent=1)
    bad = [r for r in report if not (r["records_identical"] and r["stats_identical"] and r["oc_identical"])]
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
