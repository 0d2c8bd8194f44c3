#!/usr/bin/env python
"""Mint the golden parity fixtures from the UNMODIFIED reference binaries.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference, via
oracle/build_ref.sh -> oracle/_ref/bin):

    python tests/golden/make_golden.py

For each store it writes, under tests/golden/:
    <store>.fasta.gz              the seeded synthetic reads fed to sqStoreCreate
    <store>.seqStore/             the reference-made sqStore (input of the drop-in)
    <store>.dump.fasta.gz         reference `sqStoreDumpFASTQ -fasta` of the default read version
                                  (what overlapInCore sees; pins our sqStore reader)
and for each case (store x flags):
    <case>.ovl.txt.gz             sorted `overlapConvert -unaligned` of the reference .ovb
    <case>.stats                  reference -s file
    <case>.oc                     reference per-read overlap counts
    <case>.ovb                    (one small case only) the raw reference .ovb, to pin the ovb reader/writer
and cases.json describing all of it (flags, parameter values, -t used).

The reference drops the last ref read when a work chunk starts exactly on it
(SURVEY.md 7.5); -t is chosen per case so that quirk does not fire, and the
script double-checks by comparing against a second -t.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from canu_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
TMP = "/tmp/ovl_golden"


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-4000:])
        raise SystemExit("FAILED: " + " ".join(cmd))
    return r.stdout


def safe_threads(n_ref, prefer=(4, 3, 2, 5, 6, 7, 1)):
    """-t values for which the reference's last-ref-read drop does not fire."""
    ok = []
    for t in prefer:
        per = 1 + (n_ref - 1) // t // 8
        if (n_ref - 1) % per != 0:
            ok.append(t)
    return ok


def make_store(name, reads, create_flags):
    fa = os.path.join(TMP, name + ".fasta")
    synth.write_fasta(fa, reads)
    st = os.path.join(HERE, name + ".seqStore")
    shutil.rmtree(st, ignore_errors=True)
    run([os.path.join(BIN, "sqStoreCreate"), "-o", st, "-minlength", "1000"] + create_flags + ["lib", fa])
    for junk in ("errorLog", "info.txt", "readNames.txt", "load.dat"):
        p = os.path.join(st, junk)
        if os.path.exists(p):
            os.remove(p)
    os.chmod(os.path.join(st, "blobs.0001"), 0o644)
    with open(fa, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".fasta.gz"), "wb", mtime=0) as g:
        g.write(f.read())
    dump = run([os.path.join(BIN, "sqStoreDumpFASTQ"), "-S", st, "-fasta", "-o", "-"])
    with gzip.GzipFile(os.path.join(HERE, name + ".dump.fasta.gz"), "wb", mtime=0) as g:
        g.write(dump)
    n = sum(1 for line in dump.splitlines() if line.startswith(b">"))
    return st, n


def ref_overlaps(store, out_prefix, flags, t):
    ovb = out_prefix + ".ovb"
    for ext in (".ovb", ".oc", ".stats"):
        if os.path.exists(out_prefix + ext):
            os.remove(out_prefix + ext)
    run([os.path.join(BIN, "overlapInCore"), "-t", str(t)] + flags + ["-o", ovb, "-s", out_prefix + ".stats", store])
    txt = run([os.path.join(BIN, "overlapConvert"), "-S", store, "-unaligned", ovb])
    lines = sorted(txt.splitlines())
    return lines


def main():
    os.makedirs(TMP, exist_ok=True)
    if not os.path.exists(os.path.join(BIN, "overlapInCore")):
        raise SystemExit("run oracle/build_ref.sh first")

    cases = {"stores": {}, "cases": []}

    # ---- store A: ~1% error CLR-like, planted repeat, a few reads with N ----
    gA = synth.make_genome(60000, seed=11, repeat_len=2500, repeat_copies=6)
    rA = synth.simulate_reads(gA, 12, 1500, 5000, 0.01, seed=12, n_frac=0.003, n_reads_with_n=6)
    stA, nA = make_store("A", rA, ["-corrected", "-trimmed", "-pacbio"])
    cases["stores"]["A"] = {"reads": nA, "create": "-corrected -trimmed -pacbio", "genome": 60000, "err": 0.01}

    # skip k-mer list for store A: every 22-mer (canonical pairs counted per strand as-is) seen >= 30 times
    K = 22
    from collections import Counter
    cnt = Counter()
    for r in rA:
        s = r.tobytes()
        for i in range(len(s) - K + 1):
            cnt[s[i:i + K]] += 1
    skip = sorted(k for k, v in cnt.items() if v >= 30 and b"N" not in k)
    # add a handful of k-mers that are absent from the reads (exercise the "extra string" path)
    rng = np.random.default_rng(99)
    for _ in range(8):
        skip.append(bytes(b"ACGT"[i] for i in rng.integers(0, 4, size=K)))
    skip_path = os.path.join(HERE, "A.skip.dump")
    with open(skip_path, "wb") as f:
        for k in skip:
            f.write(k + b"\t%d\n" % cnt.get(k, 0))
    cases["stores"]["A"]["skip_kmers"] = len(skip)

    # ---- store B: ~3% error ----
    gB = synth.make_genome(50000, seed=21)
    rB = synth.simulate_reads(gB, 10, 1500, 4500, 0.03, seed=22)
    stB, nB = make_store("B", rB, ["-corrected", "-trimmed", "-pacbio"])
    cases["stores"]["B"] = {"reads": nB, "create": "-corrected -trimmed -pacbio", "genome": 50000, "err": 0.03}

    # ---- store C: HiFi-like, homopolymer-compressed store ----
    gC = synth.make_genome(60000, seed=31)
    rC = synth.simulate_reads(gC, 12, 2500, 7000, 0.002, seed=32)
    stC, nC = make_store("C", rC, ["-homopolycompress", "-pacbio-hifi"])
    cases["stores"]["C"] = {"reads": nC, "create": "-homopolycompress -pacbio-hifi", "genome": 60000, "err": 0.002}

    common = ["-k", "22", "--hashbits", "22", "--hashload", "0.8", "--minlength", "500"]

    def add(case, store_name, store, n, extra, hr=None, keep_ovb=False):
        hb, he, rb, re_ = hr if hr else (1, n, 1, n)
        flags = common + extra + ["-h", "%d-%d" % (hb, he), "-r", "%d-%d" % (rb, re_)]
        ts = safe_threads(re_ - rb + 1)
        t0, t1 = ts[0], ts[1]
        prefix = os.path.join(TMP, case)
        lines = ref_overlaps(store, prefix, flags, t0)
        stats0 = open(prefix + ".stats").read()
        shutil.copy(prefix + ".stats", os.path.join(HERE, case + ".stats"))
        shutil.copy(prefix + ".oc", os.path.join(HERE, case + ".oc"))
        if keep_ovb:
            shutil.copy(prefix + ".ovb", os.path.join(HERE, case + ".ovb"))
        lines2 = ref_overlaps(store, prefix + "_chk", flags, t1)
        stats1 = open(prefix + "_chk.stats").read()
        if "-partial" in extra:
            # Output_Partial_Overlap bumps the GLOBAL Total_Overlaps without a lock
            # (overlapInCore-Output.C:204; SURVEY.md 5 "known benign races"), so the
            # "Total overlaps produced" line can come out short under -t > 1.  Keep the
            # stats of a run where the race did not fire (total == records written).
            want = " Total overlaps produced = %d\n" % len(lines)
            tries = 0
            while want not in stats0:
                tries += 1
                if tries > 8:
                    raise SystemExit("case %s: could not get a race-free partial stats file" % case)
                ref_overlaps(store, prefix, flags, t0)
                stats0 = open(prefix + ".stats").read()
            shutil.copy(prefix + ".stats", os.path.join(HERE, case + ".stats"))
            stats1 = stats0
        if lines != lines2 or stats0 != stats1:
            raise SystemExit("case %s: reference output depends on -t (%d vs %d)" % (case, t0, t1))
        with gzip.GzipFile(os.path.join(HERE, case + ".ovl.txt.gz"), "wb", mtime=0) as g:
            g.write(b"\n".join(lines) + b"\n" if lines else b"")
        cases["cases"].append({"name": case, "store": store_name, "flags": extra, "h": [hb, he], "r": [rb, re_],
                               "threads": t0, "overlaps": len(lines), "ovb": keep_ovb})
        print("%-14s store %s  %6d overlaps  (-t %d/%d)  %s" % (case, store_name, len(lines), t0, t1, " ".join(extra)))

    add("A_default", "A", stA, nA, ["--maxerate", "0.045"], keep_ovb=True)
    add("A_partial", "A", stA, nA, ["--maxerate", "0.045", "-partial"])
    add("A_multi", "A", stA, nA, ["--maxerate", "0.045", "-m"])
    add("A_minkmers", "A", stA, nA, ["--maxerate", "0.045", "--minkmers"])
    add("A_skip", "A", stA, nA, ["--maxerate", "0.045", "-k", skip_path])
    add("A_nohopeless", "A", stA, nA, ["--maxerate", "0.045", "-z"])
    add("A_blocks", "A", stA, nA, ["--maxerate", "0.045", "--hashdatalen", "150000"])
    add("A_ranges", "A", stA, nA, ["--maxerate", "0.045"], hr=(60, nA - 20, 1, 150))
    add("A_hifi01", "A", stA, nA, ["--maxerate", "0.01"])
    add("B_e06", "B", stB, nB, ["--maxerate", "0.06"])
    add("B_e12", "B", stB, nB, ["--maxerate", "0.12"])
    add("B_e15_partial", "B", stB, nB, ["--maxerate", "0.15", "-partial"])
    add("C_hpc", "C", stC, nC, ["--maxerate", "0.01"])

    for c in cases["cases"]:
        c["flags"] = [("A.skip.dump" if f == skip_path else f) for f in c["flags"]]
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote", os.path.join(HERE, "cases.json"))


if __name__ == "__main__":
    main()
