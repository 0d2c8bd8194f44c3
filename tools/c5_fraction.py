#!/usr/bin/env python
"""BASELINE.json configs[4] (C5): 3.1 Gbp human-size genome, 30x HiFi-like reads, --maxerate 0.01 -- a stated FRACTION of the
job's tile grid, run through the drop-in executable, and the extrapolation to the whole grid (SURVEY.md 8d allows this:
generating 93 Gbases of reads is hours by itself).

The fraction: ONE hash block of the size the executable would pick on a B200 (~456 Mbases, 41 k reads) and `--ref-batches`
ref batches of 256 Mbases, all reads sampled uniformly from the SAME 3.1 Gbp random genome, so the pair density of a tile is
the real job's (a hash block covers 0.15x of the genome, a ref batch 0.08x: tiles are index build + lookup, hardly any
extension).  Ref reads get the low IDs (only refID < hashID pairs are computed).  The whole job is
n_blocks = total bases / hash block, and block i is crossed with the ref reads before it: n_tiles = sum_i (bases before block
i) / ref batch.  Extrapolation: T(1 GPU) = n_blocks x t(load + index) + n_tiles x t(tile), / 8 for whole hash blocks dealt
over 8 GPUs (no index is built twice).  Not a bench: bench.py is the bench.

    python tools/c5_fraction.py [--genome 3.1e9] [--coverage 30] [--ref-batches 4] [--out profiles/r2_c5_fraction.json]
"""
import argparse
import json
import multiprocessing as mp
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
OURS = os.path.join(ROOT, "canu_b200", "bin")
_G = None
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _part(a):
    from canu_b200 import synth
    i, bases, seed, path = a
    rng = np.random.default_rng(seed + 7919 * (i + 1))
    G = _G.size
    n = tot = 0
    with open(path, "wb") as f:
        while tot < bases:
            L = int(np.clip(rng.lognormal(9.25, 0.3), 3000, 30000))
            p = int(rng.integers(0, G - L + 1))
            r = _G[p:p + L]
            if rng.random() < 0.5:
                r = synth.revcomp(r)
            r = synth.inject_errors(r, 0.001, rng)
            f.write(b">p%d_%d\n" % (i, n)); f.write(np.ascontiguousarray(r).tobytes()); f.write(b"\n")
            n += 1; tot += L
    return n, tot


def main():
    global _G
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=float, default=3.1e9)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--hash-block", type=float, default=456e6)
    ap.add_argument("--ref-batch", type=float, default=256e6)
    ap.add_argument("--ref-batches", type=int, default=4)
    ap.add_argument("--procs", type=int, default=16)
    ap.add_argument("--store", default="/tmp/c5frac")
    ap.add_argument("--out", default="")
    ap.add_argument("--own-block", action="store_true", help="let the executable size the hash block itself (its memory model) instead of passing --hashblock")
    ap.add_argument("--ncu-launches", default="", help="third run under `ncu --metrics gpu__time_duration.sum`: the launch list goes to this CSV (not timed)")
    args = ap.parse_args()
    os.makedirs(args.store, exist_ok=True)
    t0 = time.perf_counter()
    Gn = int(args.genome)
    _G = np.empty(Gn, dtype=np.uint8)
    rng = np.random.default_rng(99)
    for s in range(0, Gn, 1 << 27):
        e = min(Gn, s + (1 << 27))
        _G[s:e] = _ACGT[rng.integers(0, 4, size=e - s, dtype=np.uint8)]
    t_gen = time.perf_counter() - t0
    # ref reads first (low IDs), then the hash block
    P = args.procs
    ref_bases = args.ref_batches * args.ref_batch
    jobs = [(i, ref_bases / P, 1, os.path.join(args.store, "ref%02d.fasta" % i)) for i in range(P)]
    jobs += [(P + i, args.hash_block / P, 2, os.path.join(args.store, "hash%02d.fasta" % i)) for i in range(P)]
    with mp.get_context("fork").Pool(min(2 * P, os.cpu_count() or 1)) as pool:
        res = pool.map(_part, jobs)
    n_ref = sum(r[0] for r in res[:P]); n_hash = sum(r[0] for r in res[P:])
    b_ref = sum(r[1] for r in res[:P]); b_hash = sum(r[1] for r in res[P:])
    del _G
    t_reads = time.perf_counter() - t0 - t_gen
    st = os.path.join(args.store, "c5.seqStore")
    subprocess.check_call([os.path.join(REF, "sqStoreCreate"), "-o", st, "-minlength", "1000", "-pacbio-hifi", "lib"] + [j[3] for j in jobs],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for j in jobs:
        os.remove(j[3])
    t_store = time.perf_counter() - t0 - t_gen - t_reads
    N = n_ref + n_hash
    rows = []
    for rep in range(3):                                   # later runs: CUDA start-up warm; the extrapolation uses the run with the fastest tiles
        cmd = [os.path.join(OURS, "overlapInCore"), "-k", "22", "--maxerate", "0.01", "--minlength", "500",
               "-h", "%d-%d" % (n_ref + 1, N), "-r", "1-%d" % n_ref] + ([] if args.own_block else ["--hashblock", str(int(args.hash_block * 1.2))]) + [
               "--refbatch", str(int(args.ref_batch)), "--gpu", "0", "-o", os.path.join(args.store, "x.ovb"), "-s", os.path.join(args.store, "x.stats"), st]
        t1 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True)
        wall = time.perf_counter() - t1
        log = r.stderr.decode()
        assert r.returncode == 0, log[-3000:]
        m = re.search(r"(\d+) overlaps, (\d+) candidate pairs, (\d+) DP cells", log)
        ph = [ln.strip() for ln in log.splitlines() if ln.strip().startswith("[gpu") and "create" in ln][0]
        f = {k: float(v) for k, v in re.findall(r"(create|pack-hash|load\+index|pack-ref|stage|run|fetch|submit)\s+([0-9.]+)", ph)}
        tiles = len([ln for ln in log.splitlines() if "Processed reads" in ln])
        rows.append(dict(wall_s=round(wall, 2), tiles=tiles, hash_blocks=len([ln for ln in log.splitlines() if "Build_Hash_Index" in ln]), overlaps=int(m.group(1)), pairs=int(m.group(2)), phases=f))
    if args.ncu_launches:
        subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", args.ncu_launches] + cmd,
                       capture_output=True)
    tile_s = lambda r: sum(r["phases"][k] for k in ("run", "stage", "fetch", "pack-ref", "submit"))
    best = min(rows, key=tile_s)
    f = best["phases"]; tiles = best["tiles"]
    total_bases = args.genome * args.coverage
    n_blocks = total_bases / b_hash
    n_tiles = sum((i * b_hash) / args.ref_batch for i in range(int(n_blocks)))          # block i meets the bases before it
    t_index = f["load+index"] + f["pack-hash"]
    t_tile = (f["run"] + f["stage"] + f["fetch"] + f["pack-ref"] + f["submit"]) / max(tiles, 1)
    T1 = n_blocks * t_index + n_tiles * t_tile
    out = {"workload": "C5 fraction: %.2f Gbp genome, %gx HiFi-like reads (log-normal ~11 kb, 0.1%% error), --maxerate 0.01" % (args.genome / 1e9, args.coverage),
           "fraction_run": {"hash_block_reads": n_hash, "hash_block_bases": b_hash, "ref_reads": n_ref, "ref_bases": b_ref, "tiles": tiles,
                            "of_the_grid": "1 of %.0f hash blocks, %d of %.0f tiles" % (n_blocks, tiles, n_tiles)},
           "prep_s": {"genome": round(t_gen, 1), "reads": round(t_reads, 1), "sqStoreCreate": round(t_store, 1)},
           "runs": rows,
           "per_hash_block_s": round(t_index, 3), "per_tile_s": round(t_tile, 4),
           "pairs_per_tile": best["pairs"] / max(tiles, 1),
           "extrapolation": {"hash_blocks": round(n_blocks), "tiles": round(n_tiles), "total_pairs": round(best["pairs"] / max(tiles, 1) * n_tiles),
                             "one_gpu_s": round(T1), "eight_gpus_s": round(T1 / 8),
                             "ref_upload_TB": round(n_tiles * args.ref_batch * 0.25 / 1e12, 2),
                             "note": "whole hash blocks dealt over the GPUs, no index built twice; host packing / PCIe of the ref batches is inside per_tile_s (pipelined with the previous tile's run)"}}
    print(json.dumps(out))
    if args.out:
        json.dump(out, open(os.path.join(ROOT, args.out), "w"), indent=1)


if __name__ == "__main__":
    main()
