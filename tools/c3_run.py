#!/usr/bin/env python
"""BASELINE.json configs[2] (C3) AT SIZE through the drop-in executable: 100 Mbp x 40x CLR-like reads, --maxerate 0.06.

    python tools/c3_run.py make   --genome 100e6 --store DIR          # genome + reads + sqStore (CPU; parallel generation)
    python tools/c3_run.py run    --store DIR --gpus all --out X.json  # canu_b200/bin/overlapInCore -h 1-N -r 1-N on the GPUs
    python tools/c3_run.py tiles  --store DIR --tiles T --outdir D     # our executable on T sampled tiles (-h a-b -r c-d), .ovb kept
    python tools/c3_run.py reftiles --store DIR --tiles T --outdir D   # the REFERENCE binary on the same tiles (CPU), compares

The store is a pure function of (--genome, --coverage, --seed, --procs): `make` gives the same reads here and on the GPU
box, so the reference tiles can run in the build container (CPU) and be compared with the GPU's .ovb files of the same
tiles.  Tiles are the ones Canu itself would cut (`overlapInCorePartition.C:176-255` via ovlb_plan_tiles with Canu's C3
block sizes -hl 160 Mbp -rl 5 Gbp, Configure.pm:585-602), sub-sampled to a ref range of `--tile-ref-reads` reads so that
the reference finishes in minutes.  Not a bench: bench.py is the bench."""
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
READ_ERR, ERATE, LEN_LO, LEN_HI = 0.03, "0.06", 10000, 20000
if os.environ.get("C3RUN_MODEL") == "C1":          # BASELINE configs[0] read model through the same tooling (pass --genome 4.6e6 --coverage 30)
    READ_ERR, ERATE, LEN_LO, LEN_HI = 0.01, "0.045", 3000, 15000

_G = None


def _part(a):
    from canu_b200 import synth
    i, cov, seed, path = a
    reads = synth.simulate_reads(_G, cov, LEN_LO, LEN_HI, READ_ERR, seed=seed + 1000 * (i + 1))
    with open(path, "wb") as f:
        for j, r in enumerate(reads):
            f.write(b">p%d_%d\n" % (i, j)); f.write(r.tobytes()); f.write(b"\n")
    return len(reads), int(sum(r.size for r in reads))


def make(args):
    global _G
    from canu_b200 import synth
    os.makedirs(args.store, exist_ok=True)
    st = os.path.join(args.store, "c3.seqStore")
    if os.path.exists(st):
        return st
    t0 = time.perf_counter()
    _G = synth.make_genome(int(args.genome), seed=args.seed)
    P = args.procs
    parts = [(i, args.coverage / P, args.seed, os.path.join(args.store, "part%02d.fasta" % i)) for i in range(P)]
    with mp.get_context("fork").Pool(min(P, os.cpu_count() or 1)) as pool:
        res = pool.map(_part, parts)
    t1 = time.perf_counter()
    subprocess.check_call([os.path.join(REF, "sqStoreCreate"), "-o", st, "-minlength", "1000", "-pacbio", "lib"] + [p[3] for p in parts],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for p in parts:
        os.remove(p[3])
    info = {"genome_bp": int(args.genome), "coverage": args.coverage, "reads": sum(r[0] for r in res), "bases": sum(r[1] for r in res),
            "gen_s": round(t1 - t0, 1), "store_s": round(time.perf_counter() - t1, 1)}
    json.dump(info, open(os.path.join(args.store, "info.json"), "w"))
    print("made", json.dumps(info), flush=True)
    return st


def store_info(args):
    return json.load(open(os.path.join(args.store, "info.json")))


def common_flags(n):
    return ["-k", "22", "--maxerate", ERATE, "--minlength", "500"]


def run(args):
    st = make(args)
    info = store_info(args)
    n = info["reads"]
    out = os.path.join(args.store, "job.ovb")
    cmd = [os.path.join(OURS, "overlapInCore")] + common_flags(n) + ["-h", "1-%d" % n, "-r", "1-%d" % n, "--gpus", args.gpus,
           "-o", out, "-s", os.path.join(args.store, "job.stats"), st]
    if args.hashblock:
        cmd[1:1] = ["--hashblock", str(int(args.hashblock))]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True)
    wall = time.perf_counter() - t0
    log = r.stderr.decode()
    assert r.returncode == 0, log[-3000:]
    m = re.search(r"(\d+) overlaps, (\d+) candidate pairs, (\d+) DP cells in ([0-9.]+) s", log)
    ovl, pairs, cells = int(m.group(1)), int(m.group(2)), int(m.group(3))
    tiles = re.search(r"(\d+) tile\(s\) on (\d+) context\(s\) of (\d+) GPU", log)
    h = subprocess.run([os.path.join(OURS, "ovltool"), "hash-ovb", out], capture_output=True).stdout.decode().strip()
    row = dict(info, gpus=int(tiles.group(3)) if tiles else None, tiles=int(tiles.group(1)) if tiles else None, wall_s=round(wall, 2),
               overlaps=ovl, read_pairs=pairs, dp_cells=cells, read_pairs_per_s=round(pairs / wall, 1),
               gcells_per_s=round(cells / 1e9 / wall, 1), records_hash=h,
               stats=open(os.path.join(args.store, "job.stats")).read().splitlines(),
               phases=[ln.strip() for ln in log.splitlines() if ln.startswith("phases") or (ln.strip().startswith("[gpu") and "create" in ln)])
    os.remove(out)
    print(json.dumps(row), flush=True)
    if args.out:
        prev = json.load(open(args.out)) if os.path.exists(args.out) else []
        prev.append(row)
        json.dump(prev, open(args.out, "w"), indent=1)


def sampled_tiles(args):
    """Canu's own tile grid for the store, `--tiles` of them picked evenly, each cut down to --tile-ref-reads ref reads."""
    from canu_b200 import api
    from canu_b200.host_util import store_read_lengths
    lens = store_read_lengths(os.path.join(args.store, "c3.seqStore"))
    tiles = api.plan_tiles(lens, 500, 160_000_000, 5_000_000_000, strict_reference=True)
    pick = [tiles[int((i + 0.5) * len(tiles) / args.tiles)] for i in range(args.tiles)] if len(tiles) > 1 else tiles
    out = []
    for k, t in enumerate(pick):
        span = t["ref_end"] - t["ref_bgn"] + 1
        rb = t["ref_bgn"] + (span // 3 if span > args.tile_ref_reads else 0)
        re_ = min(t["ref_end"], rb + args.tile_ref_reads - 1)
        out.append(dict(name="tile%d" % k, hash=(t["hash_bgn"], t["hash_end"]), ref=(rb, re_), canu_tile=(t["ref_bgn"], t["ref_end"]), n_tiles=len(tiles)))
    return out


def tiles_ours(args):
    make(args)
    os.makedirs(args.outdir, exist_ok=True)
    st = os.path.join(args.store, "c3.seqStore")
    rows = []
    for t in sampled_tiles(args):
        ovb = os.path.join(args.outdir, t["name"] + ".ours.ovb")
        cmd = [os.path.join(OURS, "overlapInCore")] + common_flags(0) + ["-h", "%d-%d" % t["hash"], "-r", "%d-%d" % t["ref"], "--gpu", "0",
               "-o", ovb, "-s", os.path.join(args.outdir, t["name"] + ".ours.stats"), st]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        rows.append(dict(t, ours_wall_s=round(time.perf_counter() - t0, 2)))
        print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(args.outdir, "tiles_ours.json"), "w"), indent=1)


def pick_threads(n_reads, cores):
    """A -t for which the reference does not drop the last ref read."""
    for t in range(cores, 0, -1):
        per = 1 + (n_reads - 1) // t // 8
        if (n_reads - 1) % per != 0:
            return t
    return 1


def tiles_ref(args):
    make(args)
    st = os.path.join(args.store, "c3.seqStore")
    cores = os.cpu_count() or 1
    rows = []
    for t in sampled_tiles(args):
        ovb = os.path.join(args.outdir, t["name"] + ".ref.ovb")
        nthr = pick_threads(t["ref"][1] - t["ref"][0] + 1, cores)
        cmd = [os.path.join(REF, "overlapInCore"), "-t", str(nthr), "-k", "22", "--hashbits", "25", "--hashload", "0.8", "--hashdatalen", str(10 ** 10),
               "--maxerate", ERATE, "--minlength", "500", "-h", "%d-%d" % t["hash"], "-r", "%d-%d" % t["ref"],
               "-o", ovb, "-s", os.path.join(args.outdir, t["name"] + ".ref.stats"), st]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        wall = time.perf_counter() - t0
        c = subprocess.run([os.path.join(OURS, "ovltool"), "cmp-ovb", ovb, os.path.join(args.outdir, t["name"] + ".ours.ovb")], capture_output=True)
        so = open(os.path.join(args.outdir, t["name"] + ".ours.stats")).read()
        sr = open(os.path.join(args.outdir, t["name"] + ".ref.stats")).read()
        rows.append(dict(t, ref_wall_s=round(wall, 1), ref_threads=nthr, records_identical=c.returncode == 0,
                         cmp=(c.stdout.decode().strip().splitlines() or ["?"])[-1], stats_identical=so == sr, stats=sr.splitlines()[:4]))
        print(json.dumps(rows[-1]), flush=True)
        os.remove(ovb)
    json.dump(rows, open(os.path.join(args.outdir, "tiles_ref.json"), "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["make", "run", "tiles", "reftiles"])
    ap.add_argument("--genome", type=float, default=100e6)
    ap.add_argument("--coverage", type=float, default=40.0)
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--procs", type=int, default=16, help="generator processes (part of the store's identity)")
    ap.add_argument("--store", default="/tmp/c3store")
    ap.add_argument("--gpus", default="all")
    ap.add_argument("--hashblock", type=float, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--tiles", type=int, default=2)
    ap.add_argument("--tile-ref-reads", type=int, default=66)
    ap.add_argument("--outdir", default="gpurun_out/c3tiles")
    args = ap.parse_args()
    {"make": make, "run": run, "tiles": tiles_ours, "reftiles": tiles_ref}[args.cmd](args)


if __name__ == "__main__":
    main()
