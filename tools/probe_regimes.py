#!/usr/bin/env python
"""Probe stage (`k_ref_probe`) timed in the two regimes the executable meets:

  self     : ref batch == hash block (a bacterial job, the first tile of every hash block): every window is in the table and
             the path-ordered slots are read coalesced;
  dense    : ref batch == hash block at 50x coverage (bench.py's C2 tile): both strands of every window hit;
  sparse   : hash block and ref batch are DIFFERENT reads sampled uniformly from a genome much larger than either (the tile
             grid of a human-size job, BASELINE configs[4]): ~85 % of the windows are not in the table, the rest hit in
             stretches.

    python tools/probe_regimes.py [--genome 400e6] [--hash-bases 60e6] [--ref-bases 120e6]

Prints one JSON line per regime: probe ms, windows/s, random 128-byte lines/s ceiling it compares with (tools/micro/rand_sector.cu:
36.9 G lines/s).  Not a bench: bench.py is the bench."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canu_b200 import api, synth  # noqa: E402


def sample(G, bases, seed):
    return synth.simulate_reads(G, bases / G.size, 3000, 30000, 0.001, seed=seed, lognormal=(9.25, 0.3))


def time_probe(hash_reads, ref_reads, first_ref_id, first_hash_id, reps=4):
    prm = api.OverlapParams(kmer_len=22, max_erate=0.01, min_olap_len=500, max_read_len=max(r.size for r in hash_reads + ref_reads))
    ov = api.Overlapper(prm)
    ph = api.PackedReads(hash_reads, first_read_id=first_hash_id, min_len=500)
    pr = ph if ref_reads is hash_reads else api.PackedReads(ref_reads, first_read_id=first_ref_id, min_len=500)
    ov.load_hash_reads(ph); ov.build_index(); ov.stage_ref_batch(pr)
    best = None
    for _ in range(reps):
        n = ov.run_staged(); t = ov.timings()
        row = {k: round(t[k], 3) for k in ("probe_ms", "expand_ms", "extend_ms", "total_ms")}
        if best is None or row["probe_ms"] < best["probe_ms"]:
            best = row
    c = ov.counters()
    windows = 2 * sum(int(r.size) for r in ref_reads)                        # both orientations
    best.update(overlaps=n, windows=windows, seed_hits=c["seed_hits"] // reps, pairs=c["pairs"] // reps)
    best["gwindows_per_s"] = round(windows / (best["probe_ms"] * 1e-3) / 1e9, 2)
    ov.close()
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=float, default=400e6, help="0 = skip the sparse and self regimes")
    ap.add_argument("--hash-bases", type=float, default=60e6)
    ap.add_argument("--ref-bases", type=float, default=120e6)
    ap.add_argument("--dense-genome", type=float, default=5e6, help="also time the C2 tile (this genome x 50x, ref == hash); 0 = skip")
    args = ap.parse_args()
    if args.genome > 0:
        G = synth.make_genome(int(args.genome), seed=5)
        ref = sample(G, args.ref_bases, 1)
        hsh = sample(G, args.hash_bases, 2)
        r = time_probe(hsh, ref, 1, len(ref) + 1)
        print(json.dumps(dict(regime="sparse", hash_reads=len(hsh), ref_reads=len(ref), **r)), flush=True)
        del G
        r = time_probe(hsh, hsh, 1, 1)
        print(json.dumps(dict(regime="self", hash_reads=len(hsh), **r)), flush=True)
        del hsh, ref
    if args.dense_genome > 0:                                              # the C2 tile of bench.py: every window hits, both strands
        G = synth.make_genome(int(args.dense_genome), seed=2001)
        d = synth.simulate_reads(G, 50.0, 3000, 30000, 0.001, seed=2002, lognormal=(9.25, 0.3))
        r = time_probe(d, d, 1, 1)
        print(json.dumps(dict(regime="dense", hash_reads=len(d), **r)), flush=True)


if __name__ == "__main__":
    main()
