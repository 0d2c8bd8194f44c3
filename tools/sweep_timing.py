#!/usr/bin/env python
"""BASELINE.json configs[3]: the error-rate / band-width sweep.  One tile per --maxerate e in 0.01 .. 0.15 (read error
e / 2, so the pairwise divergence sits at the threshold as in SURVEY.md 8d), same genome size, coverage and read-length
model; reports the extension kernel's Gcell/s, read-pairs/s and the work per pair.  Not the bench (bench.py is).

    python tools/sweep_timing.py [GENOME_BP] [--out profiles/rN_sweep.json]"""
import json
import sys
import time

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from canu_b200 import api, synth  # noqa: E402

G = int(float(sys.argv[1])) if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 400_000
out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else None
rows = []
g = synth.make_genome(G, seed=41)
for e in (0.01, 0.03, 0.045, 0.06, 0.09, 0.12, 0.15):
    reads = synth.simulate_reads(g, 40, 3000, 12000, e / 2, seed=42)
    prm = api.OverlapParams(kmer_len=22, max_erate=e, min_olap_len=500, max_read_len=max(r.size for r in reads))
    ov = api.Overlapper(prm)
    pk = api.PackedReads(reads, first_read_id=1, min_len=500)
    ov.load_hash_reads(pk); ov.build_index(); ov.stage_ref_batch(pk)
    ov.run_staged()
    ov.reset_counters()
    t0 = time.perf_counter(); ov.build_index(); n = ov.run_staged(); wall = time.perf_counter() - t0
    t = ov.timings(); c = ov.counters()
    ov.close()
    row = {"maxerate": e, "read_error": e / 2, "genome_bp": G, "reads": len(reads), "bases": int(sum(r.size for r in reads)),
           "pairs": c["pairs"], "overlaps": int(n), "extend_calls": c["extend_calls"], "dp_cells": c["dp_cells"],
           "cells_per_pair": round(c["dp_cells"] / max(c["pairs"], 1)), "calls_per_pair": round(c["extend_calls"] / max(c["pairs"], 1), 1),
           "step_ms": round(wall * 1e3, 1), "extend_ms": round(t["extend_ms"], 1), "seeding_ms": round(wall * 1e3 - t["extend_ms"], 1),
           "gcells_per_s": round(c["dp_cells"] / 1e9 / (t["extend_ms"] * 1e-3), 1), "read_pairs_per_s": round(c["pairs"] / wall)}
    rows.append(row)
    print(json.dumps(row), flush=True)
if out:
    json.dump(rows, open(out, "w"), indent=1)
