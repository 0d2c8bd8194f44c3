#!/usr/bin/env python
"""Per-stage device timings of one tile for a synthetic workload (not a bench: for tuning).
usage: tools/stage_timing.py GENOME_BP COVERAGE READ_ERR MAXERATE [uniform LO HI | lognormal]"""
import sys, time
sys.path.insert(0, __file__.rsplit('/', 2)[0])
import numpy as np
from canu_b200 import api, synth

G, cov, err, erate = int(float(sys.argv[1])), float(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4])
g = synth.make_genome(G, seed=11)
if len(sys.argv) > 5 and sys.argv[5] == 'uniform':
    reads = synth.simulate_reads(g, cov, int(sys.argv[6]), int(sys.argv[7]), err, seed=12)
else:
    reads = synth.simulate_reads(g, cov, 3000, 30000, err, seed=12, lognormal=(9.25, 0.3))
prm = api.OverlapParams(kmer_len=22, max_erate=erate, min_olap_len=500, max_read_len=max(r.size for r in reads))
ov = api.Overlapper(prm)
pk = api.PackedReads(reads, first_read_id=1, min_len=500)
ov.load_hash_reads(pk); ov.build_index(); ov.stage_ref_batch(pk)
for i in range(3):
    ov.reset_counters()
    t0 = time.perf_counter(); ov.build_index(); n = ov.run_staged(); wall = time.perf_counter() - t0
    t = ov.timings(); c = ov.counters()
    print('wall %.1f ms' % (wall * 1e3), {k: round(v, 2) for k, v in t.items() if v > 0.005})
print('reads %d bases %d overlaps %d' % (len(reads), sum(r.size for r in reads), n), c)
print('Gcell/s in extend: %.2f   pairs/s overall: %.0f' % (c['dp_cells'] / t['extend_ms'] / 1e6, c['pairs'] / wall))
