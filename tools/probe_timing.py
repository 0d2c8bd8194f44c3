import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from canu_b200 import api, synth
import bench
reads = bench.make_workload(5_000_000, 50.0, seed=2001)
prm = api.OverlapParams(kmer_len=22, max_erate=0.01, min_olap_len=500, max_read_len=max(r.size for r in reads))
ov = api.Overlapper(prm)
pk = api.PackedReads(reads, first_read_id=1, min_len=500)
ov.load_hash_reads(pk); ov.build_index(); ov.stage_ref_batch(pk)
for i in range(3):
    ov.run_staged(); t = ov.timings(); print('run only  ', {k: round(v, 2) for k, v in t.items() if k in ('probe_ms','expand_ms','extend_ms','total_ms')})
for i in range(3):
    ov.build_index(); ov.run_staged(); t = ov.timings(); print('build+run ', {k: round(v, 2) for k, v in t.items() if k in ('index_sort_ms','index_table_ms','probe_ms','expand_ms','extend_ms','total_ms')})
