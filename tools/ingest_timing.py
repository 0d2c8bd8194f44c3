#!/usr/bin/env python
"""Throughput of the store-ingest step (mirror + filter + sort): ovlb_ingest_records through the
C ABI with HOST buffers (copies included) against the reference's own code (oracle/_ref/bin/ovsort_ref: ovFile reader,
ovStoreFilter::filterOverlap, std::sort with ovOverlap::operator<) on the same records, and the results compared.

    python tools/ingest_timing.py [N_RECORDS]          (default 20 M; needs a GPU and oracle/_ref)
Not the bench (bench.py is)."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canu_b200 import api, synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(9)
g = synth.make_genome(2_000_000, seed=1)
reads = synth.simulate_reads(g, 50, 1000, 1100, 0.0, seed=2)           # ~95 k short reads: only the read count matters here
max_id = len(reads)
wd = tempfile.mkdtemp(prefix="ingest_")
fa, st = os.path.join(wd, "r.fasta"), os.path.join(wd, "r.seqStore")
synth.write_fasta(fa, reads)
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/bin/sqStoreCreate"), "-o", st, "-minlength", "1000", "-pacbio-hifi", "lib", fa],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
recs = np.zeros(n, dtype=api.RECORD_DTYPE)
recs["a_iid"] = rng.integers(1, max_id + 1, n); recs["b_iid"] = rng.integers(1, max_id + 1, n)
h = lambda: rng.integers(0, 1 << 21, n, dtype=np.uint64)
recs["w0"] = h() | (h() << np.uint64(21)) | (rng.integers(0, 6000, n, dtype=np.uint64) << np.uint64(42)) | \
    (rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(58)) | (np.uint64(1) << np.uint64(61))
recs["w1"] = h() | (h() << np.uint64(21)) | (h() << np.uint64(42))
ov = api.Overlapper(api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500))
ov.ingest_records(recs[:1000], 4500, max_id)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); got = ov.ingest_records(recs, 4500, max_id); ts.append(time.perf_counter() - t0)
gpu_s = min(ts)
flat = os.path.join(wd, "in.bin"); recs.tofile(flat)
subprocess.check_call([os.path.join(ROOT, "canu_b200/bin/ovltool"), "pack-ovb", flat, os.path.join(wd, "in.ovb"), str(max_id)])
t0 = time.perf_counter()
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/bin/ovsort_ref"), st, os.path.join(wd, "in.ovb"), "0.045", os.path.join(wd, "out.bin")], stderr=subprocess.DEVNULL)
ref_s = time.perf_counter() - t0
want = np.fromfile(os.path.join(wd, "out.bin"), dtype=api.RECORD_DTYPE)
same = len(got) == len(want) and all(np.array_equal(got[f], want[f]) for f in ("a_iid", "b_iid", "w0", "w1"))
print(json.dumps({"records_in": n, "records_out": int(len(got)), "identical_to_reference": bool(same),
                  "gpu_s_with_copies": round(gpu_s, 3), "gpu_mrecords_in_per_s": round(n / gpu_s / 1e6, 1),
                  "reference_s": round(ref_s, 2), "reference_mrecords_in_per_s": round(n / ref_s / 1e6, 2),
                  "reference": "oracle/_ref/bin/ovsort_ref (reference ovFile reader + filterOverlap + std::sort, 1 thread)"}))
