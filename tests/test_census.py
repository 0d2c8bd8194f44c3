"""Skip-list census: the frequent k-mers Canu passes to `overlapInCore -k`.
CPU: the numpy restatement (oracle/census_oracle.py) against goldens minted from the REFERENCE meryl binary.
GPU: the CUDA census (ovlb_kmer_census) against the same goldens, and against the oracle on a fresh read set."""
import glob
import json
import os

import numpy as np
import pytest

import golden_util as gu
from oracle import census_oracle as co

CASES = sorted(glob.glob(os.path.join(gu.GOLDEN, "census_*.json")))


def _filter(flt):
    d, t = None, 0
    for f in flt:
        if f.startswith("distinct="):
            d = float(f.split("=")[1])
        if f.startswith("threshold="):
            t = int(f.split("=")[1])
    return d, t


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[7:-5])
def test_oracle_matches_reference_meryl(path):
    g = json.load(open(path))
    reads = gu.load_dump_reads(g["store"])
    d, t = _filter(g["filter"])
    keys, counts = co.canonical_counts(reads, g["K"])
    c2 = counts[counts >= 2]
    assert c2.size == g["statistics"]["distinct"] and int(c2.sum()) == g["statistics"]["present"]
    got, thr = co.frequent_kmers(reads, g["K"], d, t)
    assert thr == g["min_printed_count"] or (t and thr == t)
    assert got == g["kmers"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[7:-5])
def test_gpu_census_matches_reference_meryl(path):
    from canu_b200 import api
    g = json.load(open(path))
    reads = gu.load_dump_reads(g["store"])
    d, t = _filter(g["filter"])
    K = g["K"]
    ov = api.Overlapper(api.OverlapParams(kmer_len=K, max_erate=0.045, min_olap_len=0, max_read_len=max(r.size for r in reads)))
    pk = api.PackedReads(reads, first_read_id=1, min_len=0)
    ov.load_hash_reads(pk)
    for sb in (0, 2):                                   # one pass over k-mer space, and four slices
        keys, cnts, st = ov.kmer_census(distinct_fraction=-1.0 if d is None else d, min_count=t, slice_bits=sb)
        got = {co.text_canonical(co.key_to_text(k, K)): int(c) for k, c in zip(keys, cnts)}
        assert st["distinct"] == g["statistics"]["distinct"] and st["present"] == g["statistics"]["present"]
        assert got == g["kmers"], (sb, len(got), len(g["kmers"]))
    ov.close()


@pytest.mark.gpu
def test_gpu_census_on_reads_with_repeats_and_n():
    """A planted 40-copy repeat and reads with N: the census equals the oracle's, k-mers across an N are not counted."""
    from canu_b200 import api, synth
    g = synth.make_genome(150000, seed=5, repeat_len=3000, repeat_copies=40)
    reads = synth.simulate_reads(g, 15, 2000, 6000, 0.01, seed=6, n_frac=0.002, n_reads_with_n=50)
    K = 22
    ov = api.Overlapper(api.OverlapParams(kmer_len=K, max_erate=0.045, min_olap_len=0, max_read_len=max(r.size for r in reads)))
    ov.load_hash_reads(api.PackedReads(reads, first_read_id=1, min_len=0))
    keys, cnts, st = ov.kmer_census(distinct_fraction=0.98, min_count=0)
    want, thr = co.frequent_kmers(reads, K, 0.98, 0)
    got = {co.text_canonical(co.key_to_text(k, K)): int(c) for k, c in zip(keys, cnts)}
    assert st["threshold"] == thr and got == want and len(got) > 100
    ov.close()


@pytest.mark.gpu
def test_frequent_mers_executable_writes_the_reference_dump(tmp_path):
    """`ovlFrequentMers` on the reference-made store A: the file it writes holds the k-mers (either orientation) and
    counts of the reference `meryl print at-least distinct=0.90`, and overlapInCore accepts it as its -k file."""
    import subprocess
    root = os.path.dirname(gu.GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
    exe = os.path.join(root, "canu_b200", "bin", "ovlFrequentMers")
    g = json.load(open(os.path.join(gu.GOLDEN, "census_A_k22_d90.json")))
    out = str(tmp_path / "A.ms22.dump")
    r = subprocess.run([exe, "-k", "22", "-distinct", "0.90", "-o", out, os.path.join(gu.GOLDEN, "A.seqStore")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = {}
    for ln in open(out):
        k, c = ln.split()
        got[co.text_canonical(k)] = int(c)
    assert got == g["kmers"]
    ovl = os.path.join(root, "canu_b200", "bin", "overlapInCore")
    r = subprocess.run([ovl, "-k", "22", "-k", out, "--maxerate", "0.045", "--minlength", "500", "-h", "1-200", "-r", "1-200",
                        "-o", str(tmp_path / "x.ovb"), "-s", str(tmp_path / "x.stats"), os.path.join(gu.GOLDEN, "A.seqStore")], capture_output=True)
    assert r.returncode == 0 and ("Read %d kmers to mark to skip" % len(got)) in r.stderr.decode(), r.stderr.decode()[-1500:]
