"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(canu_b200/libovlb200.so via canu_b200/api.py); the oracle is only the checker."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _api():
    import canu_b200.api as api
    if api.load_library().ovlb_device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return api


def _params(api, kw):
    return api.OverlapParams(kmer_len=kw.get("kmer_len", 22), max_erate=kw.get("max_erate", 0.06),
                             partial=kw.get("partial", False), unique=kw.get("unique", True),
                             min_olap_len=kw.get("min_olap_len", 0), no_hopeless=kw.get("no_hopeless", False),
                             min_kmers=kw.get("min_kmers", False))


def _diff_msg(got, want, limit=6):
    sg, sw = set(got), set(want)
    return "only GPU (%d): %s\nonly reference (%d): %s" % (
        len(sg - sw), sorted(sg - sw)[:limit], len(sw - sg), sorted(sw - sg)[:limit])


@pytest.mark.parametrize("case", gu.case_names())
def test_golden_case_bit_exact(case):
    """Whole tile vs the reference binary's output: overlaps, -s counters, .oc counts."""
    api = _api()
    c = gu.get_case(case)
    reads = gu.load_dump_reads(c["store"])
    kw, skip = gu.flags_to_kwargs(c["flags"])
    recs, ctr = api.overlap_in_core(reads, _params(api, kw), hash_range=tuple(c["h"]), ref_range=tuple(c["r"]),
                                    skip_kmers=gu.skip_kmers(skip) if skip else None)
    got, want = gu.format_records(recs), gu.load_golden_lines(case)
    assert got == want, _diff_msg(got, want)
    ok, exp = gu.stats_match(gu.load_golden_stats(case), ctr)
    assert ok, (exp, gu.load_golden_stats(case))
    n, opr = gu.oc_from_records(recs, len(reads))
    gn, gopr = gu.load_golden_oc(case)
    assert n == gn and np.array_equal(opr, gopr)


def test_small_ref_batches_give_identical_output():
    """Output must not depend on how the ref range is batched."""
    api = _api()
    c = gu.get_case("A_default")
    reads = gu.load_dump_reads("A")
    kw, _ = gu.flags_to_kwargs(c["flags"])
    r1, c1 = api.overlap_in_core(reads, _params(api, kw))
    r2, c2 = api.overlap_in_core(reads, _params(api, kw), ref_batch_bases=100000)
    assert gu.format_records(r1) == gu.format_records(r2)
    for k in ("kmer_hits_without_olap", "kmer_hits_with_olap", "total_overlaps", "contained", "dovetail"):
        assert c1[k] == c2[k]


@pytest.mark.parametrize("store,erate", [("A", 0.045), ("B", 0.12), ("C", 0.01)])
def test_seed_lists_match_oracle(store, erate):
    """K1-K3 at kernel granularity: candidate pairs and their ordered seed lists."""
    from oracle import oracle_py as op
    api = _api()
    reads = gu.load_dump_reads(store)
    o = op.Oracle(kmer_len=22, max_erate=erate, min_olap_len=500, hash_bits=20, hash_load=0.8, no_hopeless=True)
    o.set_reads(reads)
    o.run(threads=8, trace_pairs=True)
    pt, sd = o.pair_traces()
    ost = o.stats()
    o.close()
    want = {}
    for p in pt:
        b, n = int(p["seed_begin"]), int(p["n_seeds"])
        s = sd[b:b + n]
        want[(int(p["ref_id"]), int(p["dir"]), int(p["hash_id"]))] = (
            int(p["consistent"]), int(p["diag_ct"]), int(p["diag_bgn"]), int(p["diag_end"]),
            list(zip(s["start"].tolist(), s["offset"].tolist(), s["len"].tolist())))

    prm = api.OverlapParams(kmer_len=22, max_erate=erate, min_olap_len=500, no_hopeless=True,
                            max_read_len=max(len(r) for r in reads))
    ov = api.Overlapper(prm)
    pk = api.PackedReads(reads, first_read_id=1, min_len=500)
    ov.load_hash_reads(pk)
    ov.build_index()
    ov.stage_ref_batch(pk)
    ov.run_staged()
    pairs, seeds = ov.debug_pairs()
    got = {}
    for p in pairs:
        b, n = int(p["seed_begin"]), int(p["n_seeds"])
        s = seeds[b:b + n]
        got[(int(p["ref_id"]), int(p["dir"]), int(p["hash_id"]))] = (
            int(p["consistent"]) & 1, int(p["diag_ct"]), int(p["diag_bgn"]), int(p["diag_end"]),
            list(zip(s["start"].tolist(), s["offset"].tolist(), s["len"].tolist())))
    ctr = ov.counters()
    ov.close()
    assert set(got) == set(want), (len(got), len(want), sorted(set(got) ^ set(want))[:5])
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, (len(bad), bad[:3], [(got[k], want[k]) for k in bad[:1]])
    assert ctr["seed_hits"] == ost["seed_hits"]
    assert ctr["hash_kmers"] == ost["hash_inserts"]
    assert ctr["ref_kmers"] == ost["ref_lookups"]


@pytest.mark.parametrize("store,erate,partial", [("A", 0.045, False), ("B", 0.06, False), ("B", 0.15, True), ("C", 0.01, False)])
def test_extension_kernel_matches_oracle_calls(store, erate, partial):
    """K4 at kernel granularity: every Extend_Alignment call the oracle made, re-run on the GPU from
    the same seed: (S_Lo,S_Hi,T_Lo,T_Hi,Errors,kind,delta_ct) must be identical, and the DP cell count too."""
    from oracle import oracle_py as op
    api = _api()
    reads = gu.load_dump_reads(store)
    o = op.Oracle(kmer_len=22, max_erate=erate, min_olap_len=500, hash_bits=20, hash_load=0.8, partial=partial)
    o.set_reads(reads)
    o.run(threads=8, trace_exts=True)
    et = o.ext_traces()
    ost = o.stats()
    o.close()
    assert len(et) > 100

    prm = api.OverlapParams(kmer_len=22, max_erate=erate, min_olap_len=500, partial=partial,
                            max_read_len=max(len(r) for r in reads))
    ov = api.Overlapper(prm)
    pk = api.PackedReads(reads, first_read_id=1, min_len=500)
    ov.load_hash_reads(pk)
    ov.build_index()
    ov.stage_ref_batch(pk)
    ov.reset_counters()
    out, _ = ov.debug_extend(et["ref_id"] - 1, et["dir"], et["hash_id"] - 1, et["seed_start"], et["seed_offset"], et["seed_len"])
    ctr = ov.counters()
    ov.close()
    want = np.stack([et[f] for f in ("s_lo", "s_hi", "t_lo", "t_hi", "errors", "kind", "delta_ct")], axis=1)
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, "%d of %d extensions differ:\n%s" % (bad.size, len(et), "\n".join(
        "in %s\n  want %s\n  got  %s" % (et[i].tolist()[:6], want[i].tolist(), out[i].tolist()) for i in bad[:6]))
    assert ctr["dp_cells"] == ost["dp_cells"], (ctr["dp_cells"], ost["dp_cells"])
    assert ctr["extend_calls"] == ost["extend_calls"]


def test_delta_encoding_matches_oracle():
    """Left_Delta itself (only delta_ct reaches the output, but Lies_On_Alignment walks the values)."""
    from oracle import oracle_py as op
    api = _api()
    reads = gu.load_dump_reads("B")
    o = op.Oracle(kmer_len=22, max_erate=0.12, min_olap_len=500, hash_bits=20, hash_load=0.8)
    o.set_reads(reads)
    o.run(threads=8, trace_exts=True)
    et = o.ext_traces()
    sel = et[np.argsort(-et["delta_ct"], kind="stable")[:300]]
    prm = api.OverlapParams(kmer_len=22, max_erate=0.12, min_olap_len=500, max_read_len=max(len(r) for r in reads))
    ov = api.Overlapper(prm)
    pk = api.PackedReads(reads, first_read_id=1, min_len=500)
    ov.load_hash_reads(pk)
    ov.build_index()
    ov.stage_ref_batch(pk)
    stride = int(sel["delta_ct"].max()) + 4
    out, deltas = ov.debug_extend(sel["ref_id"] - 1, sel["dir"], sel["hash_id"] - 1, sel["seed_start"], sel["seed_offset"],
                                  sel["seed_len"], delta_stride=stride)
    ov.close()
    lower = [bytes(r).lower() for r in reads]
    comp = bytes.maketrans(b"acgtn", b"tgcan")
    for i, t in enumerate(sel):
        S = lower[t["ref_id"] - 1]
        if t["dir"]:
            S = S.translate(comp)[::-1]
        T = lower[t["hash_id"] - 1]
        res, d = o.extend_one(S, T, int(t["seed_start"]), int(t["seed_offset"]), int(t["seed_len"]))
        assert int(res["delta_ct"]) == int(out[i, 6]) == int(t["delta_ct"])
        assert np.array_equal(d, deltas[i, : len(d)]), (i, d[:10], deltas[i, :10])
    o.close()


def test_empty_and_degenerate_inputs():
    api = _api()
    prm = api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500)
    # no reads at all
    recs, ctr = api.overlap_in_core([], prm)
    assert len(recs) == 0 and ctr["total_overlaps"] == 0
    # reads all shorter than --minlength: nothing hashed, nothing searched
    short = [np.frombuffer(b"ACGT" * 50, dtype=np.uint8)] * 4
    recs, ctr = api.overlap_in_core(short, prm)
    assert len(recs) == 0
    # two identical reads: one contained overlap with zero hangs, zero error
    import canu_b200.synth as synth
    g = synth.make_genome(3000, 5)
    recs, ctr = api.overlap_in_core([g, g.copy()], prm)
    assert len(recs) == 1 and ctr["contained"] == 1
    f = recs[0]
    # equal hangs: Output_Overlap picks the hash read as A (Output.C:60-75); the oracle agrees: (2, 1)
    assert (f["a_iid"], f["b_iid"]) == (2, 1) and (f["w0"] & ((1 << 58) - 1)) == 0
    # an all-N read is stored but can never seed
    n_read = np.full(2000, ord("N"), dtype=np.uint8)
    recs, ctr = api.overlap_in_core([g, n_read, g.copy()], prm)
    assert len(recs) == 1 and (recs[0]["a_iid"], recs[0]["b_iid"]) == (3, 1)


@pytest.mark.parametrize("case,build_env", [("A_default", {}), ("A_skip", {}), ("A_partial", {}), ("A_ranges", {}), ("C_hpc", {}),
                                            ("A_default", {"OVLB_BUCKETED": "0"}), ("A_skip", {"OVLB_BUCKETED": "0"}),
                                            ("B_e06", {"OVLB_BUCKETED": "0"})])
def test_drop_in_executable_on_sqstore(case, build_env, tmp_path):
    """The C++ host driver end to end: same argv as the reference, reads the reference-made sqStore,
    writes .ovb/.oc/.stats; compared with the golden output of the reference binary.  OVLB_BUCKETED=0 forces the
    sorted index build (the fallback of the bucketed one) so that both builds stay pinned to the goldens."""
    import os
    import subprocess
    _api()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "canu_b200", "bin", "overlapInCore")
    tool = os.path.join(root, "canu_b200", "bin", "ovltool")
    assert os.path.exists(exe), "build the host driver first (__graft_entry__.build())"
    c = gu.get_case(case)
    store = os.path.join(gu.GOLDEN, c["store"] + ".seqStore")
    flags = [os.path.join(gu.GOLDEN, f) if f.endswith(".dump") else f for f in c["flags"]]
    ovb = str(tmp_path / "out.ovb")
    cmd = [exe, "-t", "4", "-k", "22", "--hashbits", "22", "--hashload", "0.8", "--minlength", "500"] + flags + [
        "-h", "%d-%d" % tuple(c["h"]), "-r", "%d-%d" % tuple(c["r"]), "-o", ovb, "-s", str(tmp_path / "out.stats"), store]
    r = subprocess.run(cmd, capture_output=True, env=dict(os.environ, **build_env))
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = subprocess.check_output([tool, "dump-ovb", ovb]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=[("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    got, want = gu.format_records(recs), gu.load_golden_lines(case)
    assert got == want, _diff_msg(got, want)
    assert open(str(tmp_path / "out.stats")).read() == open(os.path.join(gu.GOLDEN, case + ".stats")).read()
    assert open(str(tmp_path / "out.oc"), "rb").read() == open(os.path.join(gu.GOLDEN, case + ".oc"), "rb").read()


@pytest.mark.parametrize("extra", [["--gpus", "0,0", "--hashblock", "250000", "--refbatch", "120000"],
                                   ["--gpus", "0,0,0", "--hashblock", "60000", "--refbatch", "400000"],
                                   ["--gpus", "all"], ["--gpu", "0", "--streams", "2", "--refbatch", "300000"]])
def test_drop_in_executable_multi_worker(extra, tmp_path):
    """The multi-GPU path of the host driver (tile plan -> LPT owners -> one worker thread and context per
    device -> one writer): with a device listed several times the same code runs on a one-GPU box.  Output must
    be the golden single-tile output whatever the tiling."""
    import os
    import subprocess
    _api()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "canu_b200", "bin", "overlapInCore")
    tool = os.path.join(root, "canu_b200", "bin", "ovltool")
    c = gu.get_case("A_default")
    store = os.path.join(gu.GOLDEN, "A.seqStore")
    ovb = str(tmp_path / "out.ovb")
    cmd = [exe, "-t", "4", "-k", "22", "--hashbits", "22", "--hashload", "0.8", "--minlength", "500"] + c["flags"] + extra + [
        "-h", "1-220", "-r", "1-220", "-o", ovb, "-s", str(tmp_path / "out.stats"), store]
    r = subprocess.run(cmd, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = subprocess.check_output([tool, "dump-ovb", ovb]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=[("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    got, want = gu.format_records(recs), gu.load_golden_lines("A_default")
    assert got == want, _diff_msg(got, want)
    assert open(str(tmp_path / "out.stats")).read() == open(os.path.join(gu.GOLDEN, "A_default.stats")).read()
    assert open(str(tmp_path / "out.oc"), "rb").read() == open(os.path.join(gu.GOLDEN, "A_default.oc"), "rb").read()


def test_index_build_paths_agree_and_overflow_falls_back():
    """The bucketed index build is the default; a k-mer with more occurrences than a bucket holds (a 300-copy repeat
    at 30x) makes it fall back to the sorted build.  Both must give the oracle's overlaps; the repeat's k-mers are on
    the skip list (as Canu's meryl step would put them), which keeps the candidate pairs to the unique sequence but
    still sends every occurrence through the index build."""
    from oracle import oracle_py as op
    from canu_b200 import synth
    api = _api()
    K = 22

    def run(genome, skip_unit):
        reads = synth.simulate_reads(genome, 30, 1500, 4000, 0.01, seed=78)
        prm = api.OverlapParams(kmer_len=K, max_erate=0.045, min_olap_len=500, max_read_len=max(r.size for r in reads))
        skip = None
        if skip_unit is not None:
            skip = [skip_unit[i:i + K].tobytes().decode() for i in range(len(skip_unit) - K + 1)]   # both sides add the reverse complements
        ov = api.Overlapper(prm)
        pk = api.PackedReads(reads, first_read_id=1, min_len=500)
        ov.load_hash_reads(pk)
        if skip:
            ov.mark_skip_kmers(skip)
        ov.build_index()
        info = ov.debug_index_info()
        recs = ov.overlap_ref_batch(pk, cap=1 << 22)
        ov.close()
        o = op.Oracle(kmer_len=K, max_erate=0.045, min_olap_len=500, hash_bits=18, hash_load=0.8)
        o.set_reads(reads)
        if skip:
            o.set_skip_kmers(skip)
        want = op.sort_records(o.run(threads=8))
        got = np.sort(recs, order=["a_iid", "b_iid", "w0", "w1"])
        assert len(got) == len(want) and len(got) > 0
        for f in ("a_iid", "b_iid", "w0", "w1"):
            assert np.array_equal(got[f], want[f]), f
        return info

    plain = synth.make_genome(120000, seed=77)
    info = run(plain, None)
    assert info["bucketed"], info                      # 3.6 M tuples: the bucketed build handles it
    rep = synth.make_genome(300000, seed=79, repeat_len=300, repeat_copies=300)
    step = 300000 // 300
    unit = rep[step // 3: step // 3 + 300]
    info = run(rep, unit)
    assert not info["bucketed"], info                  # ~9000 occurrences of every repeat k-mer: a bucket overflowed


# ------------------------------------------------------------------------------------------------
#  store-ingest step: ovlb_ingest_records vs the reference-minted goldens and the numpy oracle
# ------------------------------------------------------------------------------------------------
def _ingest_cases():
    import json
    import os
    return json.load(open(os.path.join(gu.GOLDEN, "ingest.json")))


@pytest.mark.parametrize("case", _ingest_cases(), ids=[c["golden"] for c in _ingest_cases()])
def test_ingest_matches_reference_goldens(case):
    import gzip
    import os
    import subprocess
    from oracle import ingest_oracle as io_
    api = _api()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = subprocess.check_output([os.path.join(root, "canu_b200", "bin", "ovltool"), "dump-ovb",
                                     os.path.join(gu.GOLDEN, case["input"])]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=api.RECORD_DTYPE)
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    with gzip.open(os.path.join(gu.GOLDEN, case["golden"]), "rb") as f:
        want = np.frombuffer(f.read(), dtype=api.RECORD_DTYPE)
    ov = api.Overlapper(api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500))
    got = ov.ingest_records(recs, io_.encode_evalue(case["max_erate"]), gu.load_cases()["stores"]["A"]["reads"])
    ov.close()
    assert len(got) == len(want) == case["records"]
    for f in ("a_iid", "b_iid", "w0", "w1"):
        assert np.array_equal(got[f], want[f]), f


def test_ingest_random_records_ties_and_edges():
    """Two million random records (both orientations, all flag patterns, runs of equal (a, b) with different payloads)
    against the numpy oracle; sortedness and the mirror property checked on the result itself; empty input; everything
    filtered; an ID out of range fails like the reference."""
    from oracle import ingest_oracle as io_
    api = _api()
    rng = np.random.default_rng(5)
    n, max_id = 2_000_000, 300_000
    recs = np.zeros(n, dtype=api.RECORD_DTYPE)
    recs["a_iid"] = rng.integers(1, max_id + 1, n)
    recs["b_iid"] = rng.integers(1, max_id + 1, n)
    recs["a_iid"][: n // 50] = recs["a_iid"][n // 50: 2 * (n // 50)]          # repeated (a, b): ties resolved by the payload
    recs["b_iid"][: n // 50] = recs["b_iid"][n // 50: 2 * (n // 50)]
    h = lambda: rng.integers(0, 1 << 21, n, dtype=np.uint64)
    flags = rng.integers(0, 16, n, dtype=np.uint64)                          # flipped | obt | dup | utg
    recs["w0"] = h() | (h() << np.uint64(21)) | (rng.integers(0, 6000, n, dtype=np.uint64) << np.uint64(42)) | (flags << np.uint64(58))
    recs["w1"] = h() | (h() << np.uint64(21)) | (h() << np.uint64(42))
    ov = api.Overlapper(api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500))
    got = ov.ingest_records(recs, 4500, max_id)
    want = io_.ingest(recs, 4500, max_id)
    assert len(got) == len(want) and len(got) > n
    for f in ("a_iid", "b_iid", "w0", "w1"):
        assert np.array_equal(got[f], want[f]), f
    k = (got["a_iid"].astype(np.uint64) << np.uint64(32)) | got["b_iid"]
    assert np.all(k[1:] >= k[:-1])
    assert len(ov.ingest_records(recs[:0], 4500, max_id)) == 0
    assert len(ov.ingest_records(recs[:1000], 0, max_id)) == int(2 * np.sum(((recs["w0"][:1000] >> np.uint64(42)) & np.uint64(0xFFFF) == 0) &
                                                                            ((recs["w0"][:1000] >> np.uint64(59)) & np.uint64(7) != 0)))
    bad = recs[:10].copy(); bad["b_iid"][3] = max_id + 1
    with pytest.raises(api.OvlError):
        ov.ingest_records(bad, 4500, max_id)
    ov.close()


@pytest.mark.parametrize("K,expect_bucketed", [(17, True), (24, True), (28, False)])
def test_other_kmer_lengths_through_the_bucketed_build(K, expect_bucketed):
    """The index build mixes and buckets the k-mer inside its 2K+3 key bits: check a short and a long K against the
    oracle on a block large enough for the bucketed build.  The 8-byte bucket tuple holds the key bits below the bucket
    number next to the position, which always fits for K <= 24; K = 28 takes the sorted build."""
    from oracle import oracle_py as op
    from canu_b200 import synth
    api = _api()
    g = synth.make_genome(60000, seed=90 + K)
    reads = synth.simulate_reads(g, 12, 1500, 4000, 0.01, seed=91 + K)
    prm = api.OverlapParams(kmer_len=K, max_erate=0.045, min_olap_len=500, max_read_len=max(r.size for r in reads))
    ov = api.Overlapper(prm)
    pk = api.PackedReads(reads, first_read_id=1, min_len=500)
    ov.load_hash_reads(pk)
    ov.build_index()
    assert ov.debug_index_info()["bucketed"] == expect_bucketed
    recs = ov.overlap_ref_batch(pk, cap=1 << 20)
    ctr = ov.counters()
    ov.close()
    o = op.Oracle(kmer_len=K, max_erate=0.045, min_olap_len=500, hash_bits=18, hash_load=0.8)
    o.set_reads(reads)
    want = op.sort_records(o.run(threads=8))
    got = np.sort(recs, order=["a_iid", "b_iid", "w0", "w1"])
    assert len(got) == len(want) and len(got) > 0
    for f in ("a_iid", "b_iid", "w0", "w1"):
        assert np.array_equal(got[f], want[f]), f
    st = o.stats()
    assert ctr["kmer_hits_with_olap"] == st["kmer_hits_with_olap"] and ctr["kmer_hits_without_olap"] == st["kmer_hits_without_olap"]


@pytest.mark.parametrize("sorted_build", [False, True])
def test_hash_table_fingerprint_collisions_are_resolved_by_the_slot(monkeypatch, sorted_build):
    """The index hash table keeps a 32-bit fingerprint per entry, not the k-mer; the slot confirms a match.  With the
    fingerprint narrowed to 2 bits (test hook OVLB_HT_FPMASK) nearly every bucket holds different k-mers with the same
    fingerprint, so lookups must go on past false candidates -- records, counters and the skip-k-mer marking (which
    looks k-mers up and inserts absent ones) must not change.  A ref batch of other reads exercises the miss path."""
    from canu_b200 import synth
    api = _api()
    g = synth.make_genome(80000, seed=301, repeat_len=300, repeat_copies=12)
    hreads = synth.simulate_reads(g, 10, 1500, 4000, 0.01, seed=302)
    rreads = synth.simulate_reads(g, 6, 1500, 4000, 0.01, seed=303)
    rng = np.random.default_rng(304)
    skip = [bytes(b"ACGT"[j] for j in rng.integers(0, 4, 22)) for _ in range(500)]             # absent k-mers (appended to the index)
    skip += [hreads[3][i:i + 22].tobytes() for i in range(0, 1500, 7)]                         # and present ones (flagged)
    skip = sorted(set(k for k in skip if b"N" not in k))

    def run():
        if sorted_build:
            monkeypatch.setenv("OVLB_BUCKETED", "0")
        prm = api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500, max_read_len=max(r.size for r in hreads + rreads))
        ov = api.Overlapper(prm)
        ph = api.PackedReads(hreads, first_read_id=len(rreads) + 1, min_len=500)
        pr = api.PackedReads(rreads, first_read_id=1, min_len=500)
        ov.load_hash_reads(ph); ov.mark_skip_kmers(skip); ov.build_index()
        a = np.sort(ov.overlap_ref_batch(pr, cap=1 << 20), order=["a_iid", "b_iid", "w0", "w1"])
        c = ov.counters()
        ov.close()
        return a, c

    want, cw = run()
    drop = ("ext_busy_ns", "ext_capacity_ns")
    #  2-bit fingerprints; then a table filled to 92 % (hook OVLB_HT_PERCENT: nearly every bucket is full, keys sit several
    #  buckets past their home, a k-mer and its reverse complement in different buckets, the probe sequence wraps around the
    #  end of the table) -- the lookups, and what the forward pass tells the reverse pass not to look up, must not change
    for env in ({"OVLB_HT_FPMASK": "0x3"}, {"OVLB_HT_PERCENT": "92"}, {"OVLB_HT_FPMASK": "0x1", "OVLB_HT_PERCENT": "97"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got, cg = run()
        for k in env:
            monkeypatch.delenv(k)
        assert len(want) > 100 and len(got) == len(want), env
        for f in ("a_iid", "b_iid", "w0", "w1"):
            assert np.array_equal(got[f], want[f]), (env, f)
        assert {k: v for k, v in cg.items() if k not in drop} == {k: v for k, v in cw.items() if k not in drop}, env


@pytest.mark.parametrize("homopoly", [False, True])
def test_reads_prepared_on_the_device_match_reads_prepared_on_the_host(homopoly):
    """Row f3: sqStore blobs uploaded as stored, homopolymer compression (sequence-v1.C:203-261) and clear-range trimming
    (sqStore.H:397-413) done by k_encode_raw.  The same reads prepared with numpy and uploaded the ordinary way must
    give the same records and counters -- clear ranges are random (not byte aligned), a few reads are trimmed to nothing."""
    from canu_b200 import synth
    api = _api()
    g = synth.make_genome(50000, seed=41)
    raw = synth.simulate_reads(g, 14, 1500, 4500, 0.01, seed=42)
    rng = np.random.default_rng(43)
    final, clear = [], []
    for i, r in enumerate(raw):
        full = api.homopoly_compress(r) if homopoly else r
        b = int(rng.integers(0, 40)); e = full.size - int(rng.integers(0, 40))
        if i % 37 == 0:
            e = b + 100                                   # shorter than --minlength: dropped on both paths
        clear.append((b, e)); final.append(np.ascontiguousarray(full[b:e]))
    prm = api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500, max_read_len=max(r.size for r in raw))
    out = []
    for pk in (api.RawReads(raw, clear=clear, homopoly=homopoly, first_read_id=1, min_len=500),
               api.PackedReads(final, first_read_id=1, min_len=500)):
        ov = api.Overlapper(prm)
        ov.load_hash_reads(pk); ov.build_index()
        recs = np.sort(ov.overlap_ref_batch(pk, cap=1 << 20), order=["a_iid", "b_iid", "w0", "w1"])
        ctr = ov.counters(); ov.close()
        out.append((recs, {k: ctr[k] for k in ("pairs", "hash_kmers", "ref_kmers", "seed_hits", "dp_cells", "total_overlaps")}))
    assert len(out[0][0]) > 300
    assert out[0][0].tobytes() == out[1][0].tobytes()
    assert out[0][1] == out[1][1]


def test_seed_buffer_overflow_is_split_and_counted_once(tmp_path):
    """A ref batch that overflows the device's seed buffers fails with OVLB_ERR_CAPACITY; the executable cuts it in two
    and re-queues the halves (several levels deep here: OVLB_RUN_CAP forces a tiny buffer).  Records, .oc and -- because a
    failed run rolls its partial counters back -- the .stats file must still be the reference's golden output."""
    import os
    import subprocess
    api = _api()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "canu_b200", "bin", "overlapInCore")
    tool = os.path.join(root, "canu_b200", "bin", "ovltool")
    c = gu.get_case("A_default")
    store = os.path.join(gu.GOLDEN, "A.seqStore")
    ovb = str(tmp_path / "out.ovb")
    cmd = [exe, "-k", "22", "--minlength", "500"] + c["flags"] + ["-h", "1-220", "-r", "1-220", "-o", ovb, "-s", str(tmp_path / "out.stats"), store]
    r = subprocess.run(cmd, capture_output=True, env=dict(os.environ, OVLB_RUN_CAP="6000"))
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stderr.decode().count("Processed reads") >= 4            # the one planned batch really was split
    lines = subprocess.check_output([tool, "dump-ovb", ovb]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=[("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    got, want = gu.format_records(recs), gu.load_golden_lines("A_default")
    assert got == want, _diff_msg(got, want)
    assert open(str(tmp_path / "out.stats")).read() == open(os.path.join(gu.GOLDEN, "A_default.stats")).read()
    assert open(str(tmp_path / "out.oc"), "rb").read() == open(os.path.join(gu.GOLDEN, "A_default.oc"), "rb").read()
    # through the C ABI: the error code, then the same batch in halves gives the whole result
    os.environ["OVLB_RUN_CAP"] = "6000"
    try:
        reads = gu.load_dump_reads("A")
        prm = api.OverlapParams(kmer_len=22, max_erate=0.045, min_olap_len=500, max_read_len=max(r.size for r in reads))
        ov = api.Overlapper(prm)
        pk = api.PackedReads(reads, first_read_id=1, min_len=500)
        ov.load_hash_reads(pk); ov.build_index(); ov.stage_ref_batch(pk)
        with pytest.raises(api.OvlError) as ei:
            ov.run_staged()
        assert ei.value.code == -3
        assert ov.counters()["pairs"] == 0 and ov.counters()["kmer_hits_with_olap"] == 0      # rolled back
        ov.close()
    finally:
        del os.environ["OVLB_RUN_CAP"]


def test_midsize_job_against_the_live_reference_binary(tmp_path):
    """Driver-visible bench-scale parity: a mid-size job (0.25 Mbp x 30x, 2-8 kb reads, 1.5 % error, --maxerate 0.045:
    ~10 s for the reference on the box's cores) through the REFERENCE binary and through the drop-in executable on the
    same reference-made sqStore: canonically sorted records, .stats and .oc identical."""
    import os
    import subprocess
    from canu_b200 import synth
    _api()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "oracle", "_ref", "bin")
    if not os.path.exists(os.path.join(ref, "overlapInCore")):
        pytest.skip("oracle/_ref not built")
    ours = os.path.join(root, "canu_b200", "bin")
    g = synth.make_genome(250000, seed=77)
    reads = synth.simulate_reads(g, 30, 2000, 8000, 0.015, seed=78)
    fa, st = str(tmp_path / "r.fasta"), str(tmp_path / "r.seqStore")
    synth.write_fasta(fa, reads)
    subprocess.check_call([os.path.join(ref, "sqStoreCreate"), "-o", st, "-minlength", "1000", "-pacbio", "lib", fa],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    n = len(reads)
    cores = os.cpu_count() or 1
    t = next(t for t in range(cores, 0, -1) if (n - 1) % (1 + (n - 1) // t // 8) != 0)     # no last-ref-read drop
    common = ["-k", "22", "--hashbits", "22", "--hashload", "0.8", "--hashdatalen", str(10 ** 10), "--maxerate", "0.045",
              "--minlength", "500", "-h", "1-%d" % n, "-r", "1-%d" % n]
    for who, exe in (("ref", os.path.join(ref, "overlapInCore")), ("our", os.path.join(ours, "overlapInCore"))):
        r = subprocess.run([exe, "-t", str(t)] + common + ["-o", str(tmp_path / (who + ".ovb")), "-s", str(tmp_path / (who + ".stats")), st],
                           capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-1500:]
    c = subprocess.run([os.path.join(ours, "ovltool"), "cmp-ovb", str(tmp_path / "ref.ovb"), str(tmp_path / "our.ovb")], capture_output=True)
    assert c.returncode == 0, c.stdout.decode()[-600:]
    assert int(c.stdout.decode().split()[-7]) > 20000                                       # "first N second N ..."
    assert open(str(tmp_path / "ref.stats")).read() == open(str(tmp_path / "our.stats")).read()
    assert open(str(tmp_path / "ref.oc"), "rb").read() == open(str(tmp_path / "our.oc"), "rb").read()
