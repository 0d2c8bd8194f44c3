"""CPU checks of product host/device-shared logic (no GPU needed):

  * the run-level Add_Match replay of canu_b200/csrc/ovl_common.cuh (the code the chaining kernel
    runs, compiled here for the host by tests/model/chain_model.cc) reproduces the oracle's
    Match_Node lists -- order, lengths, `consistent`, diag stats -- for every candidate pair;
  * the product's host tables (Edit_Match_Limit etc., csrc/ovl_host.cc) equal the oracle's;
  * the C-ABI library loads and exports every symbol include/ovlb200.h declares.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle_py as op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chain_lib():
    src = os.path.join(ROOT, "tests", "model", "chain_model.cc")
    so = os.path.join(ROOT, "tests", "model", "libchain_model.so")
    hdr = os.path.join(ROOT, "canu_b200", "csrc", "ovl_common.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    L = C.CDLL(so)
    L.chain_model_run.restype = C.c_int64
    L.chain_model_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                  C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.chain_model_pairs.restype = C.c_void_p
    L.chain_model_seeds.restype = C.c_void_p
    L.chain_model_num_seeds.restype = C.c_int64
    return L


def _canon(pairs, seeds):
    """dict (ref,dir,hash) -> (consistent, diag_ct, diag_bgn, diag_end, [(start,offset,len)...])"""
    out = {}
    for p in pairs:
        b, n = int(p["seed_begin"]), int(p["n_seeds"])
        s = seeds[b:b + n]
        out[(int(p["ref_id"]), int(p["dir"]), int(p["hash_id"]))] = (
            int(p["consistent"]) & 1, int(p["diag_ct"]), int(p["diag_bgn"]), int(p["diag_end"]),
            list(zip(s["start"].tolist(), s["offset"].tolist(), s["len"].tolist())))
    return out


@pytest.mark.parametrize("store", ["A", "B", "C"])
def test_run_level_chain_matches_oracle_seed_lists(store):
    reads = gu.load_dump_reads(store)
    n = len(reads)
    o = op.Oracle(kmer_len=22, max_erate=0.06, min_olap_len=500, hash_bits=20, hash_load=0.8, no_hopeless=True)
    o.set_reads(reads)
    o.run(threads=4, trace_pairs=True)
    want = _canon(*o.pair_traces())
    o.close()

    L = _chain_lib()
    lens = np.array([r.size for r in reads], dtype=np.uint32)
    offs = np.zeros(n, dtype=np.uint64)
    offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    buf = np.ascontiguousarray(np.concatenate(reads))
    npairs = L.chain_model_run(buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, n, 22, 500, 1, n, 1, n)
    pairs = op._copy(L.chain_model_pairs(), npairs, op.PAIR_DTYPE)
    seeds = op._copy(L.chain_model_seeds(), L.chain_model_num_seeds(), op.SEED_DTYPE)
    got = _canon(pairs, seeds)

    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, (len(bad), bad[:3], [(got[k], want[k]) for k in bad[:1]])
    # the interesting cases must actually occur: multi-seed and inconsistent pairs
    assert any(len(v[4]) > 3 for v in want.values())
    assert any(v[0] == 0 for v in want.values())


@pytest.mark.parametrize("erate", [0.01, 0.045, 0.06, 0.0601, 0.12, 0.15])
def test_host_tables_equal_oracle(erate):
    import canu_b200.api as api
    L = api.load_library()
    p = api._Params()
    er = L.ovlb_parse_erate(repr(erate).encode())
    assert er == op.f32(erate)
    assert L.ovlb_params_init(C.byref(p), 22, er, 1.0, 0, 1, 500, 0, 1, 0) == 0
    o = op.Oracle(kmer_len=22, max_erate=erate, min_olap_len=500, min_kmers=True)
    assert p.n_edit_match_limit == o.max_errors()
    eml = np.ctypeslib.as_array(p.edit_match_limit, shape=(p.n_edit_match_limit,))
    assert np.array_equal(eml, o.edit_match_limit())
    assert p.branch_match_value == o.branch_match_value()
    assert p.min_branch_tail_slope == (1.0 if er > 0.06 else 0.2)
    assert p.use_hopeless_check == (0 if er > 0.06 else 1)
    L.ovlb_params_free(C.byref(p))
    o.close()


def test_library_exports_every_declared_symbol():
    import canu_b200.api as api
    L = api.load_library()
    hdr = open(os.path.join(ROOT, "include", "ovlb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ovlb_[a-z_0-9]+)\s*\(", hdr))
    assert declared and declared == set(api.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_pack_reads_wire_format_and_errors():
    import canu_b200.api as api
    reads = [np.frombuffer(b"ACGTNACGTTGCA", dtype=np.uint8), np.frombuffer(b"TTT", dtype=np.uint8),
             np.frombuffer(b"gattaca", dtype=np.uint8)]
    pr = api.PackedReads(reads, first_read_id=5, min_len=4)
    v = pr.view.contents
    assert v.n_reads == 3 and v.first_read_id == 5 and v.n_n == 1
    lens = np.ctypeslib.as_array(C.cast(v.len, C.POINTER(C.c_uint32)), shape=(3,))
    assert lens.tolist() == [13, 0, 7]                      # short read keeps its slot with len 0
    packed = np.ctypeslib.as_array(C.cast(v.packed, C.POINTER(C.c_uint8)), shape=(v.packed_bytes,))
    # sqStore 2-bit: 4 bases/byte, first base in the top bits, A0 C1 G2 T3 (N packed as A)
    assert packed[0] == (0 << 6 | 1 << 4 | 2 << 2 | 3) and packed[1] == (0 << 6 | 0 << 4 | 1 << 2 | 2)
    assert np.ctypeslib.as_array(C.cast(v.n_pos, C.POINTER(C.c_uint32)), shape=(1,))[0] == 4
    pr.close()
    with pytest.raises(api.OvlError):
        api.PackedReads([np.frombuffer(b"ACGTRYACGT", dtype=np.uint8)])


def test_no_cpu_fallback_without_device():
    import canu_b200.api as api
    if api.load_library().ovlb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.OvlError):
        api.Overlapper(api.OverlapParams(max_erate=0.045))
