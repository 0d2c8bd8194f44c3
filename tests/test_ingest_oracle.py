"""CPU: the numpy restatement of the store-ingest step (oracle/ingest_oracle.py) against goldens minted from the
reference's own ovStoreFilter / swapIDs / operator< (tests/golden/make_ingest_golden.py)."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
from oracle import ingest_oracle as io_

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "canu_b200", "bin", "ovltool")
CASES = json.load(open(os.path.join(gu.GOLDEN, "ingest.json")))


def load_ovb(name):
    lines = subprocess.check_output([TOOL, "dump-ovb", os.path.join(gu.GOLDEN, name)]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=io_.RECORD_DTYPE)
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    return recs


def load_golden(name):
    with gzip.open(os.path.join(gu.GOLDEN, name), "rb") as f:
        return np.frombuffer(f.read(), dtype=io_.RECORD_DTYPE)


@pytest.mark.parametrize("case", CASES, ids=[c["golden"] for c in CASES])
def test_ingest_oracle_matches_reference(case):
    recs = load_ovb(case["input"])
    n_reads = gu.load_cases()["stores"]["A"]["reads"]
    got = io_.ingest(recs, io_.encode_evalue(case["max_erate"]), n_reads)
    want = load_golden(case["golden"])
    assert len(want) == case["records"] and len(got) == len(want)
    for f in ("a_iid", "b_iid", "w0", "w1"):
        assert np.array_equal(got[f], want[f]), f


def test_swap_ids_is_an_involution_and_rejects_bad_ids():
    recs = load_ovb("A_default.ovb")
    twice = io_.swap_ids(io_.swap_ids(recs))
    assert np.array_equal(twice, recs)
    bad = recs.copy(); bad["b_iid"][0] = 0
    with pytest.raises(ValueError):
        io_.ingest(bad, 65535, 10 ** 6)


def _ingest_model():
    import ctypes as C
    src = os.path.join(ROOT, "tests", "model", "ingest_model.cc")
    so = os.path.join(ROOT, "tests", "model", "libingest_model.so")
    hdr = os.path.join(ROOT, "canu_b200", "csrc", "ovl_common.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    L = C.CDLL(so)
    L.ingest_model.restype = C.c_int64
    L.ingest_model.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p]
    return L


@pytest.mark.parametrize("case", CASES, ids=[c["golden"] for c in CASES])
def test_product_twin_code_matches_reference_on_cpu(case):
    """The host+device function the ingest kernel runs (ovl_common.cuh: ovl_ingest_twin), compiled for the CPU."""
    recs = np.ascontiguousarray(load_ovb(case["input"]))
    out = np.zeros(2 * recs.size, dtype=io_.RECORD_DTYPE)
    n = _ingest_model().ingest_model(recs.ctypes.data, recs.size, io_.encode_evalue(case["max_erate"]), out.ctypes.data)
    want = load_golden(case["golden"])
    assert n == len(want)
    for f in ("a_iid", "b_iid", "w0", "w1"):
        assert np.array_equal(out[:n][f], want[f]), f
