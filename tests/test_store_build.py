"""Overlap-store writer (second half): a store written by canu_b200 must read back through the
REFERENCE's own ovStore class exactly like the store the reference's ovStoreBuild makes from the same .ovb.

CPU: oracle ingest (numpy) -> `ovltool write-store` (the product's writer) -> reference ovStoreDump, against the dump of
a reference-built store.  GPU: the `ovlStoreBuild` executable end to end (ovlb_ingest_records on the device)."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
from oracle import ingest_oracle as io_

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
OURS = os.path.join(ROOT, "canu_b200", "bin")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ovStoreDump")), reason="oracle/_ref/bin/ovStoreDump not built")


def _load_ovb(path):
    lines = subprocess.check_output([os.path.join(OURS, "ovltool"), "dump-ovb", path]).decode().splitlines()
    recs = np.zeros(len(lines), dtype=io_.RECORD_DTYPE)
    for i, ln in enumerate(lines):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    return recs


def _dump(store, seq, what):
    """The reference ovStoreDump (as built by oracle/build_ref.sh) aborts at exit now and then -- on its own stores too,
    after printing -- so a failed run is repeated."""
    for _ in range(8):
        r = subprocess.run([os.path.join(REF, "ovStoreDump"), "-S", seq, "-O", store] + what, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        if r.returncode == 0:
            return r.stdout.decode()
    raise RuntimeError("ovStoreDump kept failing")


def _reference_store(tmp, seq, ovb, erate=None):
    cfg = os.path.join(tmp, "cfg")
    subprocess.check_call([os.path.join(REF, "ovStoreConfig"), "-S", seq, "-M", "1", "-create", cfg, ovb], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    st = os.path.join(tmp, "ref.ovlStore")
    subprocess.check_call([os.path.join(REF, "ovStoreBuild"), "-O", st, "-S", seq, "-C", cfg] + (["-e", str(erate)] if erate else []),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return st


WHAT = [["-overlaps", "-unaligned"], ["-counts"], ["-overlaps", "17-60", "-coords"]]


@needs_ref
@pytest.mark.parametrize("erate", [None, 0.03])
def test_written_store_reads_back_like_the_reference_store(tmp_path, erate):
    seq = os.path.join(gu.GOLDEN, "A.seqStore")
    ovb = os.path.join(gu.GOLDEN, "A_default.ovb")
    ref = _reference_store(str(tmp_path), seq, ovb, erate)
    n_reads = gu.load_cases()["stores"]["A"]["reads"]
    recs = io_.ingest(_load_ovb(ovb), io_.encode_evalue(erate if erate else 1.0), n_reads)
    flat = str(tmp_path / "sorted.bin")
    np.ascontiguousarray(recs).tofile(flat)
    ours = str(tmp_path / "ours.ovlStore")
    subprocess.check_call([os.path.join(OURS, "ovltool"), "write-store", flat, ours, str(n_reads)])
    assert open(os.path.join(ours, "info"), "rb").read() == open(os.path.join(ref, "info"), "rb").read()
    assert open(os.path.join(ours, "0001-001"), "rb").read() == open(os.path.join(ref, "0001-001"), "rb").read()
    for w in (WHAT if erate else WHAT[1:]):                 # the full text dump once: the reference tool takes its time
        a, b = _dump(ours, seq, w), _dump(ref, seq, w)
        assert a == b and len(a) > 100, w


@needs_ref
@pytest.mark.gpu
def test_store_build_executable_matches_reference_store(tmp_path):
    seq = os.path.join(gu.GOLDEN, "A.seqStore")
    ovb = os.path.join(gu.GOLDEN, "A_default.ovb")
    ref = _reference_store(str(tmp_path), seq, ovb, 0.03)
    ours = str(tmp_path / "ours.ovlStore")
    r = subprocess.run([os.path.join(OURS, "ovlStoreBuild"), "-O", ours, "-S", seq, "-e", "0.03", ovb], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert open(os.path.join(ours, "0001-001"), "rb").read() == open(os.path.join(ref, "0001-001"), "rb").read()
    for w in WHAT:
        assert _dump(ours, seq, w) == _dump(ref, seq, w), w
