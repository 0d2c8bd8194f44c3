"""CPU tests of the C++ host side: sqStore reader and ovb/oc writer against reference-made files."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "canu_b200", "bin", "ovltool")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def _tool():
    if not os.path.exists(TOOL):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "canu_b200", "csrc")])
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "canu_b200", "host")])
    return TOOL


def _fasta_seqs(text):
    return [ln for ln in text.split("\n") if ln and not ln.startswith(">")]


@pytest.mark.parametrize("store", ["A", "B", "C"])
@pytest.mark.parametrize("mode", ["decode", "packed"])
def test_sqstore_reader_matches_reference_dump(store, mode):
    """Our reader vs the reference's own sqStoreDumpFASTQ of the same store (2-bit, 3-bit with N,
    homopolymer-compressed default version)."""
    cmd = [_tool(), "dump-store", os.path.join(gu.GOLDEN, store + ".seqStore")] + (["--packed"] if mode == "packed" else [])
    got = _fasta_seqs(subprocess.check_output(cmd).decode())
    with gzip.open(os.path.join(gu.GOLDEN, store + ".dump.fasta.gz"), "rt") as f:
        want = _fasta_seqs(f.read())
    assert len(got) == len(want) == gu.load_cases()["stores"][store]["reads"]
    assert got == want


def test_ovb_reader_and_writer_roundtrip(tmp_path):
    """Decode the reference-written .ovb with our snappy reader, rewrite it with our writer: records
    identical, .oc byte-identical to the reference's, and (where the reference tools are present) the
    reference's own overlapConvert reads our file back to the golden text."""
    src = os.path.join(gu.GOLDEN, "A_default.ovb")
    n_reads = gu.load_cases()["stores"]["A"]["reads"]
    out = str(tmp_path / "rt.ovb")
    subprocess.check_call([_tool(), "rewrite-ovb", src, out, str(n_reads)])
    a = subprocess.check_output([_tool(), "dump-ovb", src]).decode().splitlines()
    b = subprocess.check_output([_tool(), "dump-ovb", out]).decode().splitlines()
    assert a == b and len(a) == gu.get_case("A_default")["overlaps"]
    recs = np.zeros(len(a), dtype=[("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
    for i, ln in enumerate(a):
        x = ln.split()
        recs[i] = (int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16))
    assert gu.format_records(recs) == gu.load_golden_lines("A_default")
    assert open(str(tmp_path / "rt.oc"), "rb").read() == open(os.path.join(gu.GOLDEN, "A_default.oc"), "rb").read()
    conv = os.path.join(REFBIN, "overlapConvert")
    if os.path.exists(conv):
        txt = subprocess.check_output([conv, "-S", os.path.join(gu.GOLDEN, "A.seqStore"), "-unaligned", out]).decode()
        assert sorted(ln for ln in txt.split("\n") if ln) == gu.load_golden_lines("A_default")


def test_drop_in_cli_errors_like_the_reference(tmp_path):
    exe = os.path.join(ROOT, "canu_b200", "bin", "overlapInCore")
    _tool()
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 1 and b"No kmer length supplied" in r.stderr and b"No output file name" in r.stderr
    r = subprocess.run([exe, "-k", "22", "-o", str(tmp_path / "x.ovb"), "a.seqStore", "b.seqStore"], capture_output=True)
    assert r.returncode == 1 and b"Unknown option 'b.seqStore'" in r.stderr
    r = subprocess.run([exe, "-k", "22", "-o", str(tmp_path / "x.ovb"), str(tmp_path / "nope.seqStore")], capture_output=True)
    assert r.returncode != 0


def test_pack_ovb_roundtrip(tmp_path):
    """ovltool pack-ovb: flat {u32 a, u32 b, u64 dat0, u64 dat1} records -> snappy .ovb + .oc, read back unchanged
    (the path tools/ingest_timing.py uses to hand the same records to the reference's ovFile reader)."""
    rng = np.random.default_rng(3)
    n, n_reads = 100_003, 5000                      # more than two 43,680-record blocks, not a multiple of the block size
    recs = np.zeros(n, dtype=[("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
    recs["a_iid"] = rng.integers(1, n_reads + 1, n); recs["b_iid"] = rng.integers(1, n_reads + 1, n)
    recs["w0"] = rng.integers(0, 1 << 62, n, dtype=np.uint64); recs["w1"] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    flat, ovb = str(tmp_path / "in.bin"), str(tmp_path / "out.ovb")
    recs.tofile(flat)
    subprocess.check_call([_tool(), "pack-ovb", flat, ovb, str(n_reads)])
    lines = subprocess.check_output([_tool(), "dump-ovb", ovb]).decode().splitlines()
    assert len(lines) == n
    back = np.array([(int(x[0]), int(x[1]), int(x[2], 16), int(x[3], 16)) for x in (ln.split() for ln in lines)], dtype=recs.dtype)
    assert np.array_equal(back, recs)
    oc = gu.read_oc(str(tmp_path / "out.oc"))
    want = np.bincount(recs["a_iid"], minlength=n_reads + 1) + np.bincount(recs["b_iid"], minlength=n_reads + 1)
    assert oc[0] == n and np.array_equal(np.asarray(oc[1])[: n_reads + 1], want[: n_reads + 1])


@pytest.mark.parametrize("store", ["A", "B", "C"])
@pytest.mark.parametrize("threads", [2, 5])
def test_prefetcher_with_several_packers_hands_out_the_plan_in_order(store, threads):
    """The executable's prefetcher (canu_b200/host/prefetch.h) packs the read ranges of a worker's plan ahead of the GPU, with
    three packer threads on HiFi-like jobs.  `ovltool prefetch-check` walks a plan of hash blocks and uneven ref pieces
    (one empty) with one packer and with several, recycling the buffers as the worker does: same batches, same order."""
    r = subprocess.run([_tool(), "prefetch-check", os.path.join(gu.GOLDEN, store + ".seqStore"), "9", str(threads)], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode().startswith("prefetch-check ok: 12 items, %d threads" % threads)

