"""TEST INFRASTRUCTURE -- ctypes wrapper around oracle/libovl_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package
(canu_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) into oracle/libovl_oracle.so."""
    so = os.path.join(_HERE, "libovl_oracle.so")
    src = os.path.join(_HERE, "ovl_oracle.c")
    hdr = os.path.join(_HERE, "ovl_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-ffp-contract=off", "-w", "-shared", "-fPIC",
                               src, "-o", so, "-lm"])
    return so


class Params(C.Structure):
    _fields_ = [("kmer_len", C.c_uint32), ("max_erate", C.c_double), ("align_noise", C.c_double),
                ("partial", C.c_int32), ("unique_per_pair", C.c_int32), ("min_olap_len", C.c_int32),
                ("no_hopeless", C.c_int32), ("min_kmers", C.c_int32), ("hash_bits", C.c_uint32),
                ("hash_load", C.c_double), ("hash_data_len", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "kmer_hits_without_olap", "kmer_hits_with_olap", "kmer_hits_skipped", "multi_overlap",
        "total_overlaps", "contained", "dovetail", "extend_calls", "dp_cells", "char_compares",
        "hash_inserts", "ref_lookups", "seed_hits")]


REC_DTYPE = np.dtype([("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
PAIR_DTYPE = np.dtype([("ref_id", "<u4"), ("hash_id", "<u4"), ("dir", "<i4"), ("consistent", "<i4"),
                       ("diag_ct", "<i4"), ("diag_bgn", "<i4"), ("diag_end", "<i4"), ("n_seeds", "<i4"),
                       ("seed_begin", "<i8")])
SEED_DTYPE = np.dtype([("start", "<i4"), ("offset", "<i4"), ("len", "<i4")])
EXT_DTYPE = np.dtype([("ref_id", "<u4"), ("hash_id", "<u4"), ("dir", "<i4"), ("seed_start", "<i4"),
                      ("seed_offset", "<i4"), ("seed_len", "<i4"), ("s_lo", "<i4"), ("s_hi", "<i4"),
                      ("t_lo", "<i4"), ("t_hi", "<i4"), ("errors", "<i4"), ("kind", "<i4"), ("delta_ct", "<i4")])


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ovo_create.restype = C.c_void_p
        L.ovo_create.argtypes = [C.POINTER(Params)]
        L.ovo_destroy.argtypes = [C.c_void_p]
        L.ovo_set_reads.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ovo_set_skip_kmers.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
        L.ovo_enable_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ovo_run.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        for f in ("ovo_num_records", "ovo_num_pair_traces", "ovo_num_ext_traces"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("ovo_records", "ovo_pair_traces", "ovo_seed_traces", "ovo_ext_traces", "ovo_edit_match_limit"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.ovo_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.ovo_max_errors.restype = C.c_uint32
        L.ovo_max_errors.argtypes = [C.c_void_p]
        L.ovo_error_bound.restype = C.c_int32
        L.ovo_error_bound.argtypes = [C.c_void_p, C.c_int32]
        L.ovo_branch_match_value.restype = C.c_double
        L.ovo_branch_match_value.argtypes = [C.c_void_p]
        L.ovo_extend_one.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB = L
    return _LIB


def f32(x: float) -> float:
    """--maxerate / --alignnoise go through strtof in the reference (overlapInCore.C:380,382)."""
    return float(np.float32(x))


def _copy(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Oracle:
    def __init__(self, kmer_len=22, max_erate=0.06, align_noise=1.0, partial=False, unique=True, min_olap_len=0,
                 no_hopeless=False, min_kmers=False, hash_bits=22, hash_load=0.6, hash_data_len=100000000):
        self.L = _lib()
        self.p = Params(kmer_len, f32(max_erate), f32(align_noise), int(partial), int(unique), int(min_olap_len),
                        int(no_hopeless), int(min_kmers), hash_bits, hash_load, hash_data_len)
        self.h = self.L.ovo_create(C.byref(self.p))
        self.n_reads = 0

    def close(self):
        if self.h:
            self.L.ovo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_reads(self, reads):
        """reads: list of ASCII uint8 arrays; read ID = index + 1."""
        lens = np.array([r.size for r in reads], dtype=np.uint32)
        offs = np.zeros(len(reads), dtype=np.uint64)
        if len(reads) > 1:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        buf = np.concatenate(reads) if reads else np.zeros(0, np.uint8)
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.L.ovo_set_reads(self.h, len(reads), buf.ctypes.data, offs.ctypes.data, lens.ctypes.data)
        self.n_reads = len(reads)

    def set_skip_kmers(self, kmers):
        s = "".join(kmers).encode()
        self.L.ovo_set_skip_kmers(self.h, len(kmers), s)

    def run(self, hb=1, he=None, rb=1, re=None, threads=None, trace_pairs=False, trace_exts=False):
        he = self.n_reads if he is None else he
        re = self.n_reads if re is None else re
        threads = threads or min(8, os.cpu_count() or 1)
        self.L.ovo_enable_trace(self.h, int(trace_pairs), int(trace_exts))
        self.L.ovo_run(self.h, hb, he, rb, re, threads)
        recs = _copy(self.L.ovo_records(self.h), self.L.ovo_num_records(self.h), REC_DTYPE)
        return recs

    def stats(self):
        s = Stats()
        self.L.ovo_get_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def pair_traces(self):
        n = self.L.ovo_num_pair_traces(self.h)
        pt = _copy(self.L.ovo_pair_traces(self.h), n, PAIR_DTYPE)
        ns = int(pt["n_seeds"].sum()) if n else 0
        sd = _copy(self.L.ovo_seed_traces(self.h), ns, SEED_DTYPE)
        return pt, sd

    def ext_traces(self):
        return _copy(self.L.ovo_ext_traces(self.h), self.L.ovo_num_ext_traces(self.h), EXT_DTYPE)

    def max_errors(self):
        return self.L.ovo_max_errors(self.h)

    def edit_match_limit(self):
        n = self.max_errors()
        return _copy(self.L.ovo_edit_match_limit(self.h), n, np.dtype("<i4"))

    def error_bound(self, n):
        return self.L.ovo_error_bound(self.h, n)

    def branch_match_value(self):
        return self.L.ovo_branch_match_value(self.h)

    def extend_one(self, S, T, seed_start, seed_offset, seed_len):
        out = np.zeros(1, dtype=EXT_DTYPE)
        delta = np.zeros(max(len(S), len(T)) + 8, dtype=np.int32)
        self.L.ovo_extend_one(self.h, bytes(S), len(S), bytes(T), len(T), seed_start, seed_offset, seed_len,
                              out.ctypes.data, delta.ctypes.data, delta.size)
        return out[0], delta[: int(out[0]["delta_ct"])].copy()


def sort_records(recs: np.ndarray) -> np.ndarray:
    """Canonical order: (a_iid, b_iid, w0, w1) -- ovOverlap::operator< (stores/ovOverlap.H:263-276)."""
    return np.sort(recs, order=["a_iid", "b_iid", "w0", "w1"])


def decode_record_fields(recs: np.ndarray) -> dict:
    """Split the two 64-bit words into named fields (stores/ovOverlap.H:49-67)."""
    M = (1 << 21) - 1
    w0, w1 = recs["w0"], recs["w1"]
    return dict(a_iid=recs["a_iid"], b_iid=recs["b_iid"], ahg5=w0 & M, ahg3=(w0 >> 21) & M,
                evalue=(w0 >> 42) & 0xFFFF, flipped=(w0 >> 58) & 1, forOBT=(w0 >> 59) & 1,
                forDUP=(w0 >> 60) & 1, forUTG=(w0 >> 61) & 1, bhg5=w1 & M, bhg3=(w1 >> 21) & M,
                span=(w1 >> 42) & M)
