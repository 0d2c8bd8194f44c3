"""Python (ctypes) binding of the C ABI in include/ovlb200.h.

Mirrors the reference's operator surface for the ovl hot path:

    OverlapParams        <->  oicParameters G            (overlapInCore.H:367-473)
    Overlapper           <->  OverlapDriver()            (overlapInCore.C:162-277):
        load_hash_reads / mark_skip_kmers / build_index  = Build_Hash_Index
        overlap_ref_batch                                = Process_Overlaps over a ref range
    overlap_in_core()    <->  one `overlapInCore -h a-b -r c-d` tile, records returned in memory

There is no CPU fallback: if canu_b200/libovlb200.so is missing, or no CUDA device is
present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libovlb200.so")
_LIB = None

MAX_READLEN = (1 << 21) - 1


class OvlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ovlb200 error %d: %s" % (code, msg))
        self.code = code


class _Params(C.Structure):
    _fields_ = [("kmer_len", C.c_uint32), ("partial", C.c_int32), ("unique_per_pair", C.c_int32),
                ("min_olap_len", C.c_int32), ("use_hopeless_check", C.c_int32),
                ("filter_by_kmer_count", C.c_uint64), ("max_erate", C.c_double),
                ("branch_match_value", C.c_double), ("min_branch_tail_slope", C.c_double),
                ("minkmers_exp_factor", C.c_double), ("edit_match_limit", C.POINTER(C.c_int32)),
                ("n_edit_match_limit", C.c_uint32), ("max_read_len", C.c_uint32),
                ("device_mem_budget", C.c_uint64)]


class _Reads(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("packed_bytes", C.c_uint64), ("byte_offset", C.c_void_p),
                ("len", C.c_void_p), ("n_reads", C.c_uint32), ("first_read_id", C.c_uint32),
                ("n_read", C.c_void_p), ("n_pos", C.c_void_p), ("n_n", C.c_uint64),
                ("src_len", C.c_void_p), ("clear_bgn", C.c_void_p), ("homopoly_compress", C.c_uint32)]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "kmer_hits_without_olap", "kmer_hits_with_olap", "kmer_hits_skipped", "multi_overlap",
        "total_overlaps", "contained", "dovetail", "extend_calls", "dp_cells", "hash_kmers",
        "ref_kmers", "seed_hits", "seed_runs", "pairs", "ext_busy_ns", "ext_capacity_ns")]


class _Tile(C.Structure):
    _fields_ = [("hash_bgn", C.c_uint32), ("hash_end", C.c_uint32), ("ref_bgn", C.c_uint32), ("ref_end", C.c_uint32),
                ("hash_bases", C.c_uint64), ("ref_bases", C.c_uint64), ("cost", C.c_double), ("has_hash_reads", C.c_int32)]


class _Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "upload_ms", "encode_ms", "index_tuples_ms", "index_sort_ms", "index_table_ms", "index_skip_ms",
        "probe_ms", "expand_ms", "sort_ms", "chain_ms", "extend_ms", "download_ms", "total_ms")]


RECORD_DTYPE = np.dtype([("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
PAIR_DTYPE = np.dtype([("ref_id", "<u4"), ("hash_id", "<u4"), ("dir", "<i4"), ("consistent", "<i4"),
                       ("diag_ct", "<i4"), ("diag_bgn", "<i4"), ("diag_end", "<i4"), ("n_seeds", "<i4"),
                       ("seed_begin", "<i8")])
SEED_DTYPE = np.dtype([("start", "<i4"), ("offset", "<i4"), ("len", "<i4")])

EXPORTS = [
    "ovlb_last_error", "ovlb_device_count", "ovlb_device_memory", "ovlb_device_total_memory", "ovlb_create", "ovlb_destroy", "ovlb_load_hash_reads",
    "ovlb_mark_skip_kmers", "ovlb_build_index", "ovlb_overlap_ref_batch", "ovlb_stage_ref_batch",
    "ovlb_run_staged", "ovlb_stage_next_ref_batch", "ovlb_advance_staged", "ovlb_fetch_records", "ovlb_get_counters", "ovlb_reset_counters",
    "ovlb_get_timings", "ovlb_kernel_launches", "ovlb_timer_start", "ovlb_timer_stop", "ovlb_host_register", "ovlb_host_unregister", "ovlb_debug_pairs", "ovlb_debug_extend", "ovlb_debug_index_info", "ovlb_ingest_records",
    "ovlb_params_init", "ovlb_params_free", "ovlb_parse_erate", "ovlb_pack_reads", "ovlb_reads_view",
    "ovlb_reads_free", "ovlb_kmer_keys", "ovlb_kmer_census", "ovlb_plan_tiles", "ovlb_plan_balanced", "ovlb_hash_block_bases", "ovlb_assign_tiles",
]


def load_library():
    """Load the CUDA extension; fail loudly if it was not built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("canu_b200: %s is missing -- build it with `make -C canu_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ovlb_last_error.restype = C.c_char_p
    L.ovlb_device_count.restype = C.c_int
    L.ovlb_create.argtypes = [C.c_int, C.POINTER(_Params), C.POINTER(C.c_void_p)]
    L.ovlb_destroy.argtypes = [C.c_void_p]
    L.ovlb_load_hash_reads.argtypes = [C.c_void_p, C.c_void_p]
    L.ovlb_mark_skip_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.ovlb_build_index.argtypes = [C.c_void_p]
    L.ovlb_overlap_ref_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ovlb_stage_ref_batch.argtypes = [C.c_void_p, C.c_void_p]
    L.ovlb_run_staged.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ovlb_fetch_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ovlb_get_counters.argtypes = [C.c_void_p, C.POINTER(_Counters)]
    L.ovlb_reset_counters.argtypes = [C.c_void_p]
    L.ovlb_get_timings.argtypes = [C.c_void_p, C.POINTER(_Timings)]
    L.ovlb_kernel_launches.argtypes = [C.c_void_p]
    L.ovlb_kernel_launches.restype = C.c_uint64
    L.ovlb_host_register.argtypes = [C.c_void_p, C.c_uint64]
    L.ovlb_host_unregister.argtypes = [C.c_void_p]
    L.ovlb_timer_start.argtypes = [C.c_void_p]
    L.ovlb_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.ovlb_debug_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                   C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ovlb_debug_extend.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 8 + [C.c_uint32]
    L.ovlb_debug_index_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ovlb_ingest_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ovlb_kmer_census.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.POINTER(C.c_uint64), C.c_void_p]
    L.ovlb_params_init.argtypes = [C.POINTER(_Params), C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_uint32]
    L.ovlb_params_free.argtypes = [C.POINTER(_Params)]
    L.ovlb_parse_erate.argtypes = [C.c_char_p]
    L.ovlb_parse_erate.restype = C.c_double
    L.ovlb_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(C.c_void_p)]
    L.ovlb_reads_view.argtypes = [C.c_void_p]
    L.ovlb_reads_view.restype = C.POINTER(_Reads)
    L.ovlb_reads_free.argtypes = [C.c_void_p]
    L.ovlb_kmer_keys.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ovlb_plan_tiles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ovlb_assign_tiles.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    L.ovlb_hash_block_bases.restype = C.c_uint64
    L.ovlb_hash_block_bases.argtypes = [C.c_uint64, C.c_uint32, C.c_double, C.c_uint64]
    L.ovlb_plan_balanced.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_double, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise OvlError(rc, load_library().ovlb_last_error().decode(errors="replace"))


@dataclass
class OverlapParams:
    """Command-line level parameters, named after the reference's flags."""
    kmer_len: int = 22            # -k
    max_erate: float = 0.06       # --maxerate (rounded through float like strtof)
    align_noise: float = 1.0      # --alignnoise
    partial: bool = False         # -partial
    unique: bool = True           # -u / -m
    min_olap_len: int = 0         # --minlength
    no_hopeless: bool = False     # -z
    min_kmers: bool = False       # --minkmers
    max_read_len: int = 0         # 0 = AS_MAX_READLEN

    def erate_as_parsed(self) -> float:
        return load_library().ovlb_parse_erate(repr(float(self.max_erate)).encode())


class PackedReads:
    """Reads in the wire format of ovlb_reads (2-bit packed + N list), owned by the C side."""

    def __init__(self, reads, first_read_id=1, min_len=0):
        L = load_library()
        lens = np.array([len(r) for r in reads], dtype=np.uint32)
        offs = np.zeros(len(reads), dtype=np.uint64)
        if len(reads) > 1:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        if len(reads):
            buf = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint8) for r in reads]))
        else:
            buf = np.zeros(1, np.uint8)
        self._h = C.c_void_p()
        _check(L.ovlb_pack_reads(buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(reads),
                                 first_read_id, min_len, C.byref(self._h)))
        self.view = L.ovlb_reads_view(self._h)
        self.n_reads = len(reads)
        self.first_read_id = first_read_id
        self.total_bases = int(np.where(lens >= min_len, lens, 0).sum())
        self.packed_bytes = int(self.view.contents.packed_bytes)

    def pin(self):
        """Page-lock the packed bases and the per-read arrays (faster, asynchronous uploads)."""
        L = load_library()
        v = self.view.contents
        self._pinned = [p for p, n in ((v.packed, v.packed_bytes), (v.byte_offset, 8 * v.n_reads), (v.len, 4 * v.n_reads)) if p and n]
        for p, n in ((v.packed, v.packed_bytes), (v.byte_offset, 8 * v.n_reads), (v.len, 4 * v.n_reads)):
            if p and n:
                _check(L.ovlb_host_register(p, n))

    def close(self):
        if self._h:
            for p in getattr(self, "_pinned", []):
                load_library().ovlb_host_unregister(p)
            self._pinned = []
            load_library().ovlb_reads_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def homopoly_compress(read: np.ndarray) -> np.ndarray:
    """Every run of equal bases collapsed to one base (utility/src/sequence/sequence-v1.C:203-261); ASCII uint8 in/out."""
    r = np.asarray(read, dtype=np.uint8)
    if r.size == 0:
        return r
    keep = np.ones(r.size, dtype=bool)
    keep[1:] = (r[1:] | 0x20) != (r[:-1] | 0x20)
    return r[keep]


class RawReads(PackedReads):
    """Reads handed over AS STORED (ovlb_reads.src_len / clear_bgn / homopoly_compress): the device applies homopolymer
    compression and the clear range.  raw_reads: ASCII arrays (ACGT only); clear: list of (bgn, end) in the coordinates of
    the (compressed) read, or None for the whole read."""

    def __init__(self, raw_reads, clear=None, homopoly=False, first_read_id=1, min_len=0):
        super().__init__(raw_reads, first_read_id=first_read_id, min_len=0)
        n = len(raw_reads)
        src = np.array([len(r) for r in raw_reads], dtype=np.uint32)
        cb = np.zeros(n, dtype=np.uint32)
        out = np.zeros(n, dtype=np.uint32)
        for i, r in enumerate(raw_reads):
            full = homopoly_compress(r).size if homopoly else len(r)
            b, e = clear[i] if clear is not None else (0, full)
            assert 0 <= b <= e <= full
            cb[i] = b
            out[i] = e - b if e - b >= min_len else 0
        src[out == 0] = 0
        self._src, self._cb, self._out = src, cb, out
        v = self.view.contents
        self._raw_view = _Reads(v.packed, v.packed_bytes, v.byte_offset, out.ctypes.data, v.n_reads, v.first_read_id,
                                None, None, 0, src.ctypes.data, cb.ctypes.data, 1 if homopoly else 0)
        self.view = C.pointer(self._raw_view)
        self.total_bases = int(out.sum())


class Overlapper:
    """One ovl context on one GPU."""

    def __init__(self, params: OverlapParams, device: int = 0, device_mem_budget: int = 0):
        L = load_library()
        self.L = L
        self.params = params
        self._p = _Params()
        _check(L.ovlb_params_init(C.byref(self._p), params.kmer_len, params.erate_as_parsed(),
                                  float(np.float32(params.align_noise)), int(params.partial), int(params.unique),
                                  int(params.min_olap_len), int(params.no_hopeless), int(params.min_kmers),
                                  int(params.max_read_len)))
        self._p.device_mem_budget = device_mem_budget
        self._h = C.c_void_p()
        try:
            _check(L.ovlb_create(device, C.byref(self._p), C.byref(self._h)))
        except Exception:
            L.ovlb_params_free(C.byref(self._p))
            raise

    # --- host tables, for tests ---
    def edit_match_limit(self) -> np.ndarray:
        n = self._p.n_edit_match_limit
        return np.ctypeslib.as_array(self._p.edit_match_limit, shape=(n,)).copy()

    def close(self):
        if getattr(self, "_h", None):
            self.L.ovlb_destroy(self._h)
            self._h = C.c_void_p()
            self.L.ovlb_params_free(C.byref(self._p))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- hash side ---
    def load_hash_reads(self, packed: PackedReads):
        _check(self.L.ovlb_load_hash_reads(self._h, C.cast(packed.view, C.c_void_p)))

    def mark_skip_kmers(self, kmers):
        keys = np.zeros(2 * len(kmers), dtype=np.uint64)
        f, r = C.c_uint64(), C.c_uint64()
        K = self.params.kmer_len
        for i, k in enumerate(kmers):
            kb = k.encode() if isinstance(k, str) else bytes(k)
            if len(kb) != K:
                raise ValueError("skip k-mer of length %d, expected %d" % (len(kb), K))
            _check(self.L.ovlb_kmer_keys(kb, K, C.byref(f), C.byref(r)))
            keys[2 * i], keys[2 * i + 1] = f.value, r.value
        _check(self.L.ovlb_mark_skip_kmers(self._h, keys.ctypes.data, keys.size))

    def build_index(self):
        _check(self.L.ovlb_build_index(self._h))

    # --- ref side ---
    def overlap_ref_batch_into(self, packed: PackedReads, out: np.ndarray) -> int:
        """Same as overlap_ref_batch but into a caller-owned (e.g. pinned) record array; returns the count."""
        n = C.c_uint64()
        _check(self.L.ovlb_overlap_ref_batch(self._h, C.cast(packed.view, C.c_void_p), out.ctypes.data, out.size, C.byref(n)))
        return n.value

    def overlap_ref_batch(self, packed: PackedReads, cap: int = 1 << 20) -> np.ndarray:
        out = np.zeros(cap, dtype=RECORD_DTYPE)
        n = C.c_uint64()
        rc = self.L.ovlb_overlap_ref_batch(self._h, C.cast(packed.view, C.c_void_p), out.ctypes.data, cap, C.byref(n))
        if rc == -3 and n.value > cap:          # record buffer too small: fetch again with the right size
            out = np.zeros(n.value, dtype=RECORD_DTYPE)
            _check(self.L.ovlb_fetch_records(self._h, out.ctypes.data, out.size, C.byref(n)))
        else:
            _check(rc)
        return out[: n.value].copy()

    def stage_ref_batch(self, packed: PackedReads):
        """Upload + encode a ref batch on the copy stream; may precede build_index() (the upload then overlaps it)."""
        self._staged_src = packed                      # the buffers must outlive the asynchronous upload
        _check(self.L.ovlb_stage_ref_batch(self._h, C.cast(packed.view, C.c_void_p)))

    def fetch_records_into(self, out: np.ndarray) -> int:
        n = C.c_uint64()
        _check(self.L.ovlb_fetch_records(self._h, out.ctypes.data, out.size, C.byref(n)))
        return n.value

    def run_staged(self) -> int:
        n = C.c_uint64()
        _check(self.L.ovlb_run_staged(self._h, C.byref(n)))
        return n.value

    def fetch_records(self, n_hint: int = 0) -> np.ndarray:
        out = np.zeros(max(n_hint, 1), dtype=RECORD_DTYPE)
        n = C.c_uint64()
        rc = self.L.ovlb_fetch_records(self._h, out.ctypes.data, out.size, C.byref(n))
        if rc == -3:
            out = np.zeros(n.value, dtype=RECORD_DTYPE)
            _check(self.L.ovlb_fetch_records(self._h, out.ctypes.data, out.size, C.byref(n)))
        else:
            _check(rc)
        return out[: n.value].copy()

    # --- stats ---
    def counters(self) -> dict:
        c = _Counters()
        _check(self.L.ovlb_get_counters(self._h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in _Counters._fields_}

    def reset_counters(self):
        _check(self.L.ovlb_reset_counters(self._h))

    def timings(self) -> dict:
        t = _Timings()
        _check(self.L.ovlb_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in _Timings._fields_}

    def kernel_launches(self) -> int:
        return self.L.ovlb_kernel_launches(self._h)

    def timer_start(self):
        _check(self.L.ovlb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self.L.ovlb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def ingest_records(self, recs: np.ndarray, max_evalue: int, max_id: int) -> np.ndarray:
        """Mirror + filter + sort overlap records the way the overlap-store build does (ovlb_ingest_records)."""
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        out = np.zeros(max(2 * recs.size, 1), dtype=RECORD_DTYPE)
        n = C.c_uint64()
        _check(self.L.ovlb_ingest_records(self._h, recs.ctypes.data, recs.size, max_evalue, max_id, out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value]

    def kmer_census(self, distinct_fraction=-1.0, min_count=0, slice_bits=0, cap=1 << 20):
        """Frequent canonical k-mers of the loaded hash reads (the `-k` skip list): (keys, counts, stats)."""
        while True:
            keys = np.zeros(cap, dtype=np.uint64); cnts = np.zeros(cap, dtype=np.uint32)
            n = C.c_uint64(); st = (C.c_uint64 * 4)()
            rc = self.L.ovlb_kmer_census(self._h, slice_bits, float(distinct_fraction), int(min_count), keys.ctypes.data, cnts.ctypes.data,
                                         cap, C.byref(n), st)
            if rc == -3 and n.value >= cap:
                cap *= 4
                continue
            _check(rc)
            return keys[: n.value], cnts[: n.value], {"distinct": st[0], "present": st[1], "unique": st[2], "threshold": st[3]}

    # --- debug taps (tests) ---
    def debug_index_info(self) -> dict:
        out = (C.c_uint64 * 4)()
        _check(self.L.ovlb_debug_index_info(self._h, out))
        return {"distinct": out[0], "occurrences": out[1], "slots": out[2], "bucketed": bool(out[3])}

    def debug_pairs(self):
        np_, ns_ = C.c_uint64(), C.c_uint64()
        self.L.ovlb_debug_pairs(self._h, None, 0, C.byref(np_), None, 0, C.byref(ns_))
        pairs = np.zeros(max(np_.value, 1), dtype=PAIR_DTYPE)
        seeds = np.zeros(max(ns_.value, 1), dtype=SEED_DTYPE)
        _check(self.L.ovlb_debug_pairs(self._h, pairs.ctypes.data, pairs.size, C.byref(np_),
                                       seeds.ctypes.data, seeds.size, C.byref(ns_)))
        return pairs[: np_.value], seeds[: ns_.value]

    def debug_extend(self, ref_index, direction, hash_index, seed_start, seed_offset, seed_len, delta_stride=0):
        n = len(ref_index)
        a = [np.ascontiguousarray(x, dtype=t) for x, t in (
            (ref_index, np.uint32), (direction, np.int32), (hash_index, np.uint32),
            (seed_start, np.int32), (seed_offset, np.int32), (seed_len, np.int32))]
        out = np.zeros((n, 7), dtype=np.int32)
        deltas = np.zeros((n, delta_stride), dtype=np.int32) if delta_stride else None
        _check(self.L.ovlb_debug_extend(self._h, n, *[x.ctypes.data for x in a], out.ctypes.data,
                                        deltas.ctypes.data if delta_stride else None, delta_stride))
        return out, deltas


def plan_tiles(read_lens, min_olap_len, hash_block_len, ref_block_len, hash_range=None, ref_range=None,
               strict_reference=False):
    """Cut a read set into hash-block x ref-block tiles like overlapInCorePartition
    (overlapInCorePartition.C:127-257).  read_lens[i] is the length of read ID i+1.
    Returns a list of dicts (hash_bgn, hash_end, ref_bgn, ref_end, hash_bases, ref_bases, cost)."""
    L = load_library()
    n = len(read_lens)
    rl = np.zeros(n + 2, dtype=np.uint32)
    rl[1:n + 1] = np.asarray(read_lens, dtype=np.uint32)
    hb, he = hash_range if hash_range else (1, n)
    rb, re_ = ref_range if ref_range else (1, n)
    cnt = C.c_uint64()
    _check(L.ovlb_plan_tiles(rl.ctypes.data, n, min_olap_len, hash_block_len, ref_block_len, hb, he, rb, re_,
                             int(strict_reference), None, 0, C.byref(cnt)))
    arr = (_Tile * max(cnt.value, 1))()
    _check(L.ovlb_plan_tiles(rl.ctypes.data, n, min_olap_len, hash_block_len, ref_block_len, hb, he, rb, re_,
                             int(strict_reference), C.cast(arr, C.c_void_p), cnt.value, C.byref(cnt)))
    return [{f: getattr(arr[i], f) for f, _ in _Tile._fields_} for i in range(cnt.value)]


def hash_block_bases(budget_bytes, max_read_len, max_erate, ref_batch_bases=256_000_000):
    """Bases of hash reads one context with `budget_bytes` of device memory can index (ovlb_hash_block_bases)."""
    return int(load_library().ovlb_hash_block_bases(int(budget_bytes), int(max_read_len), float(max_erate), int(ref_batch_bases)))


def plan_balanced(read_lens, min_olap_len, n_parts, hash_range=None, ref_range=None, lookup_weight=0.002):
    """Cut one hash block's ref range into n_parts contiguous tiles of equal estimated work (ovlb_plan_balanced)."""
    L = load_library()
    n = len(read_lens)
    rl = np.zeros(n + 2, dtype=np.uint32)
    rl[1:n + 1] = np.asarray(read_lens, dtype=np.uint32)
    hb, he = hash_range if hash_range else (1, n)
    rb, re_ = ref_range if ref_range else (1, n)
    cnt = C.c_uint64()
    arr = (_Tile * max(n_parts, 1))()
    _check(L.ovlb_plan_balanced(rl.ctypes.data, n, min_olap_len, hb, he, rb, re_, n_parts, float(lookup_weight), C.cast(arr, C.c_void_p), n_parts, C.byref(cnt)))
    return [{f: getattr(arr[i], f) for f, _ in _Tile._fields_} for i in range(cnt.value)]


def assign_tiles(tiles, n_workers):
    """Longest-processing-time-first owner of every tile (deterministic); returns a list of worker indices."""
    L = load_library()
    arr = (_Tile * max(len(tiles), 1))()
    for i, t in enumerate(tiles):
        for f, _ in _Tile._fields_:
            setattr(arr[i], f, t[f])
    owner = np.zeros(max(len(tiles), 1), dtype=np.uint32)
    _check(L.ovlb_assign_tiles(C.cast(arr, C.c_void_p), len(tiles), n_workers, owner.ctypes.data))
    return owner[:len(tiles)].tolist()


def stats_lines(c: dict) -> str:
    """The -s statistics file, verbatim format of overlapInCore.C:550-558."""
    return (" Kmer hits without olaps = %d\n    Kmer hits with olaps = %d\n  Multiple overlaps/pair = %d\n"
            " Total overlaps produced = %d\n      Contained overlaps = %d\n       Dovetail overlaps = %d\n"
            "Rejected by short window = 0\n Rejected by long window = 0\n" % (
                c["kmer_hits_without_olap"], c["kmer_hits_with_olap"], c["multi_overlap"],
                c["total_overlaps"], c["contained"], c["dovetail"]))


def overlap_in_core(reads, params: OverlapParams, hash_range=None, ref_range=None, skip_kmers=None,
                    device: int = 0, ref_batch_bases: int = 1 << 28, return_overlapper: bool = False):
    """One `overlapInCore -h hb-he -r rb-re` tile over in-memory reads (ASCII arrays; ID = index+1).

    Returns (records, counters).  Reads shorter than --minlength are neither hashed nor searched
    (Build_Hash_Index.C:525-526, Process_Overlaps.C:59-60)."""
    n = len(reads)
    hb, he = hash_range if hash_range else (1, n)
    rb, re_ = ref_range if ref_range else (1, n)
    hb, rb = max(hb, 1), max(rb, 1)
    he, re_ = min(he, n), min(re_, n)
    mx = max((len(r) for r in reads), default=1)
    p = OverlapParams(**{**params.__dict__, "max_read_len": max(mx, 64)})
    ov = Overlapper(p, device=device)
    try:
        minlen = max(params.min_olap_len, params.kmer_len)
        hp = PackedReads(reads[hb - 1:he], first_read_id=hb, min_len=minlen)
        ov.load_hash_reads(hp)
        if skip_kmers:
            ov.mark_skip_kmers(skip_kmers)
        ov.build_index()
        hp.close()
        out = []
        i = rb
        while i <= re_:
            j, tot = i, 0
            while j <= re_ and (tot == 0 or tot + len(reads[j - 1]) <= ref_batch_bases) and j - i < 200000:
                tot += len(reads[j - 1]); j += 1
            rp = PackedReads(reads[i - 1:j - 1], first_read_id=i, min_len=minlen)
            out.append(ov.overlap_ref_batch(rp))
            rp.close()
            i = j
        recs = np.concatenate(out) if out else np.zeros(0, dtype=RECORD_DTYPE)
        ctr = ov.counters()
        if return_overlapper:
            return recs, ctr, ov
        return recs, ctr
    finally:
        if not return_overlapper:
            ov.close()
