"""Seeded synthetic read sets for parity fixtures and for bench.py.

Mirrors what SURVEY.md 8(d) prescribes for inputs: a uniform-random genome
(optionally with a planted multi-copy repeat), reads sampled from both strands
with a length distribution, and a per-base error injector with
sub:ins:del = 4:3:3 scaled to a per-read error rate.  (The reference's own
`seqrequester simulate` makes error-free reads only and `mutate` is buggy,
SURVEY.md 7.12, so the injector is ours.)

Everything is numpy + a seed; the same call makes the same reads on any box.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp(seq: np.ndarray) -> np.ndarray:
    """Reverse complement of an ASCII uint8 array."""
    return _COMP[seq[::-1]]


def make_genome(size: int, seed: int, repeat_len: int = 0, repeat_copies: int = 0) -> np.ndarray:
    """Uniform ACGT genome; optionally plant `repeat_copies` copies of one random
    `repeat_len` segment at evenly spaced positions (exercises skip k-mers)."""
    rng = np.random.default_rng(seed)
    g = _ACGT[rng.integers(0, 4, size=size)]
    if repeat_len > 0 and repeat_copies > 1:
        unit = _ACGT[rng.integers(0, 4, size=repeat_len)]
        step = size // repeat_copies
        for c in range(repeat_copies):
            p = c * step + step // 3
            if p + repeat_len <= size:
                g[p:p + repeat_len] = unit
    return g


def inject_errors(read: np.ndarray, rate: float, rng: np.random.Generator) -> np.ndarray:
    """Per-base errors at total rate `rate`, split sub:ins:del = 4:3:3."""
    if rate <= 0.0:
        return read
    n = read.size
    u = rng.random(n)
    sub = u < rate * 0.4
    ins = (u >= rate * 0.4) & (u < rate * 0.7)
    dele = (u >= rate * 0.7) & (u < rate)
    out = read.copy()
    # substitutions: shift to a different base
    k = int(sub.sum())
    if k:
        idx = np.searchsorted(_ACGT, out[sub])
        out[sub] = _ACGT[(idx + rng.integers(1, 4, size=k)) % 4]
    keep = ~dele
    # insertions: a random base after the position
    reps = np.ones(n, dtype=np.int64)
    reps[ins] = 2
    reps[~keep] = 0
    res = np.repeat(out, reps)
    # positions of the inserted copies (second of each pair)
    ends = np.cumsum(reps)
    ipos = ends[ins & keep] - 1
    if ipos.size:
        res[ipos] = _ACGT[rng.integers(0, 4, size=ipos.size)]
    return res


def simulate_reads(genome: np.ndarray, coverage: float, len_lo: int, len_hi: int,
                   err_rate: float, seed: int, n_frac: float = 0.0,
                   n_reads_with_n: int = 0, lognormal: tuple | None = None) -> list[np.ndarray]:
    """Sample reads until `coverage` x genome bases are drawn.

    Lengths are uniform in [len_lo, len_hi] or, with lognormal=(mean, sigma) of the
    underlying normal, log-normal clipped to that range.  Half the reads are
    reverse-complemented.  `n_reads_with_n` reads get a few 'N's (at rate n_frac)."""
    rng = np.random.default_rng(seed)
    G = genome.size
    target = int(coverage * G)
    reads = []
    total = 0
    while total < target:
        if lognormal is not None:
            L = int(np.clip(rng.lognormal(lognormal[0], lognormal[1]), len_lo, len_hi))
        else:
            L = int(rng.integers(len_lo, len_hi + 1))
        L = min(L, G)
        p = int(rng.integers(0, G - L + 1))
        r = genome[p:p + L]
        if rng.random() < 0.5:
            r = revcomp(r)
        r = inject_errors(r, err_rate, rng)
        reads.append(np.ascontiguousarray(r))
        total += L
    if n_reads_with_n > 0 and n_frac > 0:
        pick = rng.choice(len(reads), size=min(n_reads_with_n, len(reads)), replace=False)
        for i in pick:
            r = reads[i].copy()
            m = rng.random(r.size) < n_frac
            r[m] = ord("N")
            reads[i] = r
    return reads


def write_fasta(path: str, reads: list[np.ndarray]) -> None:
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">read%d\n" % (i + 1))
            f.write(r.tobytes())
            f.write(b"\n")


def read_fasta(path: str) -> list[np.ndarray]:
    """Minimal FASTA reader (one or more sequence lines per record) -> ASCII arrays."""
    reads, cur = [], []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    reads.append(np.frombuffer(b"".join(cur), dtype=np.uint8).copy())
                    cur = []
                elif reads or cur == []:
                    pass
            else:
                cur.append(line.strip())
    if cur:
        reads.append(np.frombuffer(b"".join(cur), dtype=np.uint8).copy())
    return reads
