"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the k-mer census behind Canu's `-k` skip list.

What the pipeline runs (src/pipelines/canu/Meryl.pm:529-533 count, :603-607 `greater-than 1 ... union-sum`,
:663-671 `print at-least distinct=D at-least threshold=T`) on the reads of a sqStore:

  * `meryl count` counts CANONICAL k-mers: a k-mer and its reverse complement are one k-mer; windows that hold a base
    other than ACGT are skipped (src/meryl/src/meryl/merylOp-count.C, kmerIterator);
  * `greater-than 1` keeps the k-mers of count >= 2; the statistics of THAT database feed the next step;
  * `at-least distinct=D`: threshold = the smallest count v such that the number of distinct k-mers of count <= v is
    >= D x (number of distinct k-mers), walking the histogram upwards (src/meryl/src/meryl/merylOp-nextMer.C:103-115);
    `at-least threshold=T`: count >= T.  Both given: both must hold (Meryl.pm:639 "Kmer must meet at least BOTH").

Pinned against the reference `meryl` binary itself (oracle/_ref/bin/meryl, built by oracle/build_ref.sh):
tests/golden/make_census_golden.py -> tests/golden/census_*.json; tests/test_census.py checks this file against them.
Only tests/ import this module; the product path is canu_b200/csrc/ovl_index.cu (k_part1<true>, k_bucket_census).
"""
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


def canonical_counts(reads, K):
    """(sorted unique canonical keys, counts); key = sum(code_j << 2j) with A0 C1 G2 T3, canonical = min(fwd, rc)."""
    allk = []
    sh = np.arange(K, dtype=np.uint64) * np.uint64(2)
    for r in reads:
        c = _CODE[np.asarray(r, dtype=np.uint8)]
        n = c.size - K + 1
        if n <= 0:
            continue
        bad = (c == 255).astype(np.int32)
        cs = np.concatenate([[0], np.cumsum(bad)])
        ok = (cs[K:] - cs[:-K]) == 0
        c64 = np.where(c == 255, 0, c).astype(np.uint64)
        win = np.lib.stride_tricks.sliding_window_view(c64, K)
        fwd = (win << sh).sum(axis=1, dtype=np.uint64)
        rc = ((np.uint64(3) - win[:, ::-1]) << sh).sum(axis=1, dtype=np.uint64)
        allk.append(np.minimum(fwd, rc)[ok])
    if not allk:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64)
    return np.unique(np.concatenate(allk), return_counts=True)


def threshold_for(counts, distinct_fraction, min_count):
    """Threshold meryl derives from the histogram of the count >= 2 database, combined with an absolute one."""
    c2 = counts[counts >= 2]
    t_d = 0
    if distinct_fraction is not None and distinct_fraction >= 0 and c2.size:
        vals, occ = np.unique(c2, return_counts=True)
        target = int(distinct_fraction * c2.size)          # uint64 nKmersTarget = _fracDist * numDistinct
        cum = np.cumsum(occ)
        i = int(np.searchsorted(cum, target, side="left"))
        t_d = int(vals[min(i, vals.size - 1)])
    return max(t_d, int(min_count or 0), 2)


def key_to_text(key, K):
    return "".join("ACGT"[(int(key) >> (2 * j)) & 3] for j in range(K))


def text_canonical(s):
    """Orientation-free name of a k-mer text: the smaller of the text and its reverse complement."""
    rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    return min(s, rc)


def frequent_kmers(reads, K, distinct_fraction=None, min_count=0):
    """{orientation-free k-mer text: count} of the k-mers the pipeline would pass to overlapInCore -k, and the threshold."""
    keys, counts = canonical_counts(reads, K)
    thr = threshold_for(counts, distinct_fraction, min_count)
    sel = counts >= thr
    return {text_canonical(key_to_text(k, K)): int(c) for k, c in zip(keys[sel], counts[sel])}, thr
