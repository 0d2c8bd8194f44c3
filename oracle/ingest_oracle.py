"""TEST INFRASTRUCTURE -- a CPU (numpy) restatement of the overlap-store ingest step, the checker of
ovlb_ingest_records.  Only tests/ may import it.

What the reference does with every overlap an overlapper wrote, before the store is written (sequential build:
stores/ovStoreBuild.C:176-263; parallel build: ovStoreBucketizer.C:205-230 + ovStoreSorter.C:223):

  * ovStoreFilter::filterOverlap (stores/ovStoreFilter.C:71-150): IDs must be in 1..maxID (else the reference exits);
    the mirrored twin r = swapIDs(f) is made; if evalue > AS_OVS_encodeEvalue(maxErate) the forUTG/forOBT/forDUP flags
    of BOTH are cleared (the skipReadOBT table is all false in this version: its loop is `#if 0`);
  * ovOverlap::swapIDs (stores/ovOverlap.C:215-246): a/b IDs swap; not flipped: (ahg5,ahg3,bhg5,bhg3) <-
    (bhg5,bhg3,ahg5,ahg3); flipped: <- (bhg3,bhg5,ahg3,ahg5); every other field is copied;
  * a record is kept iff it still carries one of the three flags;
  * std::sort with ovOverlap::operator< (stores/ovOverlap.H:265-279): a_iid, b_iid, dat[0], dat[1] ascending (unsigned).

Record layout (stores/ovOverlap.H:49-67, AS_MAX_READLEN_BITS = 21): dat[0] = ahg5:21 | ahg3:21 | evalue:16 | flipped |
forOBT | forDUP | forUTG | 2 spare (LSB first); dat[1] = bhg5:21 | bhg3:21 | span:21 | 1 spare.

Pinned by tests/golden/ingest_*.bin.gz, minted by tests/golden/make_ingest_golden.py from the reference's own code.
"""
import numpy as np

RECORD_DTYPE = np.dtype([("a_iid", "<u4"), ("b_iid", "<u4"), ("w0", "<u8"), ("w1", "<u8")])
_M21 = np.uint64((1 << 21) - 1)
_FLAGS = np.uint64(0b111 << 59)          # forOBT (59), forDUP (60), forUTG (61)


def encode_evalue(erate: float) -> int:
    """AS_OVS_encodeEvalue (stores/ovOverlap.H:31-35)."""
    return int(100000.0 * erate + 0.5) if erate < 65535 / 100000.0 else 65535


def swap_ids(recs: np.ndarray) -> np.ndarray:
    out = recs.copy()
    out["a_iid"], out["b_iid"] = recs["b_iid"], recs["a_iid"]
    w0, w1 = recs["w0"], recs["w1"]
    ahg5, ahg3 = w0 & _M21, (w0 >> np.uint64(21)) & _M21
    bhg5, bhg3 = w1 & _M21, (w1 >> np.uint64(21)) & _M21
    flipped = ((w0 >> np.uint64(58)) & np.uint64(1)).astype(bool)
    na5 = np.where(flipped, bhg3, bhg5); na3 = np.where(flipped, bhg5, bhg3)
    nb5 = np.where(flipped, ahg3, ahg5); nb3 = np.where(flipped, ahg5, ahg3)
    keep0 = w0 & ~np.uint64((1 << 42) - 1)
    keep1 = w1 & ~np.uint64((1 << 42) - 1)
    out["w0"] = keep0 | na5 | (na3 << np.uint64(21))
    out["w1"] = keep1 | nb5 | (nb3 << np.uint64(21))
    return out


def ingest(recs: np.ndarray, max_evalue: int, max_id: int) -> np.ndarray:
    """records (RECORD_DTYPE) -> mirrored, filtered, sorted records."""
    if recs.size and (recs["a_iid"].min() == 0 or recs["b_iid"].min() == 0 or
                      recs["a_iid"].max() > max_id or recs["b_iid"].max() > max_id):
        raise ValueError("Overlap has IDs out of range")
    both = np.empty(2 * recs.size, dtype=RECORD_DTYPE)
    both[0::2] = recs
    both[1::2] = swap_ids(recs)
    evalue = (both["w0"] >> np.uint64(42)) & np.uint64(0xFFFF)
    bad = evalue > np.uint64(max_evalue)
    both["w0"][bad] &= ~_FLAGS
    both = both[(both["w0"] & _FLAGS) != 0]
    return np.sort(both, order=["a_iid", "b_iid", "w0", "w1"])
