"""canu_b200 -- B200-native drop-in for the hot path of Canu's `overlapInCore` (ovl overlapper).

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/ovlb200.h),
host/ (the C++ `overlapInCore` replacement executable), api.py (ctypes mirror of the operator
interface) and synth.py (seeded synthetic reads for tests and bench)."""
from .api import (OverlapParams, Overlapper, PackedReads, OvlError, overlap_in_core, load_library,  # noqa: F401
                  RECORD_DTYPE, stats_lines)
