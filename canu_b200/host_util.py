"""Small host-side helpers around the companion tool `ovltool` (tools/ and tests only)."""
import os
import subprocess

import numpy as np

_BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin")


def store_read_lengths(store_path):
    """Lengths of reads 1..N of a sqStore as the overlapper sees them (0 = deleted / absent)."""
    out = subprocess.check_output([os.path.join(_BIN, "ovltool"), "lengths", store_path])
    return np.array(out.split(), dtype=np.uint32)
