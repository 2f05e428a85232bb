"""The two Green's-function tables the reference reads at start-up (kernel_initialization.f90:15,344).

Stored as float32 .npy [k][j][i][component]; regenerate with tests/golden/make_kernel_tables.py."""
import os
import numpy as np

_D = os.path.join(os.path.dirname(__file__), "data")


def fine_table() -> np.ndarray:
    t = np.load(os.path.join(_D, "wfxyzf3.npy"))
    assert t.shape == (16, 16, 16, 3) and t.dtype == np.float32
    return np.ascontiguousarray(t)


def coarse_table() -> np.ndarray:
    t = np.load(os.path.join(_D, "wfxyzc2.npy"))
    assert t.shape == (4, 4, 4, 3) and t.dtype == np.float32
    return np.ascontiguousarray(t)
