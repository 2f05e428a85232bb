import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library (cross-compiles without a GPU) and the oracle once per session."""
    import __graft_entry__ as g
    g.build()
    return True


def sort_records(a):
    import numpy as np
    a = a.reshape(-1, a.shape[-1])
    return a[np.lexsort(tuple(a[:, c] for c in range(a.shape[1] - 1, -1, -1)))]
