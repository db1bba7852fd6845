import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files(pattern="*.npz"):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def lib_built():
    """libdrops_b200.so, built on demand (nvcc cross-compiles on the CPU box)."""
    from drops_b200 import _build
    return _build.build()


@pytest.fixture(scope="session")
def gpu(lib_built):
    """Initialised CUDA path.  Fails loudly if the extension or the device is missing."""
    from drops_b200 import capi
    capi.init(int(os.environ.get("LOCAL_RANK", "0")))
    return capi


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.build()
    return p


@pytest.fixture(scope="session")
def refmod():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libdrops_ref.so not built (needs /root/reference)")
    return ref


def hash_vector(n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.float64) / 4294967296.0 - 0.5


def absAx(rowbeg, colind, val, x, rows):
    """(|A||x|)_i: the scale the SpMV tolerance 1e-14 is relative to."""
    import scipy.sparse as sp
    A = sp.csr_matrix((np.abs(val), colind.astype(np.int64), rowbeg.astype(np.int64)), shape=(rows, len(x)))
    return A @ np.abs(x)
