import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and make sure libmafrix_cuda.so exists (nvcc, no GPU needed)."""
    from oracle import oracle
    oracle.build()
    from mafrixraytracing_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    yield


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_goldens.npz"))


def have_gpu():
    from mafrixraytracing_b200 import _lib
    return _lib.load().mfx_device_count() > 0


@pytest.fixture(scope="session")
def big_goldens():
    """Full-size primary-hit checksums of BASELINE configs[3] / [4] (tests/golden/make_goldens_big.py)."""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "big_goldens.npz")))
