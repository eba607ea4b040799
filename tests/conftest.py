import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def f64(d):
    return {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}


@pytest.fixture(scope="session")
def golden_small():
    with np.load(os.path.join(GOLDEN, "small.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_default():
    with np.load(os.path.join(GOLDEN, "default_cfg.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def conv_kat():
    with np.load(os.path.join(GOLDEN, "conv_kat.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def lib():
    """Builds (if needed) and loads libsrwn.so."""
    import __graft_entry__ as ge
    ge.build()
    import sr_wavenet_b200 as srwn
    return srwn._lib.load()
