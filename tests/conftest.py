import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure only)."""
    import geordd_oracle
    return geordd_oracle


@pytest.fixture(scope="session")
def G():
    """The product package (geordd.jl_b200), loaded through the repo-root loader."""
    import georddb200
    return georddb200.load()


@pytest.fixture(scope="session")
def ctx(G):
    """A GPU context; GPU tests fail loudly (no skip, no fallback) if the device/library is missing."""
    return G.default_context()


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {"mississippi": np.load(os.path.join(d, "mississippi.npz")),
            "small": np.load(os.path.join(d, "oracle_small.npz"))}


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
