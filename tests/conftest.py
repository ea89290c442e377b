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
            "small": np.load(os.path.join(d, "oracle_small.npz")),
            "julia_notebook": np.load(os.path.join(d, "julia_notebook.npz"))}


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def julia_replay(golden):
    """Replays the global-RNG stream of the reference notebook notebooks/Mississippi_sharp_null_sims.ipynb (cells
    11-16, Julia 1.4.0) with oracle/julia_rng.py: the seeded draw Ysim and the three blocks of 10 000 x 146 standard
    normals consumed by the two boot_chi2test calls and the boot_mlltest call.  Checked against the committed
    checksums, so a drift of the restated generator cannot go unnoticed."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import julia_rng as J
    jn = golden["julia_notebook"]
    n = 146
    rng = J.MersenneTwister(int(jn["seed"]))
    z0 = J.randn_array(rng, n, bulk=False)
    assert np.array_equal(z0, jn["z0"])
    blocks = []
    for b in range(3):
        z = J.randn_array(rng, n * int(jn["nsim"]), bulk=False)
        chk = np.array([z.sum(), (z * z).sum(), z[0], z[1], z[-2], z[-1]])
        assert np.array_equal(chk, jn["block_checksums"][b])
        blocks.append(np.asfortranarray(z.reshape(int(jn["nsim"]), n).T))      # column s = draw s
    return {"Ysim": jn["Ysim"].copy(), "Z_chi_a": blocks[0], "Z_chi_b": blocks[1], "Z_mll": blocks[2]}
