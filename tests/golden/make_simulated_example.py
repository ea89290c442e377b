"""Regenerates tests/golden/simulated_example.npz: the data of the reference notebook notebooks/SimulatedExample.ipynb
(cell 5, `Random.seed!(1)`), replayed with oracle/julia_rng.py -- see oracle/simulated_example.py.  Run from the repo root:
    python tests/golden/make_simulated_example.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import simulated_example as SE  # noqa: E402

if __name__ == "__main__":
    d = SE.notebook_data()
    np.savez_compressed(os.path.join(HERE, "simulated_example.npz"), **d)
    print({k: (v.shape, v.dtype) for k, v in d.items()}, "treated:", int(d["treat"].sum()))
    print("X[:, :2] =", [repr(float(v)) for v in d["X"][:, :2].ravel(order="F")])
