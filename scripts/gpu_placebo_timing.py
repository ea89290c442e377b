"""Placebo sweep of the NYC notebook shape (notebooks/NYC_analysis-2015.ipynb:1452-1512: 90 angles x 1000 draws on one
district; the notebook's @time outputs are 73.5 / 917.1 s (mll, districts A / B) and 36.6 / 430.0 s (chi2)): wall time on one
B200 with 1, 2, 4, 8 concurrent contexts."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
angles = [float(a) for a in range(1, 180, 2)]
k = G.Mat12Iso(math.log(0.3), 0.0) + G.Const(math.log(5.0))       # the NYC notebook fits Matern-1/2 + Const
for n in (1000, 3000):
    d = O.synth_two_region(n, 10, seed=7, tau=0.0)
    X, Y = np.asfortranarray(d["X"][:, :n]), d["Y"][:n].copy()
    for which in ("mll", "chi"):
        ref = None
        for w in (1, 2, 4, 8):
            G.placebo_sweep(which, angles[:w], X, Y, k, G.MeanZero(), math.log(0.3), 1000, seed=1, workers=w)   # warm-up (contexts, pools)
            t0 = time.time()
            p = G.placebo_sweep(which, angles, X, Y, k, G.MeanZero(), math.log(0.3), 1000, seed=1, workers=w)
            dt = time.time() - t0
            ref = p if ref is None else ref
            print(f"district of {n} units, placebo_{which}, 90 angles x 1000 draws, workers={w}: {dt:.2f} s  same p-values: {np.array_equal(p, ref)}", flush=True)
