"""One pivoted-LU solve of the C3 size (200 x 200, 2 right-hand sides) for an ncu capture of lu_blocked_kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200
G = georddb200.load(); ctx = G.default_context()
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
A = rng.standard_normal((n, n)); B = rng.standard_normal((n, 2))
for _ in range(3):
    G.generic_solve(A, B)
print("ok")
