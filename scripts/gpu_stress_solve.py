"""Stress test of the batched dataflow solves: 20 000 right-hand sides against SciPy, every column checked.
(This is the test that exposed the stage-release race with TMA refills, see mainloop.cuh.)"""
import ctypes as C, os, sys
import numpy as np, scipy.linalg as sl
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context(); lib = ctx.lib; capi = G.capi
rng = np.random.default_rng(0)
for n in (1000, 2000):
    X = rng.random((2, n)); K = O.cov_sym(O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0), X, 0.09)
    L = np.asfortranarray(np.linalg.cholesky(K))
    for nr in (5632, 20000, 20000):
        B = np.asfortranarray(rng.standard_normal((n, nr)))
        for mode, ref in ((2, L @ B), (0, sl.solve_triangular(L, B, lower=True)), (1, sl.solve_triangular(L.T, B, lower=False)), (3, sl.solve_triangular(L, B, lower=True)), (4, sl.solve_triangular(L.T, B, lower=False))):
            out = B.copy(order="F"); rc = lib.geordd_dbg_trsm(ctx._h, mode, capi.dptr(L), n, capi.dptr(out), nr, None); assert rc == 0
            err = np.abs(out - ref) / np.max(np.abs(ref)); bad = np.nonzero(err.max(axis=0) > 1e-9)[0]
            print(f"n={n} nrhs={nr} mode={mode}: max err {err.max():.2e}, bad columns {len(bad)}")
            assert len(bad) == 0
print("STRESS OK")
