"""Cycle counts of the shared-memory tile operations (critical path of the dataflow Cholesky / solves)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import georddb200
G = georddb200.load(); ctx = G.default_context()
out = (C.c_double * 10)()
ctx.check(ctx.lib.geordd_bench_tileops(ctx._h, out))
for nm, v in zip(("potrf_tile", "trsm_right_tile", "solve_fwd_tile(64)", "solve_bwd_tile(64)"), out):
    print(f"{nm}: {v:.0f} cycles = {v / 1.965e3:.2f} us")
print("potrf phases (diag, panel, update):", list(out)[4:7], " trsm phases (load, subst, update):", list(out)[7:10])
for n in (1024, 2048, 4096, 8192):
    print(n, ctx.bench_potrf(n, 3))
