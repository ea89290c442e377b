"""Cycle counts of the shared-memory tile operations on the Cholesky / solve critical paths (one CTA)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200
G = georddb200.load(); ctx = G.default_context()
out = (C.c_double * 10)()
ctx.check(ctx.lib.geordd_bench_tileops(ctx._h, out))
names = ["potrf_tile", "trsm_right_tile", "solve_fwd_tile", "solve_bwd_tile", "potrf:diag32", "potrf:panel", "potrf:update",
         "trsmR:panel_load", "trsmR:subst", "trsmR:update"]
for n, v in zip(names, out):
    print(f"{n:18s} {v:10.0f} cycles")
for n in (2048, 4096, 8192, 16384):
    b, m = ctx.bench_potrf(n, 3)
    print(f"potrf n={n}: best {b:.3f} ms mean {m:.3f} ms  {n**3/3/(b*1e-3)/1e12:.2f} TFLOP/s")
