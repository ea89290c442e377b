"""Cliff-face fit at the bench size (two fits of 4 000 units + cliff face at 100 sentinels): wall time by phase and CUDA-event time by
kernel class."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200, bench
G = georddb200.load(); ctx = G.default_context()
XT, XC, YT, YC, Xb = bench.synth()
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
CLS = ("cov", "potrf", "solve_fwd", "solve_bwd", "gemm", "lu", "misc")
for it in range(8):
    prof = it >= 4
    ctx.sync()
    if prof:
        ctx.profile(True)
    t0 = time.perf_counter()
    a, b = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln)
    ctx.sync(); t1 = time.perf_counter()
    mu, Sig = G.cliff_face(a, b, Xb)
    ctx.sync(); t2 = time.perf_counter()
    line = f"fits {1e3 * (t1 - t0):.2f} ms   cliff face {1e3 * (t2 - t1):.2f} ms"
    if prof:
        line += "   | " + "  ".join(f"{c} {ctx.profile_get(c)[0]:.3f}/{ctx.profile_get(c)[1]}" for c in CLS)
        ctx.profile(False)
    print(line)
    a.close(); b.close()
