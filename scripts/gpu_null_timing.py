"""Where make_null's wall time goes at the bench size (m = 4 000 per side): per-kernel-class CUDA-event times vs the wall clock."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200, bench
G = georddb200.load(); ctx = G.default_context()
XT, XC, YT, YC, Xb = bench.synth()
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
a, b = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln)
CLS = ("cov", "potrf", "solve_fwd", "solve_bwd", "misc", "gemm", "lu", "finish")
for prof in (False, True, False):
    for it in range(4):
        ctx.sync()
        if prof:
            ctx.profile(True)
        t0 = time.perf_counter()
        nn = G.NullGPE(a, b)
        ctx.sync(); t1 = time.perf_counter()
        line = f"profile={prof}  make_null wall {1e3 * (t1 - t0):6.2f} ms"
        if prof:
            line += "   " + "  ".join(f"{c} {ctx.profile_get(c)[0]:.3f}/{ctx.profile_get(c)[1]}" for c in CLS)
            ctx.profile(False)
        print(line)
        nn.close()
