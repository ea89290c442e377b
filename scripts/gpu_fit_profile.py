"""Per-kernel-class timing of one GP fit / cliff face / small-RHS solves (latency-bound dataflow chains)."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
for m in (1000, 4000, 8000):
    d = O.synth_two_region(m // 2, 100, seed=1)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    g0 = G.GPE(d["X"], d["Y"], G.MeanZero(), k, math.log(0.3))   # warm
    ctx.profile(True); t0 = time.time()
    g = G.GPE(d["X"], d["Y"], G.MeanZero(), k, math.log(0.3))
    wall = (time.time() - t0) * 1e3
    print(f"n={m} fit wall {wall:.2f} ms:", {nm: round(ctx.profile_get(nm)[0], 3) for nm in ("cov", "potrf", "solve_fwd", "solve_bwd", "misc", "gemm")})
    ctx.profile(True); t0 = time.time()
    G.update_alpha(g)
    print(f"   update_alpha wall {(time.time()-t0)*1e3:.2f} ms:", {nm: round(ctx.profile_get(nm)[0], 3) for nm in ("solve_fwd", "solve_bwd")})
    ctx.profile(False)
gT = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, math.log(0.3)); gC = G.GPE(d["XC"], d["YC"], G.MeanZero(), k, math.log(0.3))
G.cliff_face(gT, gC, d["Xb"])
ctx.profile(True); t0 = time.time(); G.cliff_face(gT, gC, d["Xb"]); wall = (time.time() - t0) * 1e3
print(f"cliff_face (4000/side) wall {wall:.2f} ms:", {nm: round(ctx.profile_get(nm)[0], 3) for nm in ("cov", "solve_fwd", "gemm", "misc")})
