"""Timings of the fit path (round 2): single / batched Cholesky, GPRealisations (two regions in one call) + cliff face."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
for n in (1024, 2048, 4096, 8192, 16384):
    b, m = ctx.bench_potrf(n, 5)
    print(f"potrf n={n}: best {b:.3f} ms mean {m:.3f} ms  {n**3/3/(b*1e-3)/1e12:.2f} TFLOP/s  ({n**3/3/(b*1e-3)/1e12/37.07*100:.1f} % of DMMA peak)")
for m in (1000, 2000, 4000):
    d = O.synth_two_region(m, 100, seed=1)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
    rd = {True: G.RegionData(d["XT"], d["YT"], np.empty((0, m))), False: G.RegionData(d["XC"], d["YC"], np.empty((0, m)))}
    def seq():
        a = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln); b = G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln); return a, b
    def bat():
        g = G.GPRealisations(rd, k, ln, groupKeys=[True, False]); return g[True], g[False]
    for name, fn in (("two GPE calls", seq), ("GPRealisations (batched)", bat)):
        fn()
        ctx.sync(); t0 = time.time()
        for _ in range(5):
            a, b = fn(); a.close(); b.close()
        ctx.sync(); t_fit = (time.time() - t0) / 5 * 1e3
        ctx.sync(); t0 = time.time()
        for _ in range(5):
            a, b = fn(); G.cliff_face(a, b, d["Xb"]); a.close(); b.close()
        ctx.sync(); t_cf = (time.time() - t0) / 5 * 1e3
        print(f"m={m}/side  {name:26s}: fits {t_fit:.2f} ms   fits + cliff face {t_cf:.2f} ms")
    a1, b1 = seq(); a2, b2 = bat()
    print("   batched == sequential bit for bit:", np.array_equal(a1.alpha, a2.alpha) and np.array_equal(b1.alpha, b2.alpha) and a1.mll == a2.mll and b1.mll == b2.mll)
