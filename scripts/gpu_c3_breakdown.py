"""C3 (1 000 units/side, 200 sentinels, 10 000 draws): wall time of boot_chi2test by phase and by kernel class."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import georddb200, bench
G = georddb200.load(); ctx = G.default_context()
m3, nb3, S3 = 1000, 200, 10000
XT, XC, YT, YC, Xb = bench.synth(m3, nb3, seed=3)
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
a, b = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln)
CLS = ("cov", "potrf", "solve_fwd", "solve_bwd", "solve_seq", "trmm", "gemm", "rng", "finish", "lu", "misc")
for it in range(6):
    prof = it >= 3
    ctx.sync()
    if prof:
        ctx.profile(True)
    t0 = time.perf_counter()
    nn = G.make_null(a, b)
    ctx.sync(); t1 = time.perf_counter()
    obs = G.chistat(a, b, Xb)
    ctx.sync(); t2 = time.perf_counter()
    nn.set_sentinels(Xb, 0.0) if hasattr(nn, "set_sentinels") else None
    ctx.sync(); t3 = time.perf_counter()
    _, cnt = G.nsim_chi(a, b, nn, Xb, S3, chi_obs=obs, return_count=True, seed=6)
    ctx.sync(); t4 = time.perf_counter()
    line = f"make_null {1e3*(t1-t0):.2f}  chistat_obs {1e3*(t2-t1):.2f}  set_sentinels {1e3*(t3-t2):.2f}  nsim_chi {1e3*(t4-t3):.2f}  total {1e3*(t4-t0):.2f} ms"
    if prof:
        line += "\n      " + "  ".join(f"{c} {ctx.profile_get(c)[0]:.3f}/{ctx.profile_get(c)[1]}" for c in CLS)
        ctx.profile(False)
    print(line)
    nn.close()
def three_step():
    nn = G.make_null(a, b); obs = G.chistat(a, b, Xb)
    _, cnt = G.nsim_chi(a, b, nn, Xb, S3, chi_obs=obs, return_count=True, seed=6); nn.close(); return cnt / S3
for name, fn in (("boot_chi2test", lambda: G.boot_chi2test(a, b, Xb, S3, seed=6)), ("three explicit steps", three_step),
                 ("boot_invvar", lambda: G.boot_invvar(a, b, Xb, S3, seed=6))):
    ts = []
    for _ in range(25):
        ctx.sync(); t0 = time.perf_counter(); p = fn(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(f"{name}: min {ts[0]:.2f}  median {ts[len(ts)//2]:.2f}  max {ts[-1]:.2f} ms   (p = {p})")
# pivoted LU solve alone: unblocked vs blocked kernel (CUDA-event time of the kernel)
rng = np.random.default_rng(0)
for n, nrhs in ((100, 1), (200, 1), (200, 202), (604, 1)):
    A = rng.standard_normal((n, n)); B = rng.standard_normal((n, nrhs))
    for mode in (0, 1):
        ctx.lib.geordd_dbg_set_lu_blocked(mode)
        G.generic_solve(A, B)
        ctx.profile(True)
        for _ in range(5):
            G.generic_solve(A, B)
        ms, cnt = ctx.profile_get("lu"); ctx.profile(False)
        print(f"LU solve n={n} nrhs={nrhs} {'blocked  ' if mode else 'unblocked'}: {ms / cnt:.3f} ms")
ctx.lib.geordd_dbg_set_lu_blocked(1)
