"""Where the end-to-end step (3 fits from host buffers + 20 000 mll draws + D2H) spends its time."""
import math, os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, georddb200
G = georddb200.load(); ctx = G.Context(0)
XT, XC, YT, YC, Xb = bench.synth()
k = G.Mat32Iso(math.log(bench.ELL), 0.5 * math.log(bench.SIG2)) + G.Const(0.5 * math.log(bench.CONST_S2))
ln = math.log(bench.NOISE_SD); S = 20000
out_null, out_alt = np.empty(S), np.empty(S)
def step():
    t = [time.time()]
    a = G.GPE(XT, YT, G.MeanZero(), k, ln, ctx=ctx); t.append(time.time())
    b = G.GPE(XC, YC, G.MeanZero(), k, ln, ctx=ctx); t.append(time.time())
    nn = G.NullGPE(a, b); t.append(time.time())
    cnt = C.c_longlong()
    ctx.check(ctx.lib.geordd_null_mll(nn._h, S, 1, 0, None, 0.0, G.capi.dptr(out_null), G.capi.dptr(out_alt), C.byref(cnt), None, None)); t.append(time.time())
    nn.close(); a.close(); b.close(); t.append(time.time())
    return np.diff(t) * 1e3
step(); step()
ctx.profile(True)
d = step()
print("wall ms: GPE(T) %.2f GPE(C) %.2f NullGPE %.2f sims %.2f close %.2f total %.2f" % (*d, d.sum()))
print({nm: round(ctx.profile_get(nm)[0], 2) for nm in ("cov", "potrf", "solve_fwd", "solve_bwd", "solve_seq", "trmm", "gemm", "rng", "finish", "lu", "misc")})
