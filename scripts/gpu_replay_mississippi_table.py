"""The exact replay of scripts/replay_mississippi_table.py through the CUDA path: the notebook's own normals (restated
Julia RNG) are passed as host Z to nsim_invvar_pval_uncalib / nsim_logP / nsim_chi (100 000 draws each) and nsim_power
(10 000 draws, tau = 0 and 1.2); prints the size / power table next to the notebook's."""
import json, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, julia_rng as J
G = georddb200.load(); ctx = G.default_context()
Gd = os.path.join(ROOT, "tests", "golden")
ms, jn = np.load(os.path.join(Gd, "mississippi.npz")), np.load(os.path.join(Gd, "julia_notebook.npz"))
k = G.SEIso(math.log(100e3), 0.0) + G.Const(math.log(100.0))
y, Xb, n = jn["Ysim"], ms["sentinels"], 146
gT = G.GPE(ms["X_MS"], y[:82].copy(), G.MeanZero(), k, 0.0)
gC = G.GPE(ms["X_LA"], y[82:].copy(), G.MeanZero(), k, 0.0)
t0 = time.time()
def normals(rng, nsim):
    return np.asfortranarray(J.randn_array(rng, n * nsim, bulk=False).reshape(nsim, n).T)
rng = J.MersenneTwister(1)
p_invvar = G.nsim_invvar_pval_uncalib(gT, gC, Xb, 0, Z=normals(rng, 100000))
gN = G.make_null(gT, gC)
_, _, mn, ma = G.nsim_logP(gT, gC, gN, 0, Z=normals(rng, 100000), return_count=True)
chi = G.nsim_chi(gT, gC, G.make_null(gT, gC), Xb, 0, Z=normals(rng, 100000))
res = {}
for name, tau in (("altv", 1.2), ("null", 0.0)):
    ps = G.nsim_power(gT, gC, tau, Xb, chi, ma - mn, p_invvar, 0, Z=normals(J.MersenneTwister(1), 10000))
    res[name] = [float(np.mean(ps[:, i] < 0.05)) for i in range(5)]
ref = json.load(open(os.path.join(Gd, "julia_table_replay.json")))
print("GPU      ", res, round(time.time() - t0), "s")
print("oracle   ", ref["corrected"])
print("notebook ", ref["notebook"])
ok = all(f"{res[c][i]:.2f}" == f"{ref['notebook'][c][i]:.2f}" for c in ("null", "altv") for i in range(5))
print("GPU TABLE == NOTEBOOK TABLE" if ok else "MISMATCH")
