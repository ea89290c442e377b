"""Exact replay (CPU, no Julia, ~10 min) of cells 19-28 of the reference notebook
notebooks/Mississippi_sharp_null_sims.ipynb with the restated Julia RNG (oracle/julia_rng.py): three vectors of 100 000
null statistics (Random.seed!(1)), then nsim_power under the null and under tau = 1.2 (10 000 draws each, Random.seed!(1)
before each), and the size / power table of cell 28 = sprintf("%.2f", mean(p < 0.05)).  Needs the committed fixtures
only (tests/golden/mississippi.npz, julia_notebook.npz).  Prints the ten rates for the corrected and the as-shipped
analytic calibration next to the notebook's table and writes tests/golden/julia_table_replay.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import geordd_oracle as O, julia_rng as J

G = os.path.join(ROOT, "tests", "golden")
ms, jn = np.load(os.path.join(G, "mississippi.npz")), np.load(os.path.join(G, "julia_notebook.npz"))
spec = O.KernelSpec(O.SE_ISO, 100e3, 1.0, 100.0 ** 2, 0.0)
y, Xb, n = jn["Ysim"], ms["sentinels"], 146
gT = O.gp_fit(spec, np.asfortranarray(ms["X_MS"]), y[:82].copy(), 0.0)
gC = O.gp_fit(spec, np.asfortranarray(ms["X_LA"]), y[82:].copy(), 0.0)
t0 = time.time()


def normals(rng, nsim):
    return np.asfortranarray(J.randn_array(rng, n * nsim, bulk=False).reshape(nsim, n).T)


rng = J.MersenneTwister(1)                                               # cell 19
p_invvar = O.nsim_invvar_pval_uncalib(gT, gC, Xb, normals(rng, 100000))
print("cell 19 done", round(time.time() - t0), "s", flush=True)
mn, ma = O.nsim_logP(O.modifiable(gT), O.modifiable(gC), O.make_null(gT, gC), normals(rng, 100000))   # cell 20
print("cell 20 done", round(time.time() - t0), "s", flush=True)
chi = O.nsim_chi(gT, gC, O.make_null(gT, gC), Xb, normals(rng, 100000))                                  # cell 21
print("cell 21 done", round(time.time() - t0), "s", flush=True)
out = {"notebook": {"null": [0.05, 0.05, 0.09, 0.05, 0.05], "altv": [0.66, 0.63, 0.87, 0.80, 0.80]}}
for shipped in (False, True):
    res = {}
    for name, tau in (("altv", 1.2), ("null", 0.0)):                                                       # cells 23, 24
        Z = normals(J.MersenneTwister(1), 10000)
        ps = O.nsim_power(gT, gC, tau, Xb, chi, ma - mn, p_invvar, Z, as_shipped_calib=shipped)
        res[name] = [float(np.mean(ps[:, i] < 0.05)) for i in range(5)]
    out["as_shipped" if shipped else "corrected"] = res
    print("as_shipped" if shipped else "corrected", res, round(time.time() - t0), "s", flush=True)
json.dump(out, open(os.path.join(G, "julia_table_replay.json"), "w"), indent=1)
print("notebook", out["notebook"])
