"""C5-scale smoke check (n = 16 000, S_null = 1 000, 500 power draws): the largest configuration the
reference names.  Checks the z'z identity of the pooled null GP for every draw, finiteness / plausibility of all
statistics, and prints timings."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
m = 8000
d = O.synth_two_region(m, 100, seed=1)
k = G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
t0 = time.time()
gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln)
gN = G.NullGPE(gT, gC, d["Xb"], 0.0)
print(f"fits T, C (8000) + pooled null (16000): {time.time()-t0:.2f} s; mll {gT.mll:.3f} {gC.mll:.3f} {gN.mll:.3f}")
mu, Sig = G.cliff_face(gT, gC, d["Xb"])
print("cliff face: mu range", mu.min(), mu.max(), "Sigma sym err", np.abs(Sig - Sig.T).max(), "min eig", np.linalg.eigvalsh(Sig).min())
S = 1000
t0 = time.time(); _, _, mnull, malt = G.nsim_logP(gT, gC, gN, S, seed=3, return_count=True); t1 = time.time() - t0
Z = np.zeros((2 * m, S), order="F"); ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 3, 0, S, 2 * m, G.capi.dptr(Z)))
v = mnull + np.sum(Z * Z, axis=0) / 2.0; const = np.median(v)
print(f"mll null sims: {t1*1e3:.1f} ms; identity rel err max {np.max(np.abs(v-const))/abs(const):.3e}; dmll mean {np.mean(malt-mnull):.3f}")
t0 = time.time(); chi = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=3); t2 = time.time() - t0
pinv = G.nsim_invvar_pval_uncalib(gT, gC, d["Xb"], S, seed=3)
print(f"chi2 null sims: {t2*1e3:.1f} ms; mean {chi.mean():.2f} (nb=100); invvar p mean {np.mean(pinv):.3f}")
t0 = time.time(); pw = G.nsim_power(gT, gC, 0.3, d["Xb"], chi, malt - mnull, pinv, 500, seed=4); t3 = time.time() - t0
print(f"power: 500 draws x 5 p-values in {t3:.2f} s; finite {np.isfinite(pw).all()}; rejection rates at 0.05: {(pw < 0.05).mean(axis=0)}")
assert np.max(np.abs(v - const)) < 1e-8 * abs(const) and np.isfinite(chi).all() and np.isfinite(pw).all()
print("C5 OK")
