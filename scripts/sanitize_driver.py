"""Small-shape pass over the hot path for compute-sanitizer (scripts/sanitize.sh): every kernel class once, including
the shape of ADVICE r1 (mT % 128 != 0, small mC % 128, nsim a multiple of 64 that fills the batch exactly)."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
d = O.synth_two_region(145, 40, seed=1)
mT, mC = 140, 150                                   # 140 % 128 = 12, 150 % 128 = 22, 12 + 22 <= 128
XT, XC, YT, YC = d["X"][:, :mT], d["X"][:, mT:], d["Y"][:mT], d["Y"][mT:]
k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
gT, gC = G.GPE.fit_many([XT, XC], [YT, YC], G.MeanZero(), k, ln)          # batched Cholesky (merged job list)
g1 = G.GPE(XT, YT, G.MeanZero(), k, ln)                                    # single fit
assert g1.mll == gT.mll
mu, S = G.cliff_face(gT, gC, d["Xb"])
gN = G.NullGPE(gT, gC, d["Xb"], 0.0)                                       # pooled fit with the treated prefix reused
ctx.set_max_batch(64)
for nsim in (64, 128):                                                      # batch filled exactly: last slab = last columns of the buffers
    chi = G.nsim_chi(gT, gC, gN, d["Xb"], nsim, seed=1)
    _, cnt, mn, ma = G.nsim_logP(gT, gC, gN, nsim, seed=1, dmll_obs=0.0, return_count=True)
    pv = G.nsim_invvar_pval_uncalib(gT, gC, d["Xb"], nsim, seed=1, gpNull=gN)
    assert np.isfinite(chi).all() and np.isfinite(mn).all() and np.isfinite(ma).all() and np.isfinite(pv).all()
pc = G.nsim_invvar_calib(gT, gC, d["Xb"], 64, seed=2)
pw = G.nsim_power(gT, gC, 0.3, d["Xb"], chi, ma - mn, pv, 64, seed=3)
w = G.weight_at_units(gT, d["Xb"], np.ones(40))
Xp, dist = G.projection_points_all(gT, np.column_stack([np.linspace(0, 1, 50), np.linspace(0, 1, 50)]))
# more sentinels than units per side (the Mississippi fixture's shape: 82 / 64 units, 200 sentinels)
d2 = O.synth_two_region(73, 200, seed=2)
hT, hC = G.GPE.fit_many([d2["X"][:, :82], d2["X"][:, 82:]], [d2["Y"][:82], d2["Y"][82:]], G.MeanZero(), k, ln)
hN = G.NullGPE(hT, hC, d2["Xb"], 0.0)
chi2 = G.nsim_chi(hT, hC, hN, d2["Xb"], 64, seed=1)
_, _, mn2, ma2 = G.nsim_logP(hT, hC, hN, 64, seed=1, return_count=True)
pv2 = G.nsim_invvar_pval_uncalib(hT, hC, d2["Xb"], 64, seed=1)
pw2 = G.nsim_power(hT, hC, 0.2, d2["Xb"], chi2, ma2 - mn2, pv2, 64, seed=3)
pc2 = G.pval_invvar_calib(hT, hC, d2["Xb"]); pcs2 = G.nsim_invvar_calib(hT, hC, d2["Xb"], 64, seed=4)
assert np.isfinite(pw2).all() and np.isfinite(pcs2).all() and math.isfinite(pc2)
hN.close(); hT.close(); hC.close()
rd = {True: G.RegionData(XT, YT, d["D"][:, :mT]), False: G.RegionData(XC, YC, d["D"][:, mT:])}
mg = G.MultiGPCovars(rd, G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0)), G.LinIso(0.0), 1.0, groupKeys=[True, False])
G.update_mll_and_dmll(mg); G.postmean_beta(mg)
assert np.isfinite(pc).all() and np.isfinite(pw).all() and np.isfinite(w).all() and np.isfinite(Xp).all() and np.isfinite(mu).all() and np.isfinite(S).all()
gN.close(); gT.close(); gC.close(); g1.close(); mg.close()
import gc; gc.collect()
v = ctx.lib.geordd_dbg_guard_violations()
print("sanitize driver ok", float(chi.mean()), float(pw.mean()), "guard mode" if os.environ.get("GEORDD_GUARD") else "", "guard violations:", v)
assert v == 0
