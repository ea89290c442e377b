import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import georddb200, geordd_oracle as O
G = georddb200.load(); ctx = G.default_context()
for m in (1000, 2000, 4000):
    d = O.synth_two_region(m, 100, seed=1)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
    gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln)
    gN = G.NullGPE(gT, gC, d["Xb"], 0.0)
    S = 256
    _, _, mnull, malt = G.nsim_logP(gT, gC, gN, S, seed=11, return_count=True)
    Z = np.zeros((2 * m, S), order="F"); ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 11, 0, S, 2 * m, G.capi.dptr(Z)))
    zz = np.sum(Z * Z, axis=0); v = mnull + zz / 2.0; const = np.median(v)
    err = np.abs(v - const) / abs(const)
    print(f"m={m}: identity rel err max {err.max():.3e}, #bad(>1e-8) {int((err>1e-8).sum())}, bad idx {np.nonzero(err>1e-8)[0][:10]}")
    # T and C parts vs CPU on a few draws
    spec = O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0)
    if m <= 2000:
        oT, oC = O.gp_fit(spec, d["XT"], d["YT"], ln), O.gp_fit(spec, d["XC"], d["YC"], ln)
        mn_o, ma_o = O.nsim_logP(oT, oC, O.make_null(oT, oC), Z[:, :8])
        print("   null rel", np.abs(mnull[:8]-mn_o).max()/abs(mn_o).max(), "alt rel", np.abs(malt[:8]-ma_o).max()/abs(ma_o).max())
    t0=time.time(); chi = G.nsim_chi(gT, gC, gN, d["Xb"], 20000, seed=11); t1=time.time()-t0
    t0=time.time(); chi = G.nsim_chi(gT, gC, gN, d["Xb"], 20000, seed=11); t2=time.time()-t0
    print(f"   chi2 20000 draws: first {t1*1e3:.1f} ms, second {t2*1e3:.1f} ms; mean {chi.mean():.3f}")
    ctx.profile(True); G.nsim_chi(gT, gC, gN, d["Xb"], 20000, seed=12)
    print("   ", {nm: round(ctx.profile_get(nm)[0], 2) for nm in ("trmm","solve_fwd","solve_bwd","gemm","rng","finish","misc")}); ctx.profile(False)
