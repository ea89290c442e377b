"""Generates the committed fixtures under tests/golden/.  Run in the build container (it reads the
reference's data files, which do not exist on the GPU box):

    python tests/golden/make_golden.py

mississippi.npz  -- the reference's own fixture (notebooks/Mississippi_data/{X_MS,X_LA,border}.csv) with
                    the 200 sentinels of notebooks/Mississippi_sharp_null_sims.ipynb cell 7 placed by the
                    oracle's restatement of GeoRDD.sentinels (src/geometry.jl:22-30).  Coordinates only
                    (DATA, not source code); the notebook draws Y from the null prior.
oracle_small.npz -- outputs of the CPU oracle on a seeded synthetic two-region problem (SURVEY §8d shapes,
                    scaled down), used to detect drift of the oracle itself and as a second target for
                    the GPU parity tests.  NOT reference outputs.
julia_notebook.npz -- the REFERENCE'S OWN printed outputs of notebooks/Mississippi_sharp_null_sims.ipynb (Julia 1.4.0,
                    executed top to bottom: cells 11-16) together with what is needed to replay them without Julia:
                    `Random.seed!(4); Ysim = prior_rand(gpNull)` and the three bootstrap tests that follow consume the
                    global MersenneTwister stream; oracle/julia_rng.py restates that generator, this fixture stores the
                    replayed Ysim and checksums of the three blocks of 10 000 x 146 normals (so a drift of the restated
                    generator is detected), and the printed numbers: pval_invvar_uncalib 0.006981840916959539,
                    pval_invvar_calib 0.01977244834968639, boot_chi2test 0.0754 and 0.0759, boot_mlltest 0.0354.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import geordd_oracle as O  # noqa: E402

REF = "/root/reference/notebooks/Mississippi_data"


def mississippi():
    X_MS = np.loadtxt(os.path.join(REF, "X_MS.csv"), delimiter=",")
    X_LA = np.loadtxt(os.path.join(REF, "X_LA.csv"), delimiter=",")
    border = np.loadtxt(os.path.join(REF, "border.csv"), delimiter=",")
    assert X_MS.shape == (2, 82) and X_LA.shape == (2, 64) and border.shape == (1279, 2)
    sent = O.polyline_sentinels(border, 200)
    np.savez_compressed(os.path.join(HERE, "mississippi.npz"), X_MS=X_MS, X_LA=X_LA, sentinels=sent)
    print("mississippi.npz", X_MS.shape, X_LA.shape, sent.shape)


def oracle_small():
    out = {}
    for tag, kind, ell, nb in (("mat32", O.MAT32, 0.3, 40), ("se", O.SE_ISO, 0.3, 20)):
        d = O.synth_two_region(90, nb, seed=7)
        # ragged split: 100 treated / 80 control
        XT, XC = d["X"][:, :100], d["X"][:, 100:]
        YT, YC = d["Y"][:100], d["Y"][100:]
        spec = O.KernelSpec(kind, ell, 1.0, 400.0, 0.0)
        ln = math.log(0.3)
        gT, gC = O.gp_fit(spec, XT, YT, ln), O.gp_fit(spec, XC, YC, ln)
        mu, S = O.cliff_face(gT, gC, d["Xb"])
        Z = np.random.default_rng(11).standard_normal((180, 96))
        p_chi, c_chi, chi_obs, chi_sims = O.boot_chi2test(gT, gC, d["Xb"], Z)
        p_mll, c_mll, d_obs, d_sims = O.boot_mlltest(gT, gC, Z)
        p_inv, c_inv, pobs, p_sims = O.boot_invvar(gT, gC, d["Xb"], Z)
        out.update({f"{tag}_XT": XT, f"{tag}_XC": XC, f"{tag}_YT": YT, f"{tag}_YC": YC, f"{tag}_Xb": d["Xb"],
                    f"{tag}_Z": Z, f"{tag}_alphaT": gT.alpha, f"{tag}_mllT": gT.mll, f"{tag}_mllC": gC.mll,
                    f"{tag}_mu": mu, f"{tag}_Sigma": S, f"{tag}_chi_obs": chi_obs, f"{tag}_chi_sims": chi_sims,
                    f"{tag}_chi_cnt": c_chi, f"{tag}_dmll_obs": d_obs, f"{tag}_dmll_sims": d_sims,
                    f"{tag}_mll_cnt": c_mll, f"{tag}_pinv_obs": pobs, f"{tag}_pinv_sims": p_sims,
                    f"{tag}_inv_cnt": c_inv, f"{tag}_calib": O.pval_invvar_calib(gT, gC, d["Xb"]),
                    f"{tag}_spec": np.array([kind, ell, 1.0, 400.0, 0.0, ln])})
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
    print("oracle_small.npz", len(out), "arrays")


def julia_notebook():
    import julia_rng as J
    ms = np.load(os.path.join(HERE, "mississippi.npz"))
    n = ms["X_MS"].shape[1] + ms["X_LA"].shape[1]
    spec = O.KernelSpec(O.SE_ISO, 100e3, 1.0, 100.0 ** 2, 0.0)          # SEIso(log(100e3), 0) + Const(log(100)), cell 9
    g0T = O.gp_fit(spec, np.asfortranarray(ms["X_MS"]), np.zeros(82), 0.0)   # cell 10: GP(X, zeros, MeanZero(), k2, 0.0)
    g0C = O.gp_fit(spec, np.asfortranarray(ms["X_LA"]), np.zeros(64), 0.0)
    g0N = O.make_null(g0T, g0C)
    rng = J.MersenneTwister(4)                                             # cell 11
    z0 = J.randn_array(rng, n, bulk=False)
    y = g0N.U.T @ z0
    y = y - np.mean(y)
    y[:82] += 1.2
    sums = []
    for _ in range(3):                                                     # cells 14, 15 (chi2) and 16 (mll)
        z = J.randn_array(rng, n * 10000, bulk=False)
        sums.append([z.sum(), (z * z).sum(), z[0], z[1], z[-2], z[-1]])
    np.savez_compressed(os.path.join(HERE, "julia_notebook.npz"), seed=4, tau=1.2, z0=z0, Ysim=y,
                        block_checksums=np.array(sums), pval_invvar_uncalib=0.006981840916959539,
                        pval_invvar_calib=0.01977244834968639, boot_chi2=np.array([0.0754, 0.0759]), boot_mll=0.0354,
                        nsim=10000)
    print("julia_notebook.npz", y.shape, np.array(sums)[:, :2])


if __name__ == "__main__":
    mississippi()
    oracle_small()
    julia_notebook()
