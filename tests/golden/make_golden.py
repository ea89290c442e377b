"""Generates the committed fixtures under tests/golden/.  Run in the build container (it reads the
reference's data files, which do not exist on the GPU box):

    python tests/golden/make_golden.py

mississippi.npz  -- the reference's own fixture (notebooks/Mississippi_data/{X_MS,X_LA,border}.csv) with
                    the 200 sentinels of notebooks/Mississippi_sharp_null_sims.ipynb cell 7 placed by the
                    oracle's restatement of GeoRDD.sentinels (src/geometry.jl:22-30).  Coordinates only
                    (DATA, not source code); the notebook draws Y from the null prior.
oracle_small.npz -- outputs of the CPU oracle on a seeded synthetic two-region problem (SURVEY §8d shapes,
                    scaled down), used to detect drift of the oracle itself and as a second target for
                    the GPU parity tests.  NOT reference outputs: parity is unpinned by the reference.
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


if __name__ == "__main__":
    mississippi()
    oracle_small()
