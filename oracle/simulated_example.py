"""TEST INFRASTRUCTURE (oracle/): replay of the reference notebook notebooks/SimulatedExample.ipynb (Julia 1.4.0, committed
with its outputs).  Only tests/ and tests/golden/make_simulated_example.py import this.

The notebook seeds the global RNG (`Random.seed!(1)`, cell 5), simulates 300 units, fits the joint model with covariates,
and prints (cells 10, 11, 15, 17, 18) the optimum of the marginal likelihood, four fitted hyper-parameters, two LATE
estimates and two p-values.  With oracle/julia_rng.py the data are regenerated bit for bit, and with them the oracle
(oracle/geordd_oracle.py) reproduces every one of those printed numbers -- which pins, against the REFERENCE's own outputs,
MultiGPCovars' mll, postmean_beta and residuals_data (a5), GPRealisations (a6), cliff_face (a7),
inverse_variance and pval_invvar_uncalib (a11), the 3-argument pval_invvar_calib (a12), sentinels / projection_points /
proj_estimator (f4), and decides SURVEY App. B-6: `regions_from_dataframe` (src/regiondata.jl:138-146) installs
FullDummyCoding for EVERY term of the linear formula, numeric ones included, so the covariate matrix D has one indicator
column per distinct value of covarA and of covarB: p = 1 + 300 + 300 + 3 = 604 at n = 300, not 6.  (With the 6-column
design the optimum is 116.47, with 604 columns 123.1640 = the notebook's.)
"""
import math

import numpy as np

import julia_rng as J

# printed by the notebook (cell numbers of the raw .ipynb "cells" list)
NOTEBOOK = {
    "minimum": 123.1640,                       # cell 10: "Minimum:   1.231640e+02"  (= -mll at the optimiser's end point)
    "minimizer_head": (-1.20, -5.16e-3, 0.792),   # cell 10: logNoise, log ell, log sigma_f
    "grad_inf_norm": 5.08e-4,                  # cell 10: |g(x)| at the end point: the optimiser had NOT fully converged
    "sigma_y": 0.3008, "sigma_f": 2.2089, "ell": 0.9949, "sigma_beta": 0.0981,   # cell 11
    "tau_inv": (0.314, 0.090), "tau_inv_cdf0_pct": 0.025,                         # cell 15
    "tau_proj": (0.398, 0.116), "tau_proj_cdf0_pct": 0.031,
    "pval_invvar_uncalib": 0.0005060443694744312,                                 # cell 17
    "pval_invvar_calib": 0.00031629020838950655,                                  # cell 18
    "initial_point": (1.0, math.log(0.5), 0.0, math.log(20.0), 0.0),              # cell 9: logNoise, SEIso(ll, lsigma), Const, LinIso
}
N_UNITS = 300
TREATMENT_RADIUS = 0.8


def notebook_data():
    """Cell 5, call for call: Random.seed!(1); X = rand(2, n); covarA = randn(n); covarB = randn(n);
    categ = rand(["A","B","C"], n); Y = m + f_X + 0.3*randn(n) + tau*treat + 0.1*covarA + 0.2*(categ == "B")."""
    rng = J.MersenneTwister(1)
    n = N_UNITS
    X = rng.rand_array(2 * n).reshape((2, n), order="F")
    covarA = J.randn_array(rng, n, bulk=False)
    covarB = J.randn_array(rng, n, bulk=False)
    categ = np.array([rng.rand_range(3) for _ in range(n)])            # 1, 2, 3 = "A", "B", "C"
    noise = J.randn_array(rng, n, bulk=False)
    f_X = np.sin(X[0] * 2.0) + 3.0 * X[1] ** 2.0
    treat = np.sqrt(X[0] ** 2 + X[1] ** 2) < TREATMENT_RADIUS
    Y = -0.8 + f_X + 0.3 * noise + 0.3 * treat + covarA * 0.1 + (categ == 2) * 0.2
    return {"X": np.asfortranarray(X), "Y": Y, "treat": treat, "covarA": covarA, "covarB": covarB, "categ": categ}


def full_dummy_design(covarA, covarB, categ) -> np.ndarray:
    """The p x n matrix D that regions_from_dataframe builds for `outcome ~ covarA + covarB + categ`
    (src/regiondata.jl:138-146): implicit intercept, then FullDummyCoding of every term -- for a numeric column that is one
    indicator per distinct value, levels sorted (StatsModels 0.6: a contrasts hint makes any column a CategoricalTerm)."""
    def dummy(v):
        return (np.asarray(v)[None, :] == np.sort(np.unique(v))[:, None]).astype(np.float64)
    n = len(covarA)
    return np.asfortranarray(np.vstack([np.ones((1, n)), dummy(covarA), dummy(covarB), dummy(categ)]))


def border_vertices() -> np.ndarray:
    """Cell 6: the LineString is built from the first n = 300 of the 1000 arc vertices (`for i in 1:n`)."""
    th = np.linspace(0.0, math.pi / 2.0, 1000)[:N_UNITS]
    return np.stack([TREATMENT_RADIUS * np.cos(th), TREATMENT_RADIUS * np.sin(th)], axis=1)


def grouped(d):
    """Units ordered control (region == false) first, then treated, as `levels(group_col)` orders the groups
    (src/regiondata.jl:148-165); returns X, Y, D, offsets."""
    order = np.concatenate([np.where(~d["treat"])[0], np.where(d["treat"])[0]])
    D = full_dummy_design(d["covarA"], d["covarB"], d["categ"])
    nC = int((~d["treat"]).sum())
    return np.asfortranarray(d["X"][:, order]), d["Y"][order], np.asfortranarray(D[:, order]), [0, nC, len(order)]
