"""CPU oracle for the GeoRDD.jl hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``geordd.jl_b200``) never imports it and has no CPU fallback.

PARITY STATUS.  The reference (maximerischard/GeoRDD.jl @ /root/reference) ships no golden vectors, Julia is not
installed, and the arithmetic lives in un-vendored third-party packages (GaussianProcesses.jl 0.11.2, PDMats
0.9.12, Distributions 0.22.6, Distances 0.8.2 -- Manifest.toml:222-226,376-380,140-144,130-134).  What pins this
oracle to the reference's own outputs:

  PINNED by the printed outputs of notebooks/Mississippi_sharp_null_sims.ipynb (Julia 1.4.0, cells 11-16), replayed
  with the restated Julia MersenneTwister / randn of oracle/julia_rng.py (tests/test_oracle.py::
  test_reference_notebook_outputs_pin_the_oracle, fixture tests/golden/julia_notebook.npz):
    * make_null, prior_rand, update_alpha!, mll of the treated / control / pooled GPs, nsim_logP, boot_mlltest:
      the printed p-value 0.0354 = 354 exceedances of 10 000, reproduced EXACTLY;
    * nsim_chi (cliff-face mean at the sentinels, the make_posdef!-factored sentinel system): the two printed
      boot_chi2test values 0.0754 and 0.0759 are reproduced for one common observed statistic within 0.11 % of this
      oracle's (that statistic is an LU solve at condition number 1.8e14 and is not reproducible to more digits);
    * pval_invvar_uncalib / pval_invvar_calib: within the noise of the same ill-conditioned solve (1 % / 16 %);
    * the published size / power table (cells 19-28): all ten entries reproduced by an exact replay
      (scripts/replay_mississippi_table.py -> tests/golden/julia_table_replay.json): nsim_invvar_pval_uncalib, nsim_logP,
      nsim_chi, nsim_power with the corrected analytic calibration;
  PINNED (round 2) by the printed outputs of notebooks/SimulatedExample.ipynb (Julia 1.4.0, cells 5-18), whose data are
  replayed bit for bit by oracle/simulated_example.py (tests/test_oracle.py::test_simulated_example_notebook_pins_the_oracle,
  fixture tests/golden/simulated_example.npz):
    * MultiGPCovars mll and its stationary point: the printed optimum 123.1640 to all seven digits and the four printed
      hyper-parameters to the optimiser's own tolerance (1e-4) -- with the 604-column full-dummy covariate design
      (SURVEY App. B-6 settled: p = 2n + 4, not 6);
    * postmean_beta, residuals_data, GPRealisations, sentinels, cliff_face, inverse_variance (N(0.314, 0.090)),
      projection_points / proj_estimator (N(0.398, 0.116)), pval_invvar_uncalib (7e-5 relative) and the 3-argument
      pval_invvar_calib (9e-5 relative; the remaining difference is the notebook optimiser's |g| = 5e-4 stopping point);
  PARITY UNPINNED by the reference for what neither notebook prints: weight_at_units / weights_*, points_in_region /
  infinite_proj_*, the placebo drivers, the Matern kernels and the gradient entries themselves (cross-checked
  independently of the reference: a 50-digit mpmath variant, identities such as y'K^-1 y = z'z for y = Lz, analytic
  gradient vs finite differences as in test/test_gprealisations.jl:12-40, the statistical acceptance tests of
  test/test_simulated.jl:53-71).

This file restates, in NumPy/SciPy float64 calling the same BLAS/LAPACK routines in the same order, the algorithm of

    src/gp_utils.jl, src/hypotest.jl, src/cliff_face.jl:3-9, src/hypotest_chi2.jl:4-59,
    src/hypotest_invvar.jl:1-137, src/hypotest_mll.jl:1-45, src/power.jl,
    src/average_treatment_effect.jl, src/GPrealisationsCovars.jl:101-193,284-344,
    src/GPrealisations.jl:21-109, src/region_gp_fit.jl, src/weight_at_units.jl, src/border_projection.jl,
    src/placebo_geometry.jl:1-32,60-136 and the placebo_* drivers (src/hypotest_chi2.jl:62-76,
    src/hypotest_invvar.jl:162-175, src/hypotest_mll.jl:47-60)

plus the published semantics of the third-party calls those files make.

Conventions (Julia): all matrices column-major in spirit; coordinates are dim x n
(unit i = column i); D is p x n; sentinels are 2 x nb; Cholesky factors are UPPER
(U'U = K) as PDMats stores them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
from scipy.linalg import lapack, blas, lu_factor, lu_solve
from scipy.special import erfc

SE_ISO, MAT12, MAT32, MAT52 = 0, 1, 2, 3
EPS = np.finfo(np.float64).eps
INVSQRT2 = 0.7071067811865476  # StatsFuns.invsqrt2
LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# Kernel specification.  The ABI carries exactly this struct.
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class KernelSpec:
    """Isotropic spatial kernel (+ optional Const) and constant mean.

    kind: SE_ISO / MAT12 / MAT32 / MAT52;  ell: length scale (a length, not squared);
    sigma2: signal variance; const_sigma2: variance of an added Const kernel (0 = none);
    mean_const: MeanConst value (0.0 = MeanZero).
    GaussianProcesses parameterisation: SEIso(ll, ls): l2 = exp(2 ll), s2 = exp(2 ls);
    MatXXIso(ll, ls): l = exp(ll), s2 = exp(2 ls); Const(ls): s2 = exp(2 ls).
    """
    kind: int
    ell: float
    sigma2: float
    const_sigma2: float = 0.0
    mean_const: float = 0.0

    def with_mean(self, mean_const: float) -> "KernelSpec":
        return KernelSpec(self.kind, self.ell, self.sigma2, self.const_sigma2, mean_const)


def _sqdist(X1: np.ndarray, X2: np.ndarray, method: str = "direct") -> np.ndarray:
    """Squared Euclidean distances between columns of X1 (dim x n1) and X2 (dim x n2).

    'direct' = sum of squared differences (what the GPU path does; more accurate).
    'gram'   = Distances.jl 0.8.2 pairwise: |a|^2 + |b|^2 - 2a'b clamped at 0
                -- what the Julia package computes.
    """
    if method == "gram":
        sa = np.sum(X1 * X1, axis=0)
        sb = np.sum(X2 * X2, axis=0)
        r = X1.T @ X2
        return np.maximum(sa[:, None] + sb[None, :] - 2.0 * r, 0.0)
    out = np.zeros((X1.shape[1], X2.shape[1]))
    for d in range(X1.shape[0]):
        diff = X1[d][:, None] - X2[d][None, :]
        out += diff * diff
    return out


def kernel_of_r2(spec: KernelSpec, r2: np.ndarray) -> np.ndarray:
    """k(r) for the four isotropic kernels + Const (App. A.1)."""
    if spec.kind == SE_ISO:
        k = spec.sigma2 * np.exp(-0.5 * r2 / (spec.ell * spec.ell))   # se.s2*exp(-0.5*r/se.l2)
    else:
        r = np.sqrt(r2)
        if spec.kind == MAT12:
            k = spec.sigma2 * np.exp(-r / spec.ell)
        elif spec.kind == MAT32:
            s = math.sqrt(3.0) * r / spec.ell
            k = spec.sigma2 * (1.0 + s) * np.exp(-s)
        elif spec.kind == MAT52:
            s = math.sqrt(5.0) * r / spec.ell
            k = spec.sigma2 * (1.0 + s + s * s / 3.0) * np.exp(-s)
        else:
            raise ValueError("unknown kernel kind")
    return k + spec.const_sigma2


def cov_cross(spec: KernelSpec, X1: np.ndarray, X2: np.ndarray, method: str = "direct") -> np.ndarray:
    """cov(kernel, X1, X2): n1 x n2 (call sites src/hypotest_chi2.jl:38-39, src/power.jl:60-62)."""
    return kernel_of_r2(spec, _sqdist(X1, X2, method))


def cov_sym(spec: KernelSpec, X: np.ndarray, diag_add: float = 0.0, method: str = "direct") -> np.ndarray:
    """cov(kernel, X, X) + diag_add*I, exactly symmetric."""
    K = cov_cross(spec, X, X, method)
    K = np.tril(K) + np.tril(K, -1).T
    if diag_add != 0.0:
        K[np.diag_indices_from(K)] += diag_add
    return K


# ----------------------------------------------------------------------------
# make_posdef! and PDMat operations
# ----------------------------------------------------------------------------
class PosDefException(Exception):
    def __init__(self, info: int):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


def _potrf_upper(M: np.ndarray):
    """LAPACK dpotrf('U') on a copy; returns (U, info)."""
    U, info = lapack.dpotrf(M, lower=0, clean=1, overwrite_a=0)
    return U, int(info)


def make_posdef(M: np.ndarray):
    """GaussianProcesses.make_posdef!(M, buf) (App. A.5).

    Up to 10 attempts {factor; on failure eps = 1e-6*tr(M)/n and M[i,i] += eps (M is
    mutated so eps compounds on the jittered trace)}, then one final un-caught attempt.
    Returns (M_jittered, U, rounds) with rounds = number of jitter additions made.
    """
    M = np.array(M, dtype=np.float64, copy=True, order="F")
    n = M.shape[0]
    rounds = 0
    for _ in range(10):
        U, info = _potrf_upper(M)
        if info == 0:
            return M, U, rounds
        eps = 1e-6 * np.trace(M) / n
        M[np.diag_indices(n)] += eps
        rounds += 1
    U, info = _potrf_upper(M)
    if info != 0:
        raise PosDefException(info)
    return M, U, rounds


def pd_solve(U: np.ndarray, b: np.ndarray) -> np.ndarray:
    """PDMat \\ b  = U \\ (U' \\ b)  (two triangular solves = dpotrs)."""
    x, info = lapack.dpotrs(U, b, lower=0)
    assert info == 0
    return x


def pd_logdet(U: np.ndarray) -> float:
    return 2.0 * float(np.sum(np.log(np.diag(U))))


def whiten(U: np.ndarray, X: np.ndarray) -> np.ndarray:
    """PDMats.whiten!(pd, X) = U' \\ X."""
    V, info = lapack.dtrtrs(U, X, lower=0, trans=1)
    assert info == 0
    return V


def unwhiten_vec(U: np.ndarray, z: np.ndarray) -> np.ndarray:
    """PDMats.unwhiten!(pd, z) = U' z  (dtrmv, trans)."""
    return blas.dtrmv(U, z, lower=0, trans=1)


def generic_solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Julia's ``A \\ b`` on a square dense Matrix: lu(A) \\ b with partial pivoting
    (App. A.7) -- NOT Cholesky even when A is symmetric."""
    lu, piv = lu_factor(A)
    return lu_solve((lu, piv), b)


# ----------------------------------------------------------------------------
# GPE (App. A.3) and src/gp_utils.jl
# ----------------------------------------------------------------------------
@dataclass
class GP:
    """Mirror of the GaussianProcesses.GPE fields GeoRDD touches."""
    x: np.ndarray          # dim x n
    y: np.ndarray          # n
    spec: KernelSpec
    log_noise: float
    K: np.ndarray = None   # cK.mat   (jittered)
    U: np.ndarray = None   # cK.chol.U
    alpha: np.ndarray = None
    mll: float = float("nan")
    jitter_rounds: int = 0

    @property
    def nobs(self) -> int:
        return self.x.shape[1]

    def mean_at(self, X: np.ndarray) -> np.ndarray:
        return np.full(X.shape[1], self.spec.mean_const)


def gp_mll(gp: GP) -> float:
    """src/gp_utils.jl:19-22."""
    mu = gp.mean_at(gp.x)
    return -float(np.dot(gp.y - mu, gp.alpha)) / 2.0 - pd_logdet(gp.U) / 2.0 - gp.nobs * LOG2PI / 2.0


def update_alpha(gp: GP) -> None:
    """src/gp_utils.jl:15-18."""
    gp.alpha = pd_solve(gp.U, gp.y - gp.mean_at(gp.x))


def gp_fit(spec: KernelSpec, x: np.ndarray, y: np.ndarray, log_noise: float, method: str = "direct") -> GP:
    """GPE(x, y, mean, kernel, logNoise) (src/region_gp_fit.jl:12,28; src/hypotest.jl:4).

    K = k(X,X) + (exp(2 logNoise) + eps)*I ; make_posdef! ; alpha ; mll  (App. A.3).
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    noise = math.exp(2.0 * log_noise) + EPS
    K = cov_sym(spec, x, noise, method)
    K, U, rounds = make_posdef(K)
    gp = GP(x=x, y=y.copy(), spec=spec, log_noise=log_noise, K=K, U=U, jitter_rounds=rounds)
    update_alpha(gp)
    gp.mll = gp_mll(gp)
    return gp


def modifiable(gp: GP) -> GP:
    """src/gp_utils.jl:10-14: shares x and cK, copies y (re-factoring gives the same values)."""
    g = GP(x=gp.x, y=gp.y.copy(), spec=gp.spec, log_noise=gp.log_noise, K=gp.K, U=gp.U,
           jitter_rounds=gp.jitter_rounds)
    update_alpha(g)
    g.mll = gp_mll(g)
    return g


def predict_mu(gp: GP, Xb: np.ndarray, cK: np.ndarray) -> np.ndarray:
    """src/gp_utils.jl:1-5."""
    return gp.mean_at(Xb) + cK @ gp.alpha


def predict_full(gp: GP, Xb: np.ndarray, method: str = "direct"):
    """predict_f(gp, Xb; full_cov=true) (App. A.6): latent f, no noise term."""
    Kxb = cov_cross(gp.spec, gp.x, Xb, method)          # m x nb
    Kbb = cov_cross(gp.spec, Xb, Xb, method)
    mu = gp.mean_at(Xb) + Kxb.T @ gp.alpha
    V = whiten(gp.U, Kxb)
    # syrk!('U','T',-1,V,1,Kbb) then copytri!: exactly symmetric
    S = Kbb - V.T @ V
    S = np.triu(S) + np.triu(S, 1).T
    return mu, S


# ----------------------------------------------------------------------------
# src/cliff_face.jl:3-9, src/hypotest.jl
# ----------------------------------------------------------------------------
def cliff_face(gpT: GP, gpC: GP, Xb: np.ndarray, method: str = "direct"):
    muT, ST = predict_full(gpT, Xb, method)
    muC, SC = predict_full(gpC, Xb, method)
    return muT - muC, ST + SC


def make_null(gpT: GP, gpC: GP, method: str = "direct") -> GP:
    """src/hypotest.jl:1-12: pooled GP with T's kernel, mean and noise."""
    yN = np.concatenate([gpT.y, gpC.y])
    xN = np.concatenate([gpT.x, gpC.x], axis=1)
    return gp_fit(gpT.spec, xN, yN, gpT.log_noise, method)


def prior_rand(gp: GP, z: np.ndarray) -> np.ndarray:
    """src/hypotest.jl:13-18: rand(MvNormal(mu, cK)) = mu + U' z, one randn(n) per draw."""
    return gp.mean_at(gp.x) + unwhiten_vec(gp.U, z)


# ----------------------------------------------------------------------------
# src/average_treatment_effect.jl
# ----------------------------------------------------------------------------
@dataclass
class Normal:
    mu: float
    sigma: float

    def cdf(self, x: float) -> float:      # App. A.9
        z = (x - self.mu) / self.sigma
        return float(erfc(-z * INVSQRT2)) / 2.0

    def ccdf(self, x: float) -> float:
        z = (x - self.mu) / self.sigma
        return float(erfc(z * INVSQRT2)) / 2.0

    def quantile(self, p: float) -> float:
        from scipy.special import erfcinv
        return self.mu - self.sigma * math.sqrt(2.0) * float(erfcinv(2.0 * p))


def inverse_variance(mu: np.ndarray, Sigma: np.ndarray = None, nugget: float = 1e-10, U: np.ndarray = None) -> Normal:
    """src/average_treatment_effect.jl:3-11.  Sigma a Matrix => + nugget and generic ``\\``
    (LU); a PDMat (pass its factor as U) => no nugget and Cholesky solves."""
    n = mu.shape[0]
    ones = np.ones(n)
    if U is not None:
        denom = float(np.sum(pd_solve(U, ones)))
        num = float(np.sum(pd_solve(U, mu)))
    else:
        S = Sigma + nugget * np.eye(n)
        denom = float(np.sum(generic_solve(S, ones)))
        num = float(np.sum(generic_solve(S, mu)))
    return Normal(num / denom, math.sqrt(1.0 / denom))


def unweighted_mean(mu: np.ndarray, Sigma: np.ndarray) -> Normal:
    """src/average_treatment_effect.jl:13-22."""
    n = mu.shape[0]
    return Normal(float(np.mean(mu)), math.sqrt(float(np.sum(Sigma)) / n ** 2))


def weighted_mean(mu: np.ndarray, Sigma: np.ndarray, w: np.ndarray) -> Normal:
    """src/average_treatment_effect.jl:24-29."""
    sw = float(np.sum(w))
    return Normal(float(np.sum(mu * w)) / sw, math.sqrt(float(w @ (Sigma @ w)) / sw ** 2))


def weight_at_units(gp: GP, Xb: np.ndarray, w: np.ndarray) -> np.ndarray:
    """src/weight_at_units.jl:12-18."""
    KXb = cov_cross(gp.spec, gp.x, Xb)
    return pd_solve(gp.U, KXb) @ w / float(np.sum(w))


# ----------------------------------------------------------------------------
# Null-simulation setup shared by chi2 / invvar / power
# ----------------------------------------------------------------------------
@dataclass
class NullSetup:
    gpT: GP
    gpC: GP
    gpN: GP
    Xb: np.ndarray
    cK_T: np.ndarray        # nb x mT
    cK_C: np.ndarray        # nb x mC
    Sigma_cliff: np.ndarray  # raw cliff-face covariance
    mu_cliff: np.ndarray
    Sraw: np.ndarray        # after (nugget +) make_posdef!
    US: np.ndarray          # its upper factor
    jitter_rounds: int
    treat: np.ndarray


def null_setup(gpT: GP, gpC: GP, Xb: np.ndarray, nugget: float = 0.0, gpN: Optional[GP] = None) -> NullSetup:
    """The draw-independent part of nsim_chi (src/hypotest_chi2.jl:33-48; nugget = 0) and
    nsim_invvar_pval_uncalib (src/hypotest_invvar.jl:48-59; nugget = 1e-10 before make_posdef!)."""
    gpT_mod, gpC_mod = modifiable(gpT), modifiable(gpC)
    mu, S = cliff_face(gpT, gpC, Xb)
    cK_T = cov_cross(gpT.spec, Xb, gpT.x)
    cK_C = cov_cross(gpC.spec, Xb, gpC.x)
    if gpN is None:
        gpN = make_null(gpT, gpC)
    treat = np.zeros(gpN.nobs, dtype=bool)
    treat[: gpT.nobs] = True
    Sin = S + nugget * np.eye(S.shape[0]) if nugget != 0.0 else S.copy()
    Sraw, US, rounds = make_posdef(Sin)
    return NullSetup(gpT_mod, gpC_mod, gpN, Xb, cK_T, cK_C, S, mu, Sraw, US, rounds, treat)


def _sim_update(ns: NullSetup, z: np.ndarray, tau: float = 0.0) -> np.ndarray:
    Ysim = prior_rand(ns.gpN, z)
    if tau != 0.0:
        Ysim[ns.treat] += tau
    ns.gpT.y = Ysim[ns.treat]
    ns.gpC.y = Ysim[~ns.treat]
    update_alpha(ns.gpT)
    update_alpha(ns.gpC)
    return Ysim


# ----------------------------------------------------------------------------
# src/hypotest_chi2.jl
# ----------------------------------------------------------------------------
def chistat_obs(gpT: GP, gpC: GP, Xb: np.ndarray) -> float:
    """src/hypotest_chi2.jl:4-10: generic ``\\`` on the raw Sigma => LU."""
    mu, S = cliff_face(gpT, gpC, Xb)
    return float(np.dot(mu, generic_solve(S, mu)))


def chistat_pre(ns: NullSetup) -> float:
    """src/hypotest_chi2.jl:11-17."""
    mu = predict_mu(ns.gpT, ns.Xb, ns.cK_T) - predict_mu(ns.gpC, ns.Xb, ns.cK_C)
    return float(np.dot(mu, pd_solve(ns.US, mu)))


def nsim_chi(gpT: GP, gpC: GP, gpN: GP, Xb: np.ndarray, Z: np.ndarray) -> np.ndarray:
    """src/hypotest_chi2.jl:33-52; Z is n x nsim, column s = the randn(n) of draw s."""
    ns = null_setup(gpT, gpC, Xb, 0.0, gpN)
    out = np.empty(Z.shape[1])
    for s in range(Z.shape[1]):
        _sim_update(ns, Z[:, s])
        out[s] = chistat_pre(ns)
    return out


def boot_chi2test(gpT: GP, gpC: GP, Xb: np.ndarray, Z: np.ndarray):
    """src/hypotest_chi2.jl:54-59.  Returns (p, n_exceed, chi_obs, chi_sims)."""
    gpN = make_null(gpT, gpC)
    chi_obs = chistat_obs(gpT, gpC, Xb)
    sims = nsim_chi(gpT, gpC, gpN, Xb, Z)
    cnt = int(np.sum(sims > chi_obs))
    return cnt / sims.shape[0], cnt, chi_obs, sims


# ----------------------------------------------------------------------------
# src/hypotest_invvar.jl
# ----------------------------------------------------------------------------
def _pval_two_sided_at_zero(nrm: Normal) -> float:
    return 2.0 * min(nrm.cdf(0.0), nrm.ccdf(0.0))


def pval_invvar_uncalib_obs(gpT: GP, gpC: GP, Xb: np.ndarray, nugget: float = 1e-10) -> float:
    """src/hypotest_invvar.jl:1-6."""
    mu, S = cliff_face(gpT, gpC, Xb)
    return _pval_two_sided_at_zero(inverse_variance(mu, S, nugget=nugget))


def pval_invvar_uncalib_pre(ns: NullSetup) -> float:
    """src/hypotest_invvar.jl:11-19."""
    mu = predict_mu(ns.gpT, ns.Xb, ns.cK_T) - predict_mu(ns.gpC, ns.Xb, ns.cK_C)
    return _pval_two_sided_at_zero(inverse_variance(mu, U=ns.US))


def nsim_invvar_pval_uncalib(gpT: GP, gpC: GP, Xb: np.ndarray, Z: np.ndarray, nugget: float = 1e-10) -> np.ndarray:
    """src/hypotest_invvar.jl:48-63."""
    ns = null_setup(gpT, gpC, Xb, nugget)
    out = np.empty(Z.shape[1])
    for s in range(Z.shape[1]):
        _sim_update(ns, Z[:, s])
        out[s] = pval_invvar_uncalib_pre(ns)
    return out


def boot_invvar(gpT: GP, gpC: GP, Xb: np.ndarray, Z: np.ndarray, nugget: float = 1e-10):
    """src/hypotest_invvar.jl:65-69.  Returns (p, n_below, pval_obs, pval_sims)."""
    pobs = pval_invvar_uncalib_obs(gpT, gpC, Xb, nugget)
    sims = nsim_invvar_pval_uncalib(gpT, gpC, Xb, Z, nugget)
    cnt = int(np.sum(sims < pobs))
    return cnt / sims.shape[0], cnt, pobs, sims


def _calib_cov_mu_delta(gpT: GP, gpC: GP, cK_T: np.ndarray, cK_C: np.ndarray, KCT: np.ndarray, as_shipped: bool):
    """cov_mu_delta of src/hypotest_invvar.jl:88-96 (3-arg, correct) / :121-128 (7-arg; the
    statement ends after the first product -- missing '+')."""
    KTb = cK_T.T
    KCb = cK_C.T
    WT_c = pd_solve(gpT.U, KTb)          # KTT \ KTb   (mT x nb)
    WC_c = pd_solve(gpC.U, KCb)
    WT, WC = WT_c.T, WC_c.T
    first = WT @ gpT.K @ WT_c
    if as_shipped:
        return first
    return first + WC @ gpC.K @ WC_c - WC @ KCT @ WT_c - WT @ KCT.T @ WC_c


def pval_invvar_calib(gpT: GP, gpC: GP, Xb: np.ndarray, nugget: float = 1e-10) -> float:
    """src/hypotest_invvar.jl:74-105 (3-argument, correct formula; Sigma_b a Matrix => LU)."""
    mub, Sb = cliff_face(gpT, gpC, Xb)
    n = mub.shape[0]
    Sb = Sb + nugget * np.eye(n)
    cK_C = cov_cross(gpC.spec, Xb, gpC.x)
    KCT = cov_cross(gpC.spec, gpC.x, gpT.x)
    cK_T = cov_cross(gpT.spec, gpT.x, Xb).T
    cmd = _calib_cov_mu_delta(gpT, gpC, cK_T, cK_C, KCT, as_shipped=False)
    lu = lu_factor(Sb)
    cov_mt = float(np.sum(lu_solve(lu, cmd) @ lu_solve(lu, np.ones(n))))
    num = float(np.sum(lu_solve(lu, mub)))
    return 2.0 * Normal(0.0, math.sqrt(cov_mt)).ccdf(abs(num))


def calib_cov_mu_tau(ns: NullSetup, KCT: np.ndarray, as_shipped: bool = False) -> float:
    """The draw-independent part of the 7-argument pval_invvar_calib (src/hypotest_invvar.jl:111-131):
    cov_mu_tau = sum((Sigma \\ cov_mu_delta) * (Sigma \\ ones)).  The reference recomputes it for every
    draw; it does not depend on y, so computing it once gives bit-identical values."""
    n = ns.Sraw.shape[0]
    cmd = _calib_cov_mu_delta(ns.gpT, ns.gpC, ns.cK_T, ns.cK_C, KCT, as_shipped)
    return float(np.sum(pd_solve(ns.US, cmd) @ pd_solve(ns.US, np.ones(n))))


def pval_invvar_calib_pre(ns: NullSetup, KCT: np.ndarray, as_shipped: bool = False,
                          cov_mt: Optional[float] = None) -> float:
    """src/hypotest_invvar.jl:106-137 (7-argument variant used by sim_power!, src/power.jl:41).
    as_shipped=True reproduces the missing-'+' defect.  cov_mt: pass calib_cov_mu_tau(...) to hoist."""
    mub = predict_mu(ns.gpT, ns.Xb, ns.cK_T) - predict_mu(ns.gpC, ns.Xb, ns.cK_C)
    if cov_mt is None:
        cov_mt = calib_cov_mu_tau(ns, KCT, as_shipped)
    num = float(np.sum(pd_solve(ns.US, mub)))
    return 2.0 * Normal(0.0, math.sqrt(cov_mt)).ccdf(abs(num))


def nsim_invvar_calib(gpT: GP, gpC: GP, Xb: np.ndarray, Z: np.ndarray) -> np.ndarray:
    """src/hypotest_invvar.jl:138-160: per draw prior_rand -> update_alpha! x2 -> pval_invvar_calib(gpT, gpC, Xb)
    (3-argument variant: cliff face recomputed, nugget 1e-10, LU)."""
    gT, gC = modifiable(gpT), modifiable(gpC)
    gN = make_null(gT, gC)
    mT = gpT.nobs
    out = np.empty(Z.shape[1])
    for s in range(Z.shape[1]):
        Ysim = prior_rand(gN, Z[:, s])
        gT.y, gC.y = Ysim[:mT], Ysim[mT:]
        update_alpha(gT); update_alpha(gC)
        out[s] = pval_invvar_calib(gT, gC, Xb)
    return out


# ----------------------------------------------------------------------------
# src/hypotest_mll.jl
# ----------------------------------------------------------------------------
def nsim_logP(gpT: GP, gpC: GP, gpN: GP, Z: np.ndarray):
    """src/hypotest_mll.jl:1-30.  Returns (mll_null[S], mll_alt[S]); mutates gpN like the reference."""
    gpT_mod, gpC_mod = modifiable(gpT), modifiable(gpC)
    mT = gpT.nobs
    S = Z.shape[1]
    mnull, malt = np.empty(S), np.empty(S)
    for s in range(S):
        Ysim = prior_rand(gpN, Z[:, s])
        gpT_mod.y = Ysim[:mT]
        gpC_mod.y = Ysim[mT:]
        gpN.y = Ysim
        update_alpha(gpT_mod); update_alpha(gpC_mod); update_alpha(gpN)
        gpT_mod.mll = gp_mll(gpT_mod); gpC_mod.mll = gp_mll(gpC_mod); gpN.mll = gp_mll(gpN)
        malt[s] = gpT_mod.mll + gpC_mod.mll
        mnull[s] = gpN.mll
    return mnull, malt


def boot_mlltest(gpT: GP, gpC: GP, Z: np.ndarray):
    """src/hypotest_mll.jl:32-45.  Returns (p, n_exceed, dmll_obs, dmll_sims)."""
    mll_alt = gpT.mll + gpC.mll
    gpN = make_null(gpT, gpC)
    mll_null = gpN.mll
    mnull, malt = nsim_logP(gpT, gpC, gpN, Z)
    d_obs = mll_alt - mll_null
    d_sim = malt - mnull
    cnt = int(np.sum(d_sim > d_obs))
    return cnt / d_sim.shape[0], cnt, d_obs, d_sim


# ----------------------------------------------------------------------------
# src/power.jl
# ----------------------------------------------------------------------------
def nsim_power(gpT: GP, gpC: GP, tau: float, Xb: np.ndarray, chi_null: np.ndarray, mll_null: np.ndarray,
               pval_invvar_null: np.ndarray, Z: np.ndarray, as_shipped_calib: bool = False) -> np.ndarray:
    """src/power.jl:1-73.  Returns S x 5: (pval_mll, pval_chi2, pval_invvar_obs,
    invvar_bootcalib, invvar_calib)."""
    ns = null_setup(gpT, gpC, Xb, 0.0)   # make_posdef!(copy(Sigma_cliff)), src/power.jl:57-59
    KCT = cov_cross(gpC.spec, gpC.x, gpT.x)
    cov_mt = calib_cov_mu_tau(ns, KCT, as_shipped_calib)   # hoisted: draw-independent
    S = Z.shape[1]
    out = np.empty((S, 5))
    for s in range(S):
        Ysim = _sim_update(ns, Z[:, s], tau)
        ns.gpN.y = Ysim
        update_alpha(ns.gpN)
        ns.gpT.mll = gp_mll(ns.gpT); ns.gpC.mll = gp_mll(ns.gpC); ns.gpN.mll = gp_mll(ns.gpN)
        dmll = ns.gpT.mll + ns.gpC.mll - ns.gpN.mll
        out[s, 0] = np.mean(mll_null > dmll)
        chi = chistat_pre(ns)
        out[s, 1] = np.mean(chi_null > chi)
        pobs = pval_invvar_uncalib_pre(ns)
        out[s, 2] = pobs
        out[s, 3] = np.mean(pval_invvar_null < pobs)
        out[s, 4] = pval_invvar_calib_pre(ns, KCT, as_shipped_calib, cov_mt)
    return out


# ----------------------------------------------------------------------------
# src/GPrealisationsCovars.jl : joint fit with covariates
# ----------------------------------------------------------------------------
@dataclass
class MultiGP:
    """Numerical content of MultiGPCovars (src/GPrealisationsCovars.jl:1-50)."""
    D: np.ndarray                 # p x n
    y: np.ndarray                 # n
    X: np.ndarray                 # dim x n, groups contiguous
    group_off: Sequence[int]      # ngroups+1 offsets
    spec: KernelSpec
    lin_ell2: float               # LinIso l2 = exp(2 ll);  sigma_beta = 1/sqrt(l2)
    log_noise: float
    K: np.ndarray = None
    U: np.ndarray = None
    alpha: np.ndarray = None
    mll: float = float("nan")
    dmll: np.ndarray = None

    @property
    def nobs(self):
        return self.D.shape[1]


def _kernel_param_grads(spec: KernelSpec, r2: np.ndarray):
    """d k / d log-params for the SumKernel(spatial, Const) in get_params order
    [log ell, log sigma, (log sigma_const)]  (App. A.1)."""
    grads = []
    if spec.kind == SE_ISO:
        k = spec.sigma2 * np.exp(-0.5 * r2 / (spec.ell * spec.ell))
        grads.append(k * r2 / (spec.ell * spec.ell))
    else:
        r = np.sqrt(r2)
        if spec.kind == MAT12:
            k = spec.sigma2 * np.exp(-r / spec.ell)
            grads.append(k * r / spec.ell)
        elif spec.kind == MAT32:
            s = math.sqrt(3.0) * r / spec.ell
            k = spec.sigma2 * (1.0 + s) * np.exp(-s)
            grads.append(spec.sigma2 * s * s * np.exp(-s))
        else:
            s = math.sqrt(5.0) * r / spec.ell
            k = spec.sigma2 * (1.0 + s + s * s / 3.0) * np.exp(-s)
            grads.append(spec.sigma2 * s * s * (1.0 + s) * np.exp(-s) / 3.0)
    grads.append(2.0 * k)
    if spec.const_sigma2 != 0.0:
        grads.append(np.full_like(r2, 2.0 * spec.const_sigma2))
    return grads


def multigp_assemble(mg: MultiGP) -> np.ndarray:
    """cK of update_mll!/initialise_mll! (src/GPrealisationsCovars.jl:110-142):
    D'D/l2 + blockdiag(K_g) + max(exp(2 logNoise), 1e-8) I."""
    n = mg.nobs
    K = (mg.D.T @ mg.D) / mg.lin_ell2
    for g in range(len(mg.group_off) - 1):
        a, b = mg.group_off[g], mg.group_off[g + 1]
        K[a:b, a:b] += cov_sym(mg.spec, mg.X[:, a:b])
    K[np.diag_indices(n)] += max(math.exp(2.0 * mg.log_noise), 1e-8)
    return K


def multigp_update_mll(mg: MultiGP) -> float:
    """src/GPrealisationsCovars.jl:110-125."""
    K = multigp_assemble(mg)
    mg.K, mg.U, _ = make_posdef(K)
    mu = np.full(mg.nobs, mg.spec.mean_const)
    mg.alpha = pd_solve(mg.U, mg.y - mu)
    mg.mll = -float(np.dot(mg.y - mu, mg.alpha)) / 2.0 - pd_logdet(mg.U) / 2.0 - mg.nobs * LOG2PI / 2.0
    return mg.mll


def multigp_update_mll_and_dmll(mg: MultiGP, noise=True, domean=True, kern=True, beta=True) -> np.ndarray:
    """src/GPrealisationsCovars.jl:145-193 + get_aa_invcKI!/dmll_kern! (App. A.8).
    Order: [logNoise][mean params][kernel params: log ell, log sigma,(log sigma_c)][LinIso ll].
    MeanZero has no parameters; MeanConst (mean_const != 0) has one."""
    multigp_update_mll(mg)
    n = mg.nobs
    B = -pd_solve(mg.U, np.eye(n))
    B += np.outer(mg.alpha, mg.alpha)
    out = []
    if noise:
        out.append(math.exp(2.0 * mg.log_noise) * float(np.trace(B)))
    if domean and mg.spec.mean_const != 0.0:
        out.append(float(np.sum(mg.alpha)))
    if kern:
        nk = 3 if mg.spec.const_sigma2 != 0.0 else 2
        acc = np.zeros(nk)
        for g in range(len(mg.group_off) - 1):
            a, b = mg.group_off[g], mg.group_off[g + 1]
            r2 = _sqdist(mg.X[:, a:b], mg.X[:, a:b])
            Bg = B[a:b, a:b]
            for p, dK in enumerate(_kernel_param_grads(mg.spec, r2)):
                acc[p] += 0.5 * float(np.sum(dK * Bg))
        out.extend(acc.tolist())
    if beta:
        G = mg.D.T @ mg.D
        out.append(0.5 * float(np.sum((-2.0 * G / mg.lin_ell2) * B)))
    mg.dmll = np.array(out)
    return mg.dmll


def multigp_postmean_beta(mg: MultiGP) -> np.ndarray:
    """src/GPrealisationsCovars.jl:284-344: Sigma_Y|beta = blockdiag K_g + noise ;
    precision = X_invA_Xt(Sigma, D) + l2 I ; beta = (precision \\ D)(Sigma \\ (y - m))."""
    n = mg.nobs
    K = np.zeros((n, n))
    for g in range(len(mg.group_off) - 1):
        a, b = mg.group_off[g], mg.group_off[g + 1]
        K[a:b, a:b] += cov_sym(mg.spec, mg.X[:, a:b])
    K[np.diag_indices(n)] += max(math.exp(2.0 * mg.log_noise), 1e-8)
    U, info = _potrf_upper(K)           # PDMats.PDMat(cK): plain cholesky, no jitter
    if info != 0:
        raise PosDefException(info)
    W = whiten(U, np.ascontiguousarray(mg.D.T))      # U' \ D'  (n x p) ;  X_invA_Xt = W'W
    prec = W.T @ W
    prec[np.diag_indices(mg.D.shape[0])] += mg.lin_ell2
    m = np.full(n, mg.spec.mean_const)
    rhs = pd_solve(U, mg.y - m)
    return generic_solve(prec, mg.D) @ rhs


def residuals(y: np.ndarray, D: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """src/regiondata.jl:11: rd.y - rd.D'beta."""
    return y - D.T @ beta


def gp_realisations(spec: KernelSpec, xs: Sequence[np.ndarray], ys: Sequence[np.ndarray], log_noise: float):
    """GPRealisations(rdict, keys, spkern, logNoise, mean) (src/region_gp_fit.jl:23-32 ->
    src/GPrealisations.jl:10-19,101-109): one GPE per region with shared (kernel, logNoise);
    returns (list of fitted GPs, total mll)."""
    gps = [gp_fit(spec, x, y, log_noise) for x, y in zip(xs, ys)]
    return gps, float(sum(g.mll for g in gps))


# ----------------------------------------------------------------------------
# Counter-based RNG shared with the GPU path (device-RNG mode)
# ----------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al. 2011), vectorised over uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32); c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32); c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PH_W0); k1 = np.uint32(k1 + _PH_W1)
    return c0, c1, c2, c3


def philox_normals(seed: int, draw0: int, ndraws: int, n: int) -> np.ndarray:
    """Z (n x ndraws): column s = standard normals of draw index draw0+s.

    Element pair (2j, 2j+1) of draw d comes from Philox counter (j, 0, d_lo, d_hi), key
    (seed_lo, seed_hi): u1 = ((a<<32|b)>>11 + 0.5) 2^-53, u2 likewise from (c,d);
    Box-Muller z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = ... sin(2 pi u2).  Results depend only
    on (seed, draw index, element) => invariant to GPU count and batch size."""
    npair = (n + 1) // 2
    j = np.arange(npair, dtype=np.uint64)[:, None]
    d = (np.uint64(draw0) + np.arange(ndraws, dtype=np.uint64))[None, :]
    a, b, c, e = philox4x32_10((j & _MASK32).astype(np.uint32), np.zeros_like(j, dtype=np.uint32),
                               (d & _MASK32).astype(np.uint32), (d >> np.uint64(32)).astype(np.uint32),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w1 = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)
    w2 = (c.astype(np.uint64) << np.uint64(32)) | e.astype(np.uint64)
    u1 = ((w1 >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = ((w2 >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * u2
    z0 = rad * np.cos(np.pi * ang)
    z1 = rad * np.sin(np.pi * ang)
    Z = np.empty((2 * npair, ndraws))
    Z[0::2] = z0
    Z[1::2] = z1
    return np.asfortranarray(Z[:n])


# ----------------------------------------------------------------------------
# Synthetic two-region data of SURVEY §8(d) (shape-faithful to test/runtests.jl:4-36)
# ----------------------------------------------------------------------------
def synth_two_region(m: int, nb: int, seed: int = 1, tau: float = 0.3, p_categ: int = 3):
    """X ~ U[0,1]^2 (2 x 2m); treated = |x| < r with r the empirical median radius so that
    m_T = m_C = m; sentinels = nb points equally spaced in arclength on the quarter circle
    of radius r; covarA, covarB ~ N(0,1); categ ~ U{1..p_categ}; D = [1; covarA; covarB;
    onehot(categ)]; Y per test/runtests.jl:15-20.  Returns dict of arrays (units ordered
    treated first)."""
    rng = np.random.default_rng(seed)
    n = 2 * m
    X = rng.random((2, n))
    rad = np.sqrt(X[0] ** 2 + X[1] ** 2)
    order = np.argsort(rad, kind="stable")
    X = X[:, order]
    rad = rad[order]
    r = 0.5 * (rad[m - 1] + rad[m])
    treat = np.arange(n) < m
    fX = np.sin(X[0] * 2.0) + 3.0 * X[1] ** 2
    covarA = rng.standard_normal(n)
    covarB = rng.standard_normal(n)
    categ = rng.integers(0, p_categ, n)
    Y = -0.8 + fX + 0.3 * rng.standard_normal(n) + tau * treat + 0.1 * covarA + 0.2 * (categ == 1)
    theta = np.linspace(0.0, np.pi / 2.0, nb)
    Xb = np.vstack([r * np.cos(theta), r * np.sin(theta)])
    D = np.vstack([np.ones(n), covarA, covarB] + [(categ == c).astype(np.float64) for c in range(p_categ)])
    return dict(X=np.asfortranarray(X), Y=Y, D=np.asfortranarray(D), treat=treat, Xb=np.asfortranarray(Xb),
                r=r, XT=np.asfortranarray(X[:, :m]), XC=np.asfortranarray(X[:, m:]), YT=Y[:m].copy(), YC=Y[m:].copy())


def polyline_sentinels(border: np.ndarray, nsent: int) -> np.ndarray:
    """GeoRDD.sentinels (src/geometry.jl:22-30): nsent points equally spaced in arclength
    along a polyline (border: nv x 2), endpoints included.  Returns 2 x nsent."""
    seg = np.sqrt(np.sum(np.diff(border, axis=0) ** 2, axis=1))
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    q = np.linspace(0.0, cum[-1], nsent)
    idx = np.clip(np.searchsorted(cum, q, side="right") - 1, 0, len(seg) - 1)
    t = np.where(seg[idx] > 0, (q - cum[idx]) / np.where(seg[idx] > 0, seg[idx], 1.0), 0.0)
    pts = border[idx] + t[:, None] * (border[idx + 1] - border[idx])
    return np.asfortranarray(pts.T)


# ---------------------------------------------------------------------------------------------
# Projection of units onto the border (src/border_projection.jl:4-31, src/weight_at_units.jl:44-53)
# ---------------------------------------------------------------------------------------------
def _geos_point_to_segment(px, py, ax, ay, bx, by):
    """GEOS 3.8.1 algorithm/Distance.cpp Distance::pointToSegment [recalled; GEOS_jll 3.8.1+0, Manifest.toml:216-220],
    vectorised over points; every operation separately rounded (NumPy never contracts to FMA)."""
    if ax == bx and ay == by:
        return np.sqrt((px - ax) * (px - ax) + (py - ay) * (py - ay))
    ex, ey = bx - ax, by - ay
    len2 = ex * ex + ey * ey
    r = ((px - ax) * ex + (py - ay) * ey) / len2
    dA = np.sqrt((px - ax) * (px - ax) + (py - ay) * (py - ay))
    dB = np.sqrt((px - bx) * (px - bx) + (py - by) * (py - by))
    sn = ((ay - py) * ex - (ax - px) * ey) / len2
    dS = np.abs(sn) * np.sqrt(len2)
    return np.where(r <= 0.0, dA, np.where(r >= 1.0, dB, dS))


def projection_points_all(X: np.ndarray, border: np.ndarray, part_start=()):
    """distance(border, point) and nearestPoints(border, point)[1] for every unit (src/border_projection.jl:8-19).
    X: 2 x n; border: nv x 2 vertices of a LineString, or of a MultiLineString whose lines start at the vertex
    indices in part_start.  GEOS DistanceOp scans the segments of every line in order and keeps the FIRST minimum
    (strict <); the closest point is LineSegment::closestPoint on that segment [recalled].
    Returns (Xproj 2 x n, dist n)."""
    X = np.asarray(X, dtype=np.float64)
    V = np.asarray(border, dtype=np.float64)
    n, nv = X.shape[1], V.shape[0]
    brk = np.zeros(nv, dtype=bool)
    brk[list(part_start)] = True
    px, py = X[0], X[1]
    best = np.full(n, np.inf)
    bseg = np.zeros(n, dtype=np.int64)
    for s in range(nv - 1):
        if brk[s + 1]:
            continue
        d = _geos_point_to_segment(px, py, V[s, 0], V[s, 1], V[s + 1, 0], V[s + 1, 1])
        upd = d < best
        best = np.where(upd, d, best)
        bseg = np.where(upd, s, bseg)
    A, B = V[bseg], V[bseg + 1]
    ex, ey = B[:, 0] - A[:, 0], B[:, 1] - A[:, 1]
    len2 = ex * ex + ey * ey
    with np.errstate(invalid="ignore", divide="ignore"):
        r = ((px - A[:, 0]) * ex + (py - A[:, 1]) * ey) / len2
    r = np.where((px == A[:, 0]) & (py == A[:, 1]), 0.0, np.where((px == B[:, 0]) & (py == B[:, 1]), 1.0, r))
    inside = (r > 0.0) & (r < 1.0)
    qx = A[:, 0] + r * ex
    qy = A[:, 1] + r * ey
    dA = np.sqrt((px - A[:, 0]) ** 2 + (py - A[:, 1]) ** 2)
    dB = np.sqrt((px - B[:, 0]) ** 2 + (py - B[:, 1]) ** 2)
    ex_ = np.where(dA < dB, A[:, 0], B[:, 0])
    ey_ = np.where(dA < dB, A[:, 1], B[:, 1])
    Xp = np.vstack([np.where(inside, qx, ex_), np.where(inside, qy, ey_)])
    return np.asfortranarray(Xp), best


def projection_points(X: np.ndarray, border: np.ndarray, maxdist: float, part_start=()) -> np.ndarray:
    """projection_points(gp, border, maxdist) (src/border_projection.jl:4-22): X_proj[:, distances .<= maxdist]."""
    Xp, d = projection_points_all(X, border, part_start)
    return np.asfortranarray(Xp[:, d <= maxdist])


def proj_estimator(gpT, gpC, border: np.ndarray, maxdist: float, part_start=()):
    """proj_estimator (src/border_projection.jl:24-31): cliff face at the projected units, unweighted mean.
    Returns (mean, variance) of the Normal."""
    Xb = np.hstack([projection_points(gpT.x, border, maxdist, part_start), projection_points(gpC.x, border, maxdist, part_start)])
    mu, Sig = cliff_face(gpT, gpC, np.asfortranarray(Xb))
    return unweighted_mean(mu, Sig)


def points_in_region(P: np.ndarray, rings: np.ndarray, ring_start=()) -> np.ndarray:
    """LibGEOS.within(point, region) for the columns of P (2 x n).  rings: nr x 2 vertices of all (closed) rings of
    the (Multi)Polygon -- shells and holes; ring_start: vertex indices at which further rings start.  GEOS
    RayCrossingCounter restated [recalled: GEOS 3.8 algorithm/RayCrossingCounter.cpp] with the orientation
    determinant in plain FP64; crossing parities combined over all rings (even-odd), boundary points excluded."""
    P = np.asarray(P, dtype=np.float64)
    R = np.asarray(rings, dtype=np.float64)
    nr = R.shape[0]
    brk = np.zeros(nr, dtype=bool)
    brk[list(ring_start)] = True
    qx, qy = P[0], P[1]
    crossings = np.zeros(P.shape[1], dtype=np.int64)
    on_b = np.zeros(P.shape[1], dtype=bool)
    for s in range(nr - 1):
        if brk[s + 1]:
            continue
        p1x, p1y, p2x, p2y = R[s, 0], R[s, 1], R[s + 1, 0], R[s + 1, 1]
        active = ~((p1x < qx) & (p2x < qx))
        at_v = active & (qx == p2x) & (qy == p2y)
        on_b |= at_v
        active &= ~at_v
        horiz = active & (p1y == qy) & (p2y == qy)
        on_b |= horiz & (qx >= min(p1x, p2x)) & (qx <= max(p1x, p2x))
        active &= ~horiz
        cross = active & (((p1y > qy) & (p2y <= qy)) | ((p2y > qy) & (p1y <= qy)))
        det = (p2x - p1x) * (qy - p1y) - (p2y - p1y) * (qx - p1x)
        sign = np.sign(det).astype(np.int64)
        on_b |= cross & (sign == 0)
        if p2y < p1y:
            sign = -sign
        crossings += (cross & (sign == 1)).astype(np.int64)
    return (~on_b) & (crossings % 2 == 1)


def infinite_proj_sentinels(border: np.ndarray, rings: np.ndarray, maxdist: float, gridspace: float, density=None,
                            part_start=(), ring_start=()):
    """infinite_proj_sentinels (src/border_projection.jl:44-102): grid over the envelope of the region; keep the grid
    points within maxdist of the border and within the region; project them onto the border; the weight of a
    projected point is the summed density of the grid points that land on it (exact coordinate match, accumulated in
    grid order, s1 fastest).  Returns (X_projected 2 x k in first-appearance order, weights k).  The envelope pre-filter
    (:66) cannot change the result and is omitted; grid coordinates are min + k*step in plain FP64 (Julia's ranges use
    twice-precision stepping: last-bit differences possible)."""
    R = np.asarray(rings, dtype=np.float64)
    x1 = R[:, 0].min() + gridspace * np.arange(int(math.floor((R[:, 0].max() - R[:, 0].min()) / gridspace)) + 1)
    x2 = R[:, 1].min() + gridspace * np.arange(int(math.floor((R[:, 1].max() - R[:, 1].min()) / gridspace)) + 1)
    S1, S2 = np.meshgrid(x1, x2, indexing="ij")
    P = np.vstack([S1.ravel(order="F"), S2.ravel(order="F")])            # s1 fastest, as Iterators.product
    Xp, d = projection_points_all(P, border, part_start)
    keep = (d <= maxdist) & points_in_region(P, R, ring_start)
    Pk, Xk = P[:, keep], Xp[:, keep]
    rho = np.ones(Pk.shape[1]) if density is None else np.array([float(density(a, b)) for a, b in zip(Pk[0], Pk[1])])
    keys, order, w = {}, [], []
    for j in range(Xk.shape[1]):
        k = (Xk[0, j], Xk[1, j])
        if k in keys:
            w[keys[k]] += rho[j]
        else:
            keys[k] = len(order)
            order.append(k)
            w.append(rho[j])
    return np.asfortranarray(np.array(order).T.reshape(2, -1)), np.array(w)


def infinite_proj_estim(gpT, gpC, border, rings, maxdist, gridspace, density=None, part_start=(), ring_start=()):
    """infinite_proj_estim (src/border_projection.jl:104-116)."""
    Xb, w = infinite_proj_sentinels(border, rings, maxdist, gridspace, density, part_start, ring_start)
    mu, Sig = cliff_face(gpT, gpC, Xb)
    return weighted_mean(mu, Sig, w)


# ----------------------------------------------------------------------------
# src/placebo_geometry.jl + the placebo_* drivers
# ----------------------------------------------------------------------------
def _trig_deg_kernel(fn: str, y: float) -> float:
    """sin / cos of an angle given in DEGREES with |y| <= 45, the degree -> radian conversion carried in extended precision
    (Julia: sin_kernel / cos_kernel on deg2rad_ext(y), a double-double).  mpmath at 40 digits when present, libm otherwise."""
    try:
        import mpmath
        with mpmath.workprec(140):
            a = mpmath.mpf(y) * mpmath.pi / 180
            return float(mpmath.sin(a) if fn == "sin" else mpmath.cos(a))
    except ImportError:                                             # pragma: no cover
        return math.sin(math.radians(y)) if fn == "sin" else math.cos(math.radians(y))


def sind(x: float) -> float:
    """Base.sind (base/special/trig.jl, Julia 1.4): exact at multiples of 90 degrees."""
    rx = math.copysign(math.fmod(x, 360.0), x)
    arx = abs(rx)
    if rx == 0.0:
        return rx
    if arx < 45.0:
        return _trig_deg_kernel("sin", rx)
    if arx <= 135.0:
        return math.copysign(_trig_deg_kernel("cos", 90.0 - arx), rx)
    if arx == 180.0:
        return math.copysign(0.0, rx)
    if arx < 225.0:
        return _trig_deg_kernel("sin", (180.0 - arx) * math.copysign(1.0, rx))
    if arx <= 315.0:
        return -math.copysign(_trig_deg_kernel("cos", 270.0 - arx), rx)
    return _trig_deg_kernel("sin", rx - math.copysign(360.0, rx))


def cosd(x: float) -> float:
    """Base.cosd: exactly 0 at 90 and 270 degrees."""
    rx = abs(math.fmod(x, 360.0))
    if rx <= 45.0:
        return _trig_deg_kernel("cos", rx)
    if rx < 135.0:
        return _trig_deg_kernel("sin", 90.0 - rx)
    if rx <= 225.0:
        return -_trig_deg_kernel("cos", 180.0 - rx)
    if rx < 315.0:
        return _trig_deg_kernel("sin", rx - 270.0)
    return _trig_deg_kernel("cos", 360.0 - rx)


def tand(x: float) -> float:
    """Base.tand = sind / cosd: tand(45) == 1.0 exactly, tand(90) == Inf."""
    s, c = sind(x), cosd(x)
    if c == 0.0:
        return math.copysign(math.inf, s) if s != 0.0 else math.nan
    return s / c


def _placebo_line(angle: float, shift: float, X: np.ndarray):
    """Common head of left_points / placebo_sentinels (src/placebo_geometry.jl:9-19, 66-75): slope, centroid and the shift
    vector of the placebo line."""
    dydx = tand(angle)
    direction = float(np.sign(cosd(angle + 90.0)))
    if direction == 0.0:
        direction = 1.0
    sx = shift * cosd(angle + 90.0) * direction
    sy = shift * sind(angle + 90.0) * direction
    return dydx, float(np.mean(X[0])), float(np.mean(X[1])), sx, sy


def left_points(angle: float, shift: float, X: np.ndarray) -> np.ndarray:
    """src/placebo_geometry.jl:5-32: isLeft(a, b, p) for every unit, a on the line, b = a + (dx, dy)."""
    dydx, meanx, meany, sx, sy = _placebo_line(angle, shift, X)
    if abs(dydx) < 1.0:
        dx, dy = 1.0, dydx
    else:
        dx, dy = 1.0 / dydx, 1.0
    ax, ay = meanx + sx, meany + sy
    bx, by = meanx + sx + dx, meany + dy + sy
    return ((bx - ax) * (X[1] - ay) - (by - ay) * (X[0] - ax)) > 0


def shift_for_even_split(angle: float, X: np.ndarray) -> float:
    """src/placebo_geometry.jl:103-136: bisection (tol 1e-5, at most 100 iterations) on prop_left - 0.5."""
    maxd = math.sqrt((X[1].max() - X[1].min()) ** 2 + (X[0].max() - X[0].min()) ** 2)

    def f(sh):
        return float(np.mean(left_points(angle, sh, X))) - 0.5

    a, b = -maxd, maxd
    fa = f(a)
    if not fa * f(b) <= 0:
        raise ValueError("No real root in [a,b]")
    i, c = 0, None
    while b - a > 1e-5:
        i += 1
        if i == 100:
            raise RuntimeError("Max iteration exceeded")
        c = (a + b) / 2
        fc = f(c)
        if fc == 0:
            break
        elif fa * fc > 0:
            a, fa = c, fc
        else:
            b = c
    return c


def _julia_range(a: float, b: float, n: int) -> np.ndarray:
    """range(a, stop=b, length=n) for Float64: Julia evaluates it in twice-precision arithmetic, i.e. (to the ulp) the
    correctly rounded values of the exact interpolation -- computed here in rational arithmetic."""
    from fractions import Fraction
    fa, fb = Fraction(a), Fraction(b)
    return np.array([float(fa + (fb - fa) * i / (n - 1)) for i in range(n)]) if n > 1 else np.array([a])


def placebo_sentinels(angle: float, shift: float, X: np.ndarray, npoints: int) -> np.ndarray:
    """src/placebo_geometry.jl:60-101: npoints equally spaced between the two crossings of the bounding box."""
    minx, maxx, miny, maxy = X[0].min(), X[0].max(), X[1].min(), X[1].max()
    dydx, meanx, meany, sx, sy = _placebo_line(angle, shift, X)
    cx, cy = meanx + sx, meany + sy
    with np.errstate(divide="ignore", invalid="ignore"):
        y_from_minx = cy + dydx * (minx - cx)
        y_from_maxx = cy + dydx * (maxx - cx)
        x_from_miny = cx + np.float64(miny - cy) / np.float64(dydx)
        x_from_maxy = cx + np.float64(maxy - cy) / np.float64(dydx)
    ext = []
    if miny < y_from_minx < maxy:
        ext.append((minx, y_from_minx))
    if miny < y_from_maxx < maxy:
        ext.append((maxx, y_from_maxx))
    if minx < x_from_miny < maxx:
        ext.append((x_from_miny, miny))
    if minx < x_from_maxy < maxx:
        ext.append((x_from_maxy, maxy))
    assert len(ext) == 2
    return np.asfortranarray(np.vstack([_julia_range(float(ext[0][0]), float(ext[1][0]), npoints),
                                        _julia_range(float(ext[0][1]), float(ext[1][1]), npoints)]))


def _placebo_fits(angle: float, X: np.ndarray, Y: np.ndarray, spec: KernelSpec, log_noise: float):
    shift = shift_for_even_split(angle, X)
    left = left_points(angle, shift, X)
    gl = gp_fit(spec, np.asfortranarray(X[:, left]), Y[left], log_noise)
    gr = gp_fit(spec, np.asfortranarray(X[:, ~left]), Y[~left], log_noise)
    return shift, left, gl, gr


def placebo_chi(angle: float, X: np.ndarray, Y: np.ndarray, spec: KernelSpec, log_noise: float, Z: np.ndarray) -> float:
    """src/hypotest_chi2.jl:62-72 (100 sentinels); Z: n x nsim standard normals, rows in [left; right] order."""
    shift, _, gl, gr = _placebo_fits(angle, X, Y, spec, log_noise)
    return boot_chi2test(gl, gr, placebo_sentinels(angle, shift, X, 100), Z)[0]


def placebo_invvar(angle: float, X: np.ndarray, Y: np.ndarray, spec: KernelSpec, log_noise: float) -> float:
    """src/hypotest_invvar.jl:162-171."""
    shift, _, gl, gr = _placebo_fits(angle, X, Y, spec, log_noise)
    return pval_invvar_calib(gl, gr, placebo_sentinels(angle, shift, X, 100))


def placebo_mll(angle: float, X: np.ndarray, Y: np.ndarray, spec: KernelSpec, log_noise: float, Z: np.ndarray) -> float:
    """src/hypotest_mll.jl:47-56."""
    _, _, gl, gr = _placebo_fits(angle, X, Y, spec, log_noise)
    return boot_mlltest(gl, gr, Z)[0]
