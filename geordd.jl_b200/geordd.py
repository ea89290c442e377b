"""Host-side mirror of the GeoRDD.jl entry points on the hot path (same names, argument meaning and
error behaviour as the Julia package), calling libgeordd_b200.so through the C ABI.

Julia is not available in this environment, so this module plays the role of the thin Julia wrappers
(julia/GeoRDDB200.jl shows the `ccall` twin): it pattern-matches GaussianProcesses.jl-style kernel
objects into the `geordd_kernel` POD, keeps host arrays alive across the call, converts status codes
into exceptions (`PosDefException`, `ArgumentError` -> ValueError) and owns the device handles.

All numerics run on the GPU.  Nothing here falls back to NumPy: if the library cannot be loaded or
no B200 is present the constructors raise.

Reference citations are to /root/reference (GeoRDD.jl v0.2.0).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, Hashable, Iterable, List, Optional, Sequence

import numpy as np

from . import capi
from .capi import GeorddError, Kernel as _CKernel, PosDefException, dptr, fmat


# ----------------------------------------------------------------------------------------------
# GaussianProcesses.jl-style kernel / mean objects (parameterisation: SURVEY App. A.1)
# ----------------------------------------------------------------------------------------------
class KernelBase:
    def __add__(self, other):
        return SumKernel(self, other)


@dataclass
class SEIso(KernelBase):
    ll: float
    lsigma: float

    @property
    def ell2(self):  # field name in GaussianProcesses: l2
        return math.exp(2.0 * self.ll)

    @property
    def sigma2(self):
        return math.exp(2.0 * self.lsigma)


@dataclass
class _MatIso(KernelBase):
    ll: float
    lsigma: float

    @property
    def ell(self):
        return math.exp(self.ll)

    @property
    def sigma2(self):
        return math.exp(2.0 * self.lsigma)


class Mat12Iso(_MatIso):
    pass


class Mat32Iso(_MatIso):
    pass


class Mat52Iso(_MatIso):
    pass


@dataclass
class Const(KernelBase):
    lsigma: float

    @property
    def sigma2(self):
        return math.exp(2.0 * self.lsigma)


@dataclass
class LinIso(KernelBase):
    ll: float

    @property
    def ell2(self):
        return math.exp(2.0 * self.ll)


@dataclass
class SumKernel(KernelBase):
    kleft: KernelBase
    kright: KernelBase


@dataclass
class FixedKernel(KernelBase):
    kernel: KernelBase


def fix(k):
    """GaussianProcesses.fix(k): same values, parameters hidden from get_params."""
    return FixedKernel(k)


@dataclass
class MeanZero:
    pass


@dataclass
class MeanConst:
    beta: float


_KIND = {SEIso: 0, Mat12Iso: 1, Mat32Iso: 2, Mat52Iso: 3}


def _unfix(k):
    return k.kernel if isinstance(k, FixedKernel) else k


def kernel_spec(kernel, mean=None) -> _CKernel:
    """Pattern-match SEIso / MatXXIso / SumKernel(spatial, Const) (+ fix) and MeanZero / MeanConst into the
    ABI struct; anything else is an ArgumentError (no fallback, BASELINE north_star)."""
    k = _unfix(kernel)
    const_s2 = 0.0
    if isinstance(k, SumKernel):
        a, b = _unfix(k.kleft), _unfix(k.kright)
        if isinstance(a, Const):
            a, b = b, a
        if not isinstance(b, Const) or type(a) not in _KIND:
            raise ValueError("ArgumentError: only SEIso/Mat12Iso/Mat32Iso/Mat52Iso (+ Const) kernels are supported")
        const_s2 = b.sigma2
        k = a
    if type(k) not in _KIND:
        raise ValueError("ArgumentError: only SEIso/Mat12Iso/Mat32Iso/Mat52Iso (+ Const) kernels are supported")
    ell = math.sqrt(k.ell2) if isinstance(k, SEIso) else k.ell
    if mean is None or isinstance(mean, MeanZero):
        mc = 0.0
    elif isinstance(mean, MeanConst):
        mc = float(mean.beta)
    else:
        raise ValueError("ArgumentError: only MeanZero / MeanConst are supported")
    return _CKernel(_KIND[type(k)], ell, k.sigma2, const_s2, mc)


def get_params(kernel) -> List[float]:
    """get_params order of GaussianProcesses kernels: [log ell, log sigma] (+ [log sigma_const])."""
    if isinstance(kernel, FixedKernel):
        return []
    if isinstance(kernel, SumKernel):
        return get_params(kernel.kleft) + get_params(kernel.kright)
    if isinstance(kernel, (SEIso, _MatIso)):
        return [kernel.ll, kernel.lsigma]
    if isinstance(kernel, Const):
        return [kernel.lsigma]
    if isinstance(kernel, LinIso):
        return [kernel.ll]
    raise ValueError("unsupported kernel")


def set_params(kernel, hyp: Sequence[float]) -> None:
    if isinstance(kernel, FixedKernel):
        return
    if isinstance(kernel, SumKernel):
        nl = len(get_params(kernel.kleft))
        set_params(kernel.kleft, hyp[:nl])
        set_params(kernel.kright, hyp[nl:])
    elif isinstance(kernel, (SEIso, _MatIso)):
        kernel.ll, kernel.lsigma = float(hyp[0]), float(hyp[1])
    elif isinstance(kernel, Const):
        kernel.lsigma = float(hyp[0])
    elif isinstance(kernel, LinIso):
        kernel.ll = float(hyp[0])


# ----------------------------------------------------------------------------------------------
# Context
# ----------------------------------------------------------------------------------------------
class Context:
    """One GPU, one stream (geordd_ctx).  Raises if no B200-class device / library is available."""

    def __init__(self, device: int = 0):
        self.lib = capi.load_library()
        self._h = C.c_void_p()
        rc = self.lib.geordd_ctx_create(C.byref(self._h), int(device))
        if rc != 0:
            raise GeorddError(rc, "no usable sm_100 CUDA device (this package has no CPU fallback)")
        self.device = device

    def check(self, rc: int) -> None:
        if rc == 0:
            return
        msg = self.lib.geordd_last_error(self._h).decode()
        if rc > 0:
            raise PosDefException(rc, msg)
        if rc == capi_ERR_ARG:
            raise ValueError("ArgumentError: " + msg)
        raise GeorddError(rc, msg)

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self.check(self.lib.geordd_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def sync(self) -> None:
        self.check(self.lib.geordd_ctx_sync(self._h))

    def set_max_batch(self, n: int) -> None:
        self.check(self.lib.geordd_ctx_set_max_batch(self._h, int(n)))

    def set_mll_forward_only(self, on: bool) -> None:
        """Opt-in: quadratic forms of the mll test as |L^-1 (y-m)|^2 / z'z (no backward solves); see the header."""
        self.check(self.lib.geordd_ctx_set_mll_forward_only(self._h, 1 if on else 0))
        self._mll_forward_only = bool(on)

    def last_min_margin(self):
        """(chi2, invvar p, delta-mll) minimum |statistic - threshold| of the last Monte-Carlo call: the certificate that the
        exceedance count is exact (it can only change if statistics move by more than this)."""
        out = (C.c_double * 3)()
        self.check(self.lib.geordd_ctx_last_min_margin(self._h, out))
        return tuple(out)

    def profile(self, enable: bool) -> None:
        self.check(self.lib.geordd_ctx_profile(self._h, 1 if enable else 0))

    def profile_get(self, name: str = "*"):
        ms, n = C.c_double(), C.c_longlong()
        self.check(self.lib.geordd_ctx_profile_get(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def dmma_peak_tflops(self) -> float:
        t = C.c_double()
        self.check(self.lib.geordd_bench_dmma_peak(self._h, C.byref(t)))
        return t.value

    def bench_potrf(self, n: int, reps: int = 3):
        b, m = C.c_double(), C.c_double()
        self.check(self.lib.geordd_bench_potrf(self._h, int(n), int(reps), C.byref(b), C.byref(m)))
        return b.value, m.value

    def close(self):
        if self._h:
            self.lib.geordd_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


capi_ERR_ARG = -1
_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("GEORDD_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _default_ctx


# ----------------------------------------------------------------------------------------------
# Covariance assembly (K1)
# ----------------------------------------------------------------------------------------------
def cov(kernel, X1, X2=None, *, diag_add: float = 0.0, ctx: Optional[Context] = None) -> np.ndarray:
    """cov(kernel, X1, X2) / cov(kernel, X) of GaussianProcesses (call sites src/hypotest_chi2.jl:38-39)."""
    ctx = ctx or default_context()
    spec = kernel_spec(kernel)
    X1 = fmat(X1)
    if X2 is None:
        n = X1.shape[1]
        K = np.empty((n, n), order="F")
        ctx.check(ctx.lib.geordd_cov_sym(ctx._h, C.byref(spec), dptr(X1), X1.shape[0], n, float(diag_add), dptr(K)))
        return K
    X2 = fmat(X2)
    K = np.empty((X1.shape[1], X2.shape[1]), order="F")
    ctx.check(ctx.lib.geordd_cov_cross(ctx._h, C.byref(spec), dptr(X1), X1.shape[1], dptr(X2), X2.shape[1], X1.shape[0],
                                       dptr(K)))
    return K


# ----------------------------------------------------------------------------------------------
# GPE
# ----------------------------------------------------------------------------------------------
class Normal:
    """Distributions.Normal(mu, sigma) with the cdf/ccdf/quantile GeoRDD uses."""

    def __init__(self, mu: float, sigma: float):
        self.mu, self.sigma = float(mu), float(sigma)

    def cdf(self, x):
        return math.erfc(-((x - self.mu) / self.sigma) * 0.7071067811865476) / 2.0

    def ccdf(self, x):
        return math.erfc(((x - self.mu) / self.sigma) * 0.7071067811865476) / 2.0

    def quantile(self, p):
        from scipy.special import erfcinv
        return self.mu - self.sigma * math.sqrt(2.0) * float(erfcinv(2.0 * p))

    def __repr__(self):
        return f"Normal(mu={self.mu}, sigma={self.sigma})"


class GPE:
    """GPE(x, y, mean, kernel, logNoise) as GeoRDD constructs it (src/region_gp_fit.jl:3,12,28;
    src/hypotest.jl:4): assembles K + (exp(2 logNoise) + eps) I, make_posdef!, alpha, mll -- on the GPU."""

    def __init__(self, x, y, mean, kernel, logNoise: float, ctx: Optional[Context] = None, want_chol: bool = False):
        self.ctx = ctx or default_context()
        self.x = fmat(x)
        self.y = np.array(y, dtype=np.float64)
        self.mean, self.kernel, self.logNoise = mean, kernel, float(logNoise)
        self.dim, self.nobs = self.x.shape
        if self.y.shape[0] != self.nobs:
            raise ValueError("ArgumentError: x and y have incompatible sizes")
        spec = kernel_spec(kernel, mean)
        self._h = C.c_void_p()
        self.alpha = np.empty(self.nobs)
        mll, jr = C.c_double(), C.c_int()
        self.cholU = np.empty((self.nobs, self.nobs), order="F") if want_chol else None
        rc = self.ctx.lib.geordd_gp_fit(self.ctx._h, C.byref(spec), dptr(self.x), self.dim, self.nobs, dptr(self.y),
                                        self.logNoise, C.byref(self._h), dptr(self.alpha), C.byref(mll), dptr(self.cholU),
                                        C.byref(jr))
        self.ctx.check(rc)
        self.mll = mll.value
        self.jitter_rounds = jr.value

    # GeoRDD.GPE(rd::RegionData, args...) (src/region_gp_fit.jl:3)
    @classmethod
    def from_region(cls, rd: "RegionData", mean, kernel, logNoise, **kw):
        return cls(rd.x, rd.y, mean, kernel, logNoise, **kw)

    @classmethod
    def fit_many(cls, xs, ys, mean, kernel, logNoise: float, ctx: Optional[Context] = None) -> List["GPE"]:
        """The independent fits that GPRealisations' constructor (src/GPrealisations.jl:10-19) and the placebo drivers
        (src/hypotest_chi2.jl:66-69) make one after another, as ONE call (geordd_gp_fit_batch): all factorisations run in
        a single dataflow launch.  Same results as [GPE(x, y, mean, kernel, logNoise) for ...], bit for bit."""
        ctx = ctx or default_context()
        nfit = len(xs)
        gps = [cls.__new__(cls) for _ in range(nfit)]
        for g, x, y in zip(gps, xs, ys):
            g.ctx, g.x, g.y = ctx, fmat(x), np.array(y, dtype=np.float64)
            g.mean, g.kernel, g.logNoise = mean, kernel, float(logNoise)
            g.dim, g.nobs = g.x.shape
            if g.y.shape[0] != g.nobs:
                raise ValueError("ArgumentError: x and y have incompatible sizes")
            g._h, g.alpha, g.cholU = C.c_void_p(), np.empty(g.nobs), None
        if len({g.dim for g in gps}) != 1:
            raise ValueError("ArgumentError: all regions must have the same spatial dimension")
        spec = kernel_spec(kernel, mean)
        specs = (_CKernel * nfit)(*([spec] * nfit))
        PD = capi.c_double_p
        Xp = (PD * nfit)(*[dptr(g.x) for g in gps])
        Yp = (PD * nfit)(*[dptr(g.y) for g in gps])
        Ap = (PD * nfit)(*[dptr(g.alpha) for g in gps])
        ns = (C.c_int * nfit)(*[g.nobs for g in gps])
        lns = (C.c_double * nfit)(*([float(logNoise)] * nfit))
        hs = (C.c_void_p * nfit)()
        mll, jr, info = (C.c_double * nfit)(), (C.c_int * nfit)(), (C.c_int * nfit)()
        rc = ctx.lib.geordd_gp_fit_batch(ctx._h, nfit, specs, Xp, gps[0].dim, ns, Yp, lns, hs, Ap, mll, jr, info)
        ctx.check(rc)
        for f, g in enumerate(gps):
            g._h, g.mll, g.jitter_rounds = C.c_void_p(hs[f]), mll[f], jr[f]
        return gps

    def cK_mat(self) -> np.ndarray:
        K = np.empty((self.nobs, self.nobs), order="F")
        self.ctx.check(self.ctx.lib.geordd_gp_get_K(self._h, dptr(K)))
        return K

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.ctx.lib.geordd_gp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def update_alpha(gp: GPE) -> None:
    """update_alpha!(gp) (src/gp_utils.jl:15-18) -- also refreshes gp.mll like mll(gp) (:19-22)."""
    mll = C.c_double()
    gp.ctx.check(gp.ctx.lib.geordd_gp_set_y(gp._h, dptr(gp.y), dptr(gp.alpha), C.byref(mll)))
    gp.mll = mll.value


def mll(gp: GPE) -> float:
    """mll(gp) (src/gp_utils.jl:19-22) for the current gp.y / gp.alpha."""
    return gp.mll


def modifiable(gp: GPE) -> GPE:
    """modifiable(gp) (src/gp_utils.jl:10-14): a copy whose y may be changed.  The 8-argument GPE
    constructor re-initialises (re-factors) the shared cK; so does this."""
    return GPE(gp.x, gp.y.copy(), gp.mean, gp.kernel, gp.logNoise, ctx=gp.ctx)


def predict_mu(gp: GPE, X, cK=None) -> np.ndarray:
    """predict_mu(gp, X, cK) = mean(gp.mean, X) + cK*gp.alpha (src/gp_utils.jl:1-5); cK is recomputed on the
    device from (kernel, X, gp.x), the argument is accepted for signature compatibility."""
    X = fmat(X)
    mu = np.empty(X.shape[1])
    gp.ctx.check(gp.ctx.lib.geordd_gp_predict_mu(gp._h, dptr(X), X.shape[1], dptr(mu)))
    return mu


def prior_rand(gp: GPE, nsim: int = 1, *, seed: int = 0, draw0: int = 0, Z=None) -> np.ndarray:
    """prior_rand(gp) (src/hypotest.jl:13-18) for nsim draws: columns of m + L z."""
    Zf = fmat(Z) if Z is not None else None
    if Zf is not None:
        nsim = Zf.shape[1] if Zf.ndim == 2 else 1
    Y = np.empty((gp.nobs, nsim), order="F")
    gp.ctx.check(gp.ctx.lib.geordd_gp_prior_rand(gp._h, nsim, seed, draw0, dptr(Zf), dptr(Y)))
    return Y


# ----------------------------------------------------------------------------------------------
# cliff face and LATE reducers
# ----------------------------------------------------------------------------------------------
def cliff_face(gpT: GPE, gpC: GPE, sentinels):
    """cliff_face(gpT, gpC, sentinels) -> (mu_posterior, Sigma_posterior) (src/cliff_face.jl:3-9)."""
    Xb = fmat(sentinels)
    nb = Xb.shape[1]
    mu, S = np.empty(nb), np.empty((nb, nb), order="F")
    gpT.ctx.check(gpT.ctx.lib.geordd_cliff_face(gpT._h, gpC._h, dptr(Xb), nb, dptr(mu), dptr(S)))
    return mu, S


def add_nugget(A, nugget):
    """src/average_treatment_effect.jl:1-2."""
    return A + nugget * np.eye(A.shape[0])


def inverse_variance(mu, Sigma, nugget: float = 1e-10, ctx: Optional[Context] = None) -> Normal:
    """inverse_variance(mu, Sigma; nugget) (src/average_treatment_effect.jl:3-11)."""
    ctx = ctx or default_context()
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    S = fmat(Sigma)
    tau, var = C.c_double(), C.c_double()
    ctx.check(ctx.lib.geordd_inverse_variance(ctx._h, dptr(mu), dptr(S), mu.shape[0], float(nugget), C.byref(tau),
                                              C.byref(var)))
    return Normal(tau.value, math.sqrt(var.value))


def unweighted_mean(mu, Sigma) -> Normal:
    """src/average_treatment_effect.jl:13-22 (O(nb^2) host reduction of an nb x nb result)."""
    n = len(mu)
    return Normal(float(np.mean(mu)), math.sqrt(float(np.sum(Sigma)) / n ** 2))


def weighted_mean(mu, Sigma, w) -> Normal:
    """src/average_treatment_effect.jl:24-29."""
    w = np.asarray(w, dtype=np.float64)
    sw = float(np.sum(w))
    return Normal(float(np.sum(mu * w)) / sw, math.sqrt(float(w @ (np.asarray(Sigma) @ w)) / sw ** 2))


def weight_at_units(gp: GPE, sentinels, weights) -> np.ndarray:
    """weight_at_units(gp, sentinels, weights) (src/weight_at_units.jl:12-18)."""
    Xb = fmat(sentinels)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    assert Xb.shape[1] == w.shape[0]
    out = np.empty(gp.nobs)
    gp.ctx.check(gp.ctx.lib.geordd_weight_at_units(gp._h, dptr(Xb), Xb.shape[1], dptr(w), dptr(out)))
    return out


def weights_unif(gpT: GPE, gpC: GPE, Xb):
    """weights_unif with the sentinels already placed (src/weight_at_units.jl:23-28)."""
    nb = np.asarray(Xb).shape[1]
    wb = np.ones(nb)
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


def generic_solve(A, B, ctx: Optional[Context] = None) -> np.ndarray:
    """Julia's `A \\ B` on a square dense Matrix (lu with partial pivoting) on the device."""
    ctx = ctx or default_context()
    A = fmat(np.array(A, dtype=np.float64))
    Bm = fmat(np.array(B, dtype=np.float64).reshape(A.shape[0], -1))
    ctx.check(ctx.lib.geordd_solve_general(ctx._h, dptr(A), A.shape[0], dptr(Bm), Bm.shape[1]))
    return Bm.reshape(np.shape(B))


def weights_invvar(gpT: GPE, gpC: GPE, Xb):
    """weights_invvar (src/weight_at_units.jl:33-39) with the sentinels already placed:
    w_border = Sigma_post \\ ones(nb)."""
    _, S = cliff_face(gpT, gpC, Xb)
    wb = generic_solve(S, np.ones(S.shape[0]), ctx=gpT.ctx)
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


# ----------------------------------------------------------------------------------------------
# Border geometry feeding the projected estimators (src/geometry.jl:22-30, src/border_projection.jl:4-31,
# src/weight_at_units.jl:44-53,80-86)
# ----------------------------------------------------------------------------------------------
def sentinels(border, nsent: int) -> np.ndarray:
    """sentinels(border, nsent) (src/geometry.jl:22-30): nsent points equally spaced in arclength along the border
    polyline (nv x 2), end points included -> 2 x nsent.  O(nv + nsent) host geometry."""
    V = np.asarray(border, dtype=np.float64)
    seg = np.sqrt(np.sum(np.diff(V, axis=0) ** 2, axis=1))
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    q = np.linspace(0.0, cum[-1], int(nsent))
    idx = np.clip(np.searchsorted(cum, q, side="right") - 1, 0, len(seg) - 1)
    safe = np.where(seg[idx] > 0, seg[idx], 1.0)
    t = np.where(seg[idx] > 0, (q - cum[idx]) / safe, 0.0)
    return fmat((V[idx] + t[:, None] * (V[idx + 1] - V[idx])).T)


def _as_units(gp_or_X):
    return fmat(gp_or_X.x if isinstance(gp_or_X, GPE) else gp_or_X)


def projection_points_all(gp_or_X, border, part_start=(), ctx: Optional[Context] = None):
    """Nearest point on the border and distance for EVERY unit (the per-unit LibGEOS distance / nearestPoints calls
    of src/border_projection.jl:8-19, on the device).  border: nv x 2 vertices; part_start: vertex indices at which
    further lines of a MultiLineString start.  Returns (Xproj 2 x n, dist n)."""
    X = _as_units(gp_or_X)
    if ctx is None:
        ctx = gp_or_X.ctx if isinstance(gp_or_X, GPE) else default_context()
    V = np.asarray(border, dtype=np.float64)
    if V.ndim != 2 or V.shape[1] != 2:
        raise ValueError("border must be an nv x 2 array of vertices")
    Vt = fmat(V.T)                                                           # 2 x nv column-major
    ps = np.ascontiguousarray(part_start, dtype=np.int32)
    n = X.shape[1]
    Xp, d = np.empty((2, n), order="F"), np.empty(n)
    ctx.check(ctx.lib.geordd_projection_points(ctx._h, dptr(X), n, dptr(Vt), V.shape[0],
                                               ps.ctypes.data_as(C.POINTER(C.c_int)) if ps.size else None, int(ps.size),
                                               dptr(Xp), dptr(d)))
    return Xp, d


def projection_points(gp_or_X, border, maxdist: float, part_start=(), ctx: Optional[Context] = None) -> np.ndarray:
    """projection_points(gp, border, maxdist) (src/border_projection.jl:4-22): projections of the units that lie
    within maxdist of the border, in unit order."""
    Xp, d = projection_points_all(gp_or_X, border, part_start, ctx)
    return fmat(Xp[:, d <= maxdist])


def proj_estimator(gpT: GPE, gpC: GPE, border, maxdist: float, part_start=()) -> Normal:
    """proj_estimator(gpT, gpC, border, maxdist) (src/border_projection.jl:24-31; test/test_simulated.jl:58)."""
    Xb = np.hstack([projection_points(gpT, border, maxdist, part_start), projection_points(gpC, border, maxdist, part_start)])
    mu, S = cliff_face(gpT, gpC, fmat(Xb))
    return unweighted_mean(mu, S)


def weights_proj(gpT: GPE, gpC: GPE, border, maxdist: float, part_start=()):
    """weights_proj (src/weight_at_units.jl:44-53): unit weights of the projected finite-population estimator."""
    Xb = fmat(np.hstack([projection_points(gpT, border, maxdist, part_start), projection_points(gpC, border, maxdist, part_start)]))
    wb = np.ones(Xb.shape[1])
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


def weights_rho(gpT: GPE, gpC: GPE, density, nb: int, border):
    """weights_rho (src/weight_at_units.jl:80-86): density-weighted sentinels equally spaced along the border."""
    Xb = sentinels(border, nb)
    wb = np.array([float(density(x1, x2)) for x1, x2 in zip(Xb[0], Xb[1])])
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


def points_in_region(P, rings, ring_start=(), ctx: Optional[Context] = None) -> np.ndarray:
    """LibGEOS.within(point, region) for the columns of P (2 x n) on the device.  rings: nr x 2 vertices of all closed
    rings (shells and holes) of the (Multi)Polygon; ring_start: vertex indices at which further rings start."""
    ctx = ctx or default_context()
    Pm = fmat(P)
    R = np.asarray(rings, dtype=np.float64)
    if R.ndim != 2 or R.shape[1] != 2:
        raise ValueError("rings must be an nr x 2 array of vertices")
    Rt = fmat(R.T)
    rs = np.ascontiguousarray(ring_start, dtype=np.int32)
    out = np.empty(Pm.shape[1], dtype=np.uint8)
    ctx.check(ctx.lib.geordd_points_in_region(ctx._h, dptr(Pm), Pm.shape[1], dptr(Rt), R.shape[0],
                                              rs.ctypes.data_as(C.POINTER(C.c_int)) if rs.size else None, int(rs.size),
                                              out.ctypes.data_as(C.POINTER(C.c_ubyte))))
    return out.astype(bool)


def infinite_proj_sentinels(gpT: GPE, gpC: GPE, border, region_rings, maxdist: float, gridspace: float, density=None,
                            part_start=(), ring_start=()):
    """infinite_proj_sentinels (src/border_projection.jl:44-102): grid over the region's envelope, filtered by distance
    to the border and within(region) and projected onto the border ON THE DEVICE (the reference spends 5-28 s per call
    in LibGEOS here); the per-point density and the accumulation of weights on coinciding projections stay on the
    host.  Returns (X_projected 2 x k, weights k)."""
    ctx = gpT.ctx
    R = np.asarray(region_rings, dtype=np.float64)
    x1 = R[:, 0].min() + gridspace * np.arange(int(math.floor((R[:, 0].max() - R[:, 0].min()) / gridspace)) + 1)
    x2 = R[:, 1].min() + gridspace * np.arange(int(math.floor((R[:, 1].max() - R[:, 1].min()) / gridspace)) + 1)
    S1, S2 = np.meshgrid(x1, x2, indexing="ij")
    P = fmat(np.vstack([S1.ravel(order="F"), S2.ravel(order="F")]))       # s1 fastest, as Iterators.product
    Xp, d = projection_points_all(P, border, part_start, ctx=ctx)
    keep = (d <= maxdist) & points_in_region(P, R, ring_start, ctx=ctx)
    Pk, Xk = P[:, keep], Xp[:, keep]
    rho = np.ones(Pk.shape[1]) if density is None else np.array([float(density(a, b)) for a, b in zip(Pk[0], Pk[1])])
    # weights of coinciding projections: first-appearance order, sequential accumulation in grid order
    uniq, first, inv = np.unique(Xk.T, axis=0, return_index=True, return_inverse=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    w = np.zeros(len(first))
    np.add.at(w, rank[np.ravel(inv)], rho)
    Xb = np.empty((2, len(first)), order="F")
    Xb[:, rank] = uniq.T
    return Xb, w


def infinite_proj_estim(gpT: GPE, gpC: GPE, border, region_rings, maxdist: float, gridspace: float, density=None,
                        part_start=(), ring_start=()) -> Normal:
    """infinite_proj_estim (src/border_projection.jl:104-116)."""
    Xb, w = infinite_proj_sentinels(gpT, gpC, border, region_rings, maxdist, gridspace, density, part_start, ring_start)
    mu, S = cliff_face(gpT, gpC, Xb)
    return weighted_mean(mu, S, w)


def weights_geo(gpT: GPE, gpC: GPE, border, region_rings, maxdist: float, gridspace: float, **kw):
    """weights_geo (src/weight_at_units.jl:58-64)."""
    Xb, wb = infinite_proj_sentinels(gpT, gpC, border, region_rings, maxdist, gridspace, None, **kw)
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


def weights_pop(gpT: GPE, gpC: GPE, density, border, region_rings, maxdist: float, gridspace: float, **kw):
    """weights_pop (src/weight_at_units.jl:69-75)."""
    Xb, wb = infinite_proj_sentinels(gpT, gpC, border, region_rings, maxdist, gridspace, density, **kw)
    return Xb, wb, np.concatenate([weight_at_units(gpT, Xb, wb), -weight_at_units(gpC, Xb, wb)])


# ----------------------------------------------------------------------------------------------
# Null handle: make_null + draw-independent setup, and the Monte-Carlo tests
# ----------------------------------------------------------------------------------------------
class NullGPE:
    """make_null(gpT, gpC) (src/hypotest.jl:1-12): the pooled GP, plus -- once sentinels are attached -- the
    draw-independent state of nsim_chi / nsim_invvar_pval_uncalib / nsim_power."""

    def __init__(self, gpT: GPE, gpC: GPE, Xb=None, nugget: float = 0.0):
        self.ctx = gpT.ctx
        self.gpT, self.gpC = gpT, gpC     # keep alive: the handle borrows them
        self._h = C.c_void_p()
        self.Xb = fmat(Xb) if Xb is not None else None
        nb = self.Xb.shape[1] if self.Xb is not None else 0
        jr, m = C.c_int(), C.c_double()
        rc = self.ctx.lib.geordd_null_prepare(gpT._h, gpC._h, dptr(self.Xb), nb, float(nugget), C.byref(self._h),
                                              C.byref(jr), C.byref(m))
        self.ctx.check(rc)
        self.jitter_rounds = jr.value
        self.mll = m.value
        self.nobs = gpT.nobs + gpC.nobs
        self.nugget = nugget
        self.y = np.concatenate([gpT.y, gpC.y])
        self.alpha = None

    def set_sentinels(self, Xb, nugget: float = 0.0) -> None:
        self.Xb = fmat(Xb)
        jr = C.c_int()
        self.ctx.check(self.ctx.lib.geordd_null_set_sentinels(self._h, dptr(self.Xb), self.Xb.shape[1], float(nugget),
                                                              C.byref(jr)))
        self.jitter_rounds = jr.value
        self.nugget = nugget

    def _ensure(self, Xb, nugget):
        Xb = fmat(Xb)
        if self.Xb is None or self.Xb.shape != Xb.shape or not np.array_equal(self.Xb, Xb) or self.nugget != nugget:
            self.set_sentinels(Xb, nugget)

    def chistat_obs(self) -> float:
        """chistat(gpT, gpC, Xb) of the observed outcomes from the cliff face this handle evaluated when its sentinels were
        set (bit-identical to chistat(); saves boot_chi2test the second cliff face)."""
        out = C.c_double()
        self.ctx.check(self.ctx.lib.geordd_null_chistat_obs(self._h, C.byref(out)))
        return out.value

    def invvar_obs(self):
        """(tau_hat, var, p) of pval_invvar_uncalib(gpT, gpC, Xb; nugget) from this handle's cliff face."""
        tau, var, p = C.c_double(), C.c_double(), C.c_double()
        self.ctx.check(self.ctx.lib.geordd_null_invvar_obs(self._h, C.byref(tau), C.byref(var), C.byref(p)))
        return tau.value, var.value, p.value

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.ctx.lib.geordd_null_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_null(gpT: GPE, gpC: GPE) -> NullGPE:
    """make_null(gpT, gpC): pooled null GP with T's kernel, mean and noise (src/hypotest.jl:6-12)."""
    return NullGPE(gpT, gpC)


def _zarg(Z, n):
    if Z is None:
        return None, None
    Zf = fmat(Z)
    assert Zf.shape[0] == n, "Z must be n x nsim (column s = randn(n) of draw s)"
    return Zf, Zf.shape[1]


def chistat(gpT: GPE, gpC: GPE, Xb) -> float:
    """chistat(gpT, gpC, Xb) observed (src/hypotest_chi2.jl:4-10; generic `\\` = LU on the raw Sigma)."""
    Xb = fmat(Xb)
    out = C.c_double()
    gpT.ctx.check(gpT.ctx.lib.geordd_chistat_obs(gpT._h, gpC._h, dptr(Xb), Xb.shape[1], C.byref(out)))
    return out.value


def nsim_chi(gpT: GPE, gpC: GPE, gpNull: NullGPE, Xb, nsim: int, *, seed: int = 0, draw0: int = 0, Z=None,
             chi_obs: float = math.inf, return_count: bool = False, sharded: bool = False):
    """nsim_chi(gpT, gpC, gpNull, Xb, nsim) (src/hypotest_chi2.jl:33-52): vector of nsim chi2 null draws.
    sharded=True (under torch.distributed): every rank simulates its shard, the full vector is all-gathered."""
    if sharded:
        total = np.asarray(Z).shape[1] if Z is not None else int(nsim)
        local, kw, gather = _vector_shard(total, dict(seed=seed, draw0=draw0, Z=Z))
        part = nsim_chi(gpT, gpC, gpNull, Xb, local, chi_obs=chi_obs, **kw) if local > 0 else np.empty(0)
        return gather(part)
    gpNull._ensure(Xb, 0.0)
    Zf, nz = _zarg(Z, gpNull.nobs)
    nsim = nz if nz is not None else int(nsim)
    stats = np.empty(nsim)
    cnt = C.c_longlong()
    gpT.ctx.check(gpT.ctx.lib.geordd_null_chi2(gpNull._h, nsim, seed, draw0, dptr(Zf), float(chi_obs), dptr(stats),
                                               C.byref(cnt)))
    return (stats, cnt.value) if return_count else stats


def _draw_shard(nsim: int, kw: dict):
    """Multi-GPU boot_* drivers: when the caller runs one process per GPU under torch.distributed, every
    rank simulates its contiguous shard of the draw indices and the exceedance tallies are summed with one all-reduce;
    draws are keyed by (seed, index), so the p-value does not depend on the number of GPUs.  Returns
    (local nsim, kwargs with draw0 / Z adjusted, reducer)."""
    from . import parallel
    if not parallel.active():
        return nsim, kw, (lambda cnt: cnt)
    import torch.distributed as dist
    lo, cnt = parallel.shard_draws(nsim, dist.get_rank(), dist.get_world_size())
    kw = dict(kw)
    kw["draw0"] = int(kw.get("draw0", 0)) + lo
    if kw.get("Z") is not None:
        kw["Z"] = np.asarray(kw["Z"])[:, lo:lo + cnt]
    return cnt, kw, (lambda c: parallel.allreduce_tallies([c])[0])


def _vector_shard(nsim: int, seed_kw: dict):
    """Vector-returning nsim_* forms with sharded=True: this rank simulates its contiguous shard of the draw indices and the
    statistics are assembled with one all-gather (8 * nsim bytes).  Returns (local nsim, kwargs, gather)."""
    from . import parallel
    if not parallel.active():
        return nsim, seed_kw, (lambda v: v)
    import torch.distributed as dist
    lo, cnt = parallel.shard_draws(nsim, dist.get_rank(), dist.get_world_size())
    kw = dict(seed_kw)
    kw["draw0"] = int(kw.get("draw0", 0)) + lo
    if kw.get("Z") is not None:
        kw["Z"] = np.asarray(kw["Z"])[:, lo:lo + cnt]
    return cnt, kw, (lambda v: parallel.gather_stats(v, nsim))


def boot_chi2test(gpT: GPE, gpC: GPE, Xb, nsim: int, **kw) -> float:
    """boot_chi2test(gpT, gpC, Xb, nsim) (src/hypotest_chi2.jl:54-59): mean(chi_sims .> chi_obs)."""
    if isinstance(gpT, GroupGPE):
        return _group_boot("chi", gpT, gpC, Xb, nsim, **kw)
    gpNull = NullGPE(gpT, gpC, Xb, 0.0)       # make_null + the sentinel-dependent setup (one cliff face)
    chi_obs = gpNull.chistat_obs()            # == chistat(gpT, gpC, Xb), from that cliff face
    total = np.asarray(kw["Z"]).shape[1] if kw.get("Z") is not None else int(nsim)
    local, kw, reduce_ = _draw_shard(total, kw)
    cnt = 0                                   # a rank whose shard is empty (world > nsim) still joins the all-reduce
    if local > 0:
        _, cnt = nsim_chi(gpT, gpC, gpNull, Xb, local, chi_obs=chi_obs, return_count=True, **kw)
    gpNull.close()
    return reduce_(cnt) / total


def pval_invvar_uncalib(gpT: GPE, gpC: GPE, Xb, nugget: float = 1e-10) -> float:
    """pval_invvar_uncalib(gpT, gpC, Xb; nugget) observed (src/hypotest_invvar.jl:1-6)."""
    Xb = fmat(Xb)
    tau, var, p = C.c_double(), C.c_double(), C.c_double()
    gpT.ctx.check(gpT.ctx.lib.geordd_invvar_obs(gpT._h, gpC._h, dptr(Xb), Xb.shape[1], float(nugget), C.byref(tau),
                                                C.byref(var), C.byref(p)))
    return p.value


def nsim_invvar_pval_uncalib(gpT: GPE, gpC: GPE, Xb, nsim: int, nugget: float = 1e-10, *, seed: int = 0,
                             draw0: int = 0, Z=None, pval_obs: float = -1.0, return_count: bool = False,
                             gpNull: Optional[NullGPE] = None, sharded: bool = False):
    """nsim_invvar_pval_uncalib(gpT, gpC, Xb, nsim; nugget) (src/hypotest_invvar.jl:48-63)."""
    gpNull = gpNull or make_null(gpT, gpC)
    if sharded:
        total = np.asarray(Z).shape[1] if Z is not None else int(nsim)
        local, kw, gather = _vector_shard(total, dict(seed=seed, draw0=draw0, Z=Z))
        part = nsim_invvar_pval_uncalib(gpT, gpC, Xb, local, nugget, pval_obs=pval_obs, gpNull=gpNull, **kw) if local > 0 else np.empty(0)
        return gather(part)
    gpNull._ensure(Xb, nugget)
    Zf, nz = _zarg(Z, gpNull.nobs)
    nsim = nz if nz is not None else int(nsim)
    pv = np.empty(nsim)
    cnt = C.c_longlong()
    gpT.ctx.check(gpT.ctx.lib.geordd_null_invvar(gpNull._h, nsim, seed, draw0, dptr(Zf), float(pval_obs), dptr(pv),
                                                 C.byref(cnt)))
    return (pv, cnt.value) if return_count else pv


def boot_invvar(gpT: GPE, gpC: GPE, Xb, nsim: int, nugget: float = 1e-10, **kw) -> float:
    """boot_invvar (src/hypotest_invvar.jl:65-69): mean(pval_sims .< pval_obs)."""
    if isinstance(gpT, GroupGPE):
        return _group_boot("invvar", gpT, gpC, Xb, nsim, nugget, **kw)
    gpNull = NullGPE(gpT, gpC, Xb, nugget)
    pobs = gpNull.invvar_obs()[2]             # == pval_invvar_uncalib(gpT, gpC, Xb; nugget), from the handle's cliff face
    total = np.asarray(kw["Z"]).shape[1] if kw.get("Z") is not None else int(nsim)
    local, kw, reduce_ = _draw_shard(total, kw)
    cnt = 0
    if local > 0:
        _, cnt = nsim_invvar_pval_uncalib(gpT, gpC, Xb, local, nugget, pval_obs=pobs, return_count=True, gpNull=gpNull, **kw)
    gpNull.close()
    return reduce_(cnt) / total


def pval_invvar_calib(gpT: GPE, gpC: GPE, Xb, nugget: float = 1e-10) -> float:
    """pval_invvar_calib(gpT, gpC, Xb; nugget) analytic (src/hypotest_invvar.jl:74-105)."""
    Xb = fmat(Xb)
    p = C.c_double()
    gpT.ctx.check(gpT.ctx.lib.geordd_invvar_calib(gpT._h, gpC._h, dptr(Xb), Xb.shape[1], float(nugget), C.byref(p)))
    return p.value


def nsim_invvar_calib(gpT: GPE, gpC: GPE, Xb, nsim: int, nugget: float = 1e-10, *, seed: int = 0, draw0: int = 0, Z=None,
                      gpNull: Optional[NullGPE] = None) -> np.ndarray:
    """nsim_invvar_calib(gpT, gpC, Xb, nsim) (src/hypotest_invvar.jl:138-160): the analytically calibrated p-value
    (3-argument pval_invvar_calib: LU of Sigma_b + nugget) of every null draw; the draw-independent LU and covariance
    term are computed once instead of per draw, by the same solves."""
    gpNull = gpNull or make_null(gpT, gpC)
    gpNull._ensure(Xb, nugget)
    Zf, nz = _zarg(Z, gpNull.nobs)
    nsim = nz if nz is not None else int(nsim)
    pv = np.empty(nsim)
    gpT.ctx.check(gpT.ctx.lib.geordd_null_invvar_calib(gpNull._h, nsim, seed, draw0, dptr(Zf), dptr(pv)))
    return pv


def nsim_logP(gpT: GPE, gpC: GPE, gpNull: NullGPE, nsim: int, *, seed: int = 0, draw0: int = 0, Z=None,
              dmll_obs: float = math.inf, return_count: bool = False, sharded: bool = False):
    """nsim_logP(gpT, gpC, gpNull, nsim) (src/hypotest_mll.jl:21-30): list of (mll_null, mll_alt).
    Like the reference it leaves gpNull.y / .alpha / .mll at the last draw (:6,10,14) -- except in the opt-in
    forward-only mode (Context.set_mll_forward_only), which never forms alpha and leaves y / alpha untouched."""
    if sharded:   # every rank its shard of the draws; both vectors all-gathered (the tally is not meaningful here)
        total = np.asarray(Z).shape[1] if Z is not None else int(nsim)
        local, kw, gather = _vector_shard(total, dict(seed=seed, draw0=draw0, Z=Z))
        if local > 0:
            _, _, mn, ma = nsim_logP(gpT, gpC, gpNull, local, dmll_obs=dmll_obs, return_count=True, **kw)
        else:
            mn, ma = np.empty(0), np.empty(0)
        mn, ma = gather(mn), gather(ma)
        sims = list(zip(mn.tolist(), ma.tolist()))
        return (sims, int(np.sum((ma - mn) > dmll_obs)), mn, ma) if return_count else sims
    Zf, nz = _zarg(Z, gpNull.nobs)
    nsim = nz if nz is not None else int(nsim)
    mnull, malt = np.empty(nsim), np.empty(nsim)
    fwd_only = getattr(gpT.ctx, "_mll_forward_only", False)
    last_y, last_a = (None, None) if fwd_only else (np.empty(gpNull.nobs), np.empty(gpNull.nobs))
    cnt = C.c_longlong()
    gpT.ctx.check(gpT.ctx.lib.geordd_null_mll(gpNull._h, nsim, seed, draw0, dptr(Zf), float(dmll_obs), dptr(mnull),
                                              dptr(malt), C.byref(cnt), dptr(last_y), dptr(last_a)))
    if not fwd_only:
        gpNull.y, gpNull.alpha = last_y, last_a
    gpNull.mll = float(mnull[-1])
    sims = list(zip(mnull.tolist(), malt.tolist()))
    return (sims, cnt.value, mnull, malt) if return_count else sims


def boot_mlltest(gpT: GPE, gpC: GPE, nsim: int, **kw) -> float:
    """boot_mlltest(gpT, gpC, nsim) (src/hypotest_mll.jl:32-45): mean(dmll_sim .> dmll_obs)."""
    if isinstance(gpT, GroupGPE):
        return _group_boot("mll", gpT, gpC, None, nsim, **kw)
    mll_alt = gpT.mll + gpC.mll
    gpNull = make_null(gpT, gpC)
    d_obs = mll_alt - gpNull.mll
    total = np.asarray(kw["Z"]).shape[1] if kw.get("Z") is not None else int(nsim)
    local, kw, reduce_ = _draw_shard(total, kw)
    cnt = 0
    if local > 0:
        _, cnt, _, _ = nsim_logP(gpT, gpC, gpNull, local, dmll_obs=d_obs, return_count=True, **kw)
    return reduce_(cnt) / total


def nsim_power(gpT: GPE, gpC: GPE, tau: float, Xb, chi_null, mll_null, pval_invvar_null, nsim: int, *,
               seed: int = 0, draw0: int = 0, Z=None, compat_calib_bug: bool = False, sharded: bool = False,
               gpNull: Optional[NullGPE] = None) -> np.ndarray:
    """nsim_power (src/power.jl:46-73): nsim x 5 array of (pval_mll, pval_chi2, pval_invvar_obs,
    invvar_bootcalib, invvar_calib).  compat_calib_bug reproduces the as-shipped missing '+' of the
    7-argument pval_invvar_calib (src/hypotest_invvar.jl:125-128); default is the CORRECTED formula -- the one that
    reproduces the reference notebook's published size / power table (tests/golden/julia_table_replay.json); the as-shipped
    formula never rejects on that fixture.  sharded=True: the alternative draws are sharded over the ranks
    (torch.distributed) and the five columns all-gathered.  gpNull: reuse a prepared null handle (same sentinels, nugget 0)."""
    if sharded:
        total = np.asarray(Z).shape[1] if Z is not None else int(nsim)
        local, kw, gather = _vector_shard(total, dict(seed=seed, draw0=draw0, Z=Z))
        if local > 0:
            part = nsim_power(gpT, gpC, tau, Xb, chi_null, mll_null, pval_invvar_null, local, compat_calib_bug=compat_calib_bug,
                              gpNull=gpNull, **kw)
        else:
            part = np.empty((0, 5))
        return np.column_stack([gather(np.ascontiguousarray(part[:, j])) for j in range(5)])
    if gpNull is None:
        gpNull = NullGPE(gpT, gpC, Xb, 0.0)
    else:
        gpNull._ensure(Xb, 0.0)
    Zf, nz = _zarg(Z, gpNull.nobs)
    nsim = nz if nz is not None else int(nsim)
    chi_null = np.ascontiguousarray(chi_null, dtype=np.float64)
    mll_null = np.ascontiguousarray(mll_null, dtype=np.float64)
    pinv = np.ascontiguousarray(pval_invvar_null, dtype=np.float64)
    assert chi_null.shape == mll_null.shape == pinv.shape
    out = np.empty((5, nsim), order="F")
    gpT.ctx.check(gpT.ctx.lib.geordd_power(gpNull._h, float(tau), nsim, seed, draw0, dptr(Zf), dptr(chi_null),
                                           dptr(mll_null), dptr(pinv), chi_null.shape[0], 1 if compat_calib_bug else 0,
                                           dptr(out)))
    return np.ascontiguousarray(out.T)


# ----------------------------------------------------------------------------------------------
# Multi-GPU inside ONE process: the group layer of the C ABI (geordd_group_*).  The same driver names accept the
# group-fitted objects, so `boot_mlltest(gpT, gpC, nsim)` keeps its reference signature on 1 or 8 GPUs.
# ----------------------------------------------------------------------------------------------
class DeviceGroup:
    """geordd_group: several GPUs of one box driven by this process; fits replicated, null draws sharded by index, tallies
    summed with one NCCL all-reduce inside the library (no torch.distributed, no MPI)."""

    def __init__(self, devices: Sequence[int]):
        self.lib = capi.load_library()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._h = C.c_void_p()
        rc = self.lib.geordd_group_create(C.byref(self._h), len(devices), devs)
        if rc != 0:
            raise GeorddError(rc, "geordd_group_create failed (no usable sm_100 devices, or NCCL missing for > 1 device)")
        self.devices = list(devices)

    def check(self, rc: int) -> None:
        if rc == 0:
            return
        msg = self.lib.geordd_group_last_error(self._h).decode()
        if rc > 0:
            raise PosDefException(rc, msg)
        if rc == capi_ERR_ARG:
            raise ValueError("ArgumentError: " + msg)
        raise GeorddError(rc, msg)

    def context(self, i: int = 0) -> "Context":
        """Borrowed single-GPU context of device i (observed statistics, cliff faces, other models)."""
        ctx = Context.__new__(Context)
        ctx.lib, ctx.device = self.lib, self.devices[i]
        ctx._h = C.c_void_p(self.lib.geordd_group_ctx(self._h, i))
        ctx.close = lambda: None          # owned by the group
        return ctx

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.lib.geordd_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupGPE:
    """GPE(x, y, mean, kernel, logNoise) fitted on every device of a DeviceGroup (bit-identical replicas)."""

    def __init__(self, group: DeviceGroup, x, y, mean, kernel, logNoise: float):
        self.group = group
        self.x, self.y = fmat(x), np.array(y, dtype=np.float64)
        self.mean, self.kernel, self.logNoise = mean, kernel, float(logNoise)
        self.dim, self.nobs = self.x.shape
        spec = kernel_spec(kernel, mean)
        self._h = C.c_void_p()
        self.alpha = np.empty(self.nobs)
        mll, jr = C.c_double(), C.c_int()
        group.check(group.lib.geordd_group_gp_fit(group._h, C.byref(spec), dptr(self.x), self.dim, self.nobs, dptr(self.y),
                                                  self.logNoise, C.byref(self._h), dptr(self.alpha), C.byref(mll), C.byref(jr)))
        self.mll, self.jitter_rounds = mll.value, jr.value
        # device 0's replica as a plain (borrowed) GPE for the single-GPU entry points
        r = GPE.__new__(GPE)
        r.ctx, r.x, r.y, r.mean, r.kernel, r.logNoise = group.context(0), self.x, self.y, mean, kernel, self.logNoise
        r.dim, r.nobs, r.alpha, r.mll, r.jitter_rounds, r.cholU = self.dim, self.nobs, self.alpha, self.mll, self.jitter_rounds, None
        r._h = C.c_void_p(group.lib.geordd_group_gp_replica(self._h, 0))
        r.close = lambda: None
        self.rep0 = r

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.group.lib.geordd_group_gp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupNull:
    """make_null + the draw-independent setup on every device of the group (geordd_group_null_prepare)."""

    def __init__(self, gpT: GroupGPE, gpC: GroupGPE, Xb=None, nugget: float = 0.0):
        self.group, self.gpT, self.gpC = gpT.group, gpT, gpC
        self.Xb = fmat(Xb) if Xb is not None else None
        nb = self.Xb.shape[1] if self.Xb is not None else 0
        self._h = C.c_void_p()
        jr, m = C.c_int(), C.c_double()
        self.group.check(self.group.lib.geordd_group_null_prepare(gpT._h, gpC._h, dptr(self.Xb), nb, float(nugget),
                                                                  C.byref(self._h), C.byref(jr), C.byref(m)))
        self.jitter_rounds, self.mll, self.nugget = jr.value, m.value, nugget
        self.nobs = gpT.nobs + gpC.nobs

    def mll_draws(self, nsim: int, seed=0, draw0=0, Z=None, dmll_obs=math.inf):
        Zf, nz = _zarg(Z, self.nobs)
        nsim = nz if nz is not None else int(nsim)
        mn, ma, cnt = np.empty(nsim), np.empty(nsim), C.c_longlong()
        self.group.check(self.group.lib.geordd_group_null_mll(self._h, nsim, seed, draw0, dptr(Zf), float(dmll_obs), dptr(mn),
                                                              dptr(ma), C.byref(cnt)))
        return mn, ma, cnt.value

    def chi2_draws(self, nsim: int, seed=0, draw0=0, Z=None, chi_obs=math.inf):
        Zf, nz = _zarg(Z, self.nobs)
        nsim = nz if nz is not None else int(nsim)
        st, cnt = np.empty(nsim), C.c_longlong()
        self.group.check(self.group.lib.geordd_group_null_chi2(self._h, nsim, seed, draw0, dptr(Zf), float(chi_obs), dptr(st),
                                                               C.byref(cnt)))
        return st, cnt.value

    def invvar_draws(self, nsim: int, seed=0, draw0=0, Z=None, pval_obs=-1.0):
        Zf, nz = _zarg(Z, self.nobs)
        nsim = nz if nz is not None else int(nsim)
        pv, cnt = np.empty(nsim), C.c_longlong()
        self.group.check(self.group.lib.geordd_group_null_invvar(self._h, nsim, seed, draw0, dptr(Zf), float(pval_obs), dptr(pv),
                                                                 C.byref(cnt)))
        return pv, cnt.value

    def power_draws(self, tau, chi_null, mll_null, pinv_null, nsim: int, seed=0, draw0=0, Z=None, compat_calib_bug=False):
        Zf, nz = _zarg(Z, self.nobs)
        nsim = nz if nz is not None else int(nsim)
        a, b, c_ = (np.ascontiguousarray(v, dtype=np.float64) for v in (chi_null, mll_null, pinv_null))
        out = np.empty((5, nsim), order="F")
        self.group.check(self.group.lib.geordd_group_power(self._h, float(tau), nsim, seed, draw0, dptr(Zf), dptr(a), dptr(b),
                                                           dptr(c_), a.shape[0], 1 if compat_calib_bug else 0, dptr(out)))
        return np.ascontiguousarray(out.T)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.group.lib.geordd_group_null_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _group_boot(which: str, gpT: GroupGPE, gpC: GroupGPE, Xb, nsim: int, nugget: float = 0.0, **kw) -> float:
    """boot_chi2test / boot_mlltest / boot_invvar on a DeviceGroup: observed statistic on device 0, draws on all devices."""
    total = np.asarray(kw["Z"]).shape[1] if kw.get("Z") is not None else int(nsim)
    if which == "mll":
        gn = GroupNull(gpT, gpC)
        _, _, cnt = gn.mll_draws(total, dmll_obs=gpT.mll + gpC.mll - gn.mll, **kw)
    elif which == "chi":
        gn = GroupNull(gpT, gpC, Xb, 0.0)
        _, cnt = gn.chi2_draws(total, chi_obs=chistat(gpT.rep0, gpC.rep0, Xb), **kw)
    else:
        gn = GroupNull(gpT, gpC, Xb, nugget)
        _, cnt = gn.invvar_draws(total, pval_obs=pval_invvar_uncalib(gpT.rep0, gpC.rep0, Xb, nugget), **kw)
    gn.close()
    return cnt / total


# ----------------------------------------------------------------------------------------------
# RegionData, MultiGPCovars, GPRealisations
# ----------------------------------------------------------------------------------------------
@dataclass
class RegionData:
    """src/regiondata.jl:1-9 (shape omitted: geometry stays on the CPU side of the boundary)."""
    x: np.ndarray
    y: np.ndarray
    D: np.ndarray


def residuals_data(rd: RegionData, beta) -> RegionData:
    """residuals_data(rd, beta) (src/regiondata.jl:11-17): y - D'beta with empty covariates."""
    return RegionData(rd.x, rd.y - rd.D.T @ np.asarray(beta), np.empty((0, rd.y.shape[0])))


class MultiGPCovars:
    """MultiGPCovars(rdict, groupKeys, spkern, betakern, logNoise, mean) (src/region_gp_fit.jl:5-21 ->
    src/GPrealisationsCovars.jl:31-95): joint GP over all regions with a LinIso kernel on the covariates."""

    def __init__(self, rdict: Dict[Hashable, RegionData], spkern, betakern: LinIso, logNoise: float, mean=None,
                 groupKeys: Optional[Sequence[Hashable]] = None, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.groupKeys = list(groupKeys) if groupKeys is not None else list(rdict.keys())
        if not isinstance(betakern, LinIso):
            raise ValueError("ArgumentError: the covariate kernel must be LinIso")
        self.kernel, self.betakern, self.logNoise = spkern, betakern, float(logNoise)
        self.mean = mean if mean is not None else MeanZero()
        xs, ys, Ds, off = [], [], [], [0]
        self.groupIndices = {}
        for key in self.groupKeys:
            rd = rdict[key]
            if rd.y.shape[0] == 0:
                raise ValueError("empty group")
            self.groupIndices[key] = range(off[-1], off[-1] + rd.y.shape[0])
            xs.append(fmat(rd.x)); ys.append(np.asarray(rd.y, dtype=np.float64)); Ds.append(fmat(rd.D))
            off.append(off[-1] + rd.y.shape[0])
        self.x = fmat(np.concatenate(xs, axis=1))
        self.y = np.concatenate(ys)
        self.D = fmat(np.concatenate(Ds, axis=1))
        self.p, self.nobs = self.D.shape
        self.dim = self.x.shape[0]
        if self.D.shape[1] != self.y.shape[0]:
            raise ValueError("incompatible dimensions of covariates matrix and gaussian processes")
        self._off = np.array(off, dtype=np.int32)
        self._h = C.c_void_p()
        rc = self.ctx.lib.geordd_multigp_create(self.ctx._h, dptr(self.D), self.p, dptr(self.x), self.dim,
                                                self._off.ctypes.data_as(capi.c_int_p), len(self.groupKeys), dptr(self.y),
                                                C.byref(self._h))
        self.ctx.check(rc)
        self.alpha = np.empty(self.nobs)
        self.mll = float("nan")
        self.dmll = None
        self.jitter_rounds = 0
        update_mll(self)      # initialise_mll! (:127-142)

    def _spec(self):
        return kernel_spec(self.kernel, self.mean)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.ctx.lib.geordd_multigp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def update_mll(obj):
    """update_mll!(mgpcv) (src/GPrealisationsCovars.jl:110-125) / update_mll!(gpreals) (src/GPrealisations.jl:101-109)."""
    if isinstance(obj, GPRealisations):
        obj.mll = float(sum(gp.mll for gp in obj.mgp.values()))
        return obj.mll
    spec = obj._spec()
    m, jr = C.c_double(), C.c_int()
    obj.ctx.check(obj.ctx.lib.geordd_multigp_mll(obj._h, C.byref(spec), obj.betakern.ell2, obj.logNoise, dptr(obj.alpha),
                                                 C.byref(m), C.byref(jr)))
    obj.mll, obj.jitter_rounds = m.value, jr.value
    return obj.mll


def update_mll_and_dmll(mgpcv: MultiGPCovars, noise=True, domean=True, kern=True, beta=True) -> np.ndarray:
    """update_mll_and_dmll!(mgpcv, buf; noise, domean, kern, beta) (src/GPrealisationsCovars.jl:145-193)."""
    spec = mgpcv._spec()
    m, nd = C.c_double(), C.c_int()
    buf = np.empty(8)
    nfix = 0 if not isinstance(mgpcv.kernel, (FixedKernel,)) else 0
    mgpcv.ctx.check(mgpcv.ctx.lib.geordd_multigp_mll_grad(mgpcv._h, C.byref(spec), mgpcv.betakern.ell2, mgpcv.logNoise,
                                                          int(noise), int(domean), int(kern), int(beta), C.byref(m),
                                                          dptr(buf), C.byref(nd)))
    mgpcv.mll = m.value
    g = buf[: nd.value].copy()
    # hide fixed sub-kernel parameters (fix(Const)) exactly like get_params does
    if kern:
        g = _drop_fixed_kernel_grads(mgpcv, g, noise, domean, beta)
    mgpcv.dmll = g
    return g


def _kernel_param_mask(kernel) -> List[bool]:
    """True for each device-reported kernel gradient entry that get_params exposes."""
    k = kernel
    if isinstance(k, FixedKernel):
        inner = _kernel_param_mask(k.kernel)
        return [False] * len(inner)
    if isinstance(k, SumKernel):
        a, b = k.kleft, k.kright
        ma, mb = _kernel_param_mask(a), _kernel_param_mask(b)
        # device order is always [spatial..., const]
        if isinstance(_unfix(a), Const):
            return mb + ma
        return ma + mb
    if isinstance(k, Const):
        return [True]
    return [True, True]


def _drop_fixed_kernel_grads(mgpcv, g, noise, domean, beta):
    start = (1 if noise else 0) + (1 if (domean and isinstance(mgpcv.mean, MeanConst) and mgpcv.mean.beta != 0.0) else 0)
    mask = _kernel_param_mask(mgpcv.kernel)
    nk = len(mask)
    kg = [g[start + i] for i in range(nk) if mask[i]]
    # reorder to get_params order if Const is on the left
    k = mgpcv.kernel
    if isinstance(k, SumKernel) and isinstance(_unfix(k.kleft), Const) and len(kg) == 3:
        kg = [kg[2], kg[0], kg[1]]
    return np.concatenate([g[:start], kg, g[start + nk:]])


def mgp_get_params(mgpcv: MultiGPCovars, noise=True, domean=True, kern=True, beta=True) -> np.ndarray:
    """get_params(mgpcv; ...) (src/GPrealisationsCovars.jl:194-201)."""
    p: List[float] = []
    if noise:
        p.append(mgpcv.logNoise)
    if domean and isinstance(mgpcv.mean, MeanConst):
        p.append(mgpcv.mean.beta)
    if kern:
        p.extend(get_params(mgpcv.kernel))
    if beta:
        p.extend(get_params(mgpcv.betakern))
    return np.array(p)


def mgp_set_params(mgpcv: MultiGPCovars, hyp, noise=True, domean=True, kern=True, beta=True) -> None:
    """set_params!(mgpcv, hyp; ...) (src/GPrealisationsCovars.jl:202-222)."""
    i = 0
    if noise:
        mgpcv.logNoise = float(hyp[i]); i += 1
    if domean and isinstance(mgpcv.mean, MeanConst):
        mgpcv.mean.beta = float(hyp[i]); i += 1
    if kern:
        nk = len(get_params(mgpcv.kernel))
        set_params(mgpcv.kernel, hyp[i:i + nk]); i += nk
    if beta:
        set_params(mgpcv.betakern, hyp[i:i + 1]); i += 1


def optimize(mgpcv: MultiGPCovars, noise=True, domean=True, kern=True, beta=True, method="CG", options=None):
    """optimize!(mgpcv; ...) (src/GPrealisationsCovars.jl:224-279).  The optimiser loop stays on the host
    (Optim.jl in Julia; SciPy's nonlinear CG here); every objective / gradient evaluation is one ABI call.
    Non-finite parameters, ArgumentError and PosDefException evaluate to +Inf exactly as :232-245."""
    from scipy.optimize import minimize
    kw = dict(noise=noise, domean=domean, kern=kern, beta=beta)

    def fg(hyp):
        try:
            if not np.all(np.isfinite(hyp)):
                raise ValueError("non-finite hyper-parameters")
            mgp_set_params(mgpcv, hyp, **kw)
            g = update_mll_and_dmll(mgpcv, **kw)
            return -mgpcv.mll, -g
        except (ValueError, PosDefException, OverflowError) as err:
            print(err)
            return math.inf, np.zeros_like(hyp)

    init = mgp_get_params(mgpcv, **kw)
    res = minimize(fg, init, jac=True, method=method, options=options or {})
    mgp_set_params(mgpcv, res.x, **kw)
    update_mll(mgpcv)
    return res


def postmean_beta(mgpcv: MultiGPCovars) -> np.ndarray:
    """postmean_beta(mgpcv) (src/GPrealisationsCovars.jl:335-344)."""
    spec = mgpcv._spec()
    out = np.empty(mgpcv.p)
    mgpcv.ctx.check(mgpcv.ctx.lib.geordd_multigp_postmean_beta(mgpcv._h, C.byref(spec), mgpcv.betakern.ell2,
                                                               mgpcv.logNoise, dptr(out)))
    return out


class GPRealisations:
    """GPRealisations(rdict, groupKeys, spkern, logNoise, mean) (src/region_gp_fit.jl:23-37 ->
    src/GPrealisations.jl:10-55): one fitted GPE per region sharing (kernel, logNoise).  'Realisations'
    are regions, not random draws."""

    def __init__(self, rdict: Dict[Hashable, RegionData], spkern, logNoise: float, mean=None,
                 groupKeys: Optional[Sequence[Hashable]] = None, ctx: Optional[Context] = None):
        self.groupKeys = list(groupKeys) if groupKeys is not None else list(rdict.keys())
        self.kernel, self.logNoise = spkern, float(logNoise)
        self.mean = mean if mean is not None else MeanZero()
        for key in self.groupKeys:
            if rdict[key].y.shape[0] == 0:
                raise ValueError("empty group")
        # the reference fits the regions one after another (src/GPrealisations.jl:13-17); here all of them in one call
        fits = GPE.fit_many([rdict[k].x for k in self.groupKeys], [rdict[k].y for k in self.groupKeys], self.mean, spkern,
                            logNoise, ctx=ctx)
        self.mgp: Dict[Hashable, GPE] = dict(zip(self.groupKeys, fits))
        self.nobs = sum(g.nobs for g in fits)
        self.mll = 0.0
        update_mll(self)

    def __getitem__(self, key) -> GPE:
        return self.mgp[key]

    def keys(self):
        return self.groupKeys

    def values(self):
        return self.mgp.values()


# ----------------------------------------------------------------------------------------------
# Placebo drivers (src/placebo_geometry.jl + src/hypotest_*.jl placebo_*): O(n) host geometry feeding the
# GPU tests.  Thin wrappers, SURVEY §8 a15.
# ----------------------------------------------------------------------------------------------
def _sind(x: float) -> float:
    """Base.sind: argument reduction in DEGREES, so that multiples of 90 are exact (sind(180) == 0)."""
    rx = math.copysign(math.fmod(x, 360.0), x)
    arx = abs(rx)
    if rx == 0.0:
        return rx
    if arx < 45.0:
        return math.sin(math.radians(rx))
    if arx <= 135.0:
        return math.copysign(math.cos(math.radians(90.0 - arx)), rx)
    if arx == 180.0:
        return math.copysign(0.0, rx)
    if arx < 225.0:
        return math.sin(math.radians((180.0 - arx) * math.copysign(1.0, rx)))
    if arx <= 315.0:
        return -math.copysign(math.cos(math.radians(270.0 - arx)), rx)
    return math.sin(math.radians(rx - math.copysign(360.0, rx)))


def _cosd(x: float) -> float:
    """Base.cosd: cosd(90) == cosd(270) == 0 exactly."""
    rx = abs(math.fmod(x, 360.0))
    if rx <= 45.0:
        return math.cos(math.radians(rx))
    if rx < 135.0:
        return math.sin(math.radians(90.0 - rx))
    if rx <= 225.0:
        return -math.cos(math.radians(180.0 - rx))
    if rx < 315.0:
        return math.sin(math.radians(rx - 270.0))
    return math.cos(math.radians(360.0 - rx))


def _tand(x: float) -> float:
    """Base.tand = sind / cosd: tand(45) == 1.0 exactly (the `abs(dydx) < 1` branch of left_points depends on it),
    tand(90) == Inf."""
    s, c = _sind(x), _cosd(x)
    if c == 0.0:
        return math.copysign(math.inf, s) if s != 0.0 else math.nan
    return s / c


def _placebo_direction(angle: float) -> float:
    d = _cosd(angle + 90.0)
    return 1.0 if d == 0.0 else math.copysign(1.0, d)


def left_points(angle: float, shift: float, X) -> np.ndarray:
    """left_points(angle, shift, X) (src/placebo_geometry.jl:8-32)."""
    X = np.asarray(X)
    dydx = _tand(angle)
    meanx, meany = float(np.mean(X[0])), float(np.mean(X[1]))
    direction = _placebo_direction(angle)
    sx = shift * _cosd(angle + 90.0) * direction
    sy = shift * _sind(angle + 90.0) * direction
    ax, ay = meanx + sx, meany + sy
    if abs(dydx) < 1.0:
        dx, dy = 1.0, dydx
    else:
        dx, dy = 1.0 / dydx, 1.0
    bx, by = meanx + sx + dx, meany + dy + sy
    return ((bx - ax) * (X[1] - ay) - (by - ay) * (X[0] - ax)) > 0


def shift_for_even_split(angle: float, X, tol: float = 1e-5, maxiter: int = 100) -> float:
    """shift_for_even_split(angle, X) (src/placebo_geometry.jl:131-136): bisection (tol 1e-5) on
    prop_left(angle, shift, X) - 0.5 over [-maxdistance, maxdistance] (:112-129)."""
    X = np.asarray(X)
    maxd = float(math.hypot(np.ptp(X[1]), np.ptp(X[0])))
    f = lambda sh: float(np.mean(left_points(angle, sh, X))) - 0.5
    a, b = -maxd, maxd
    fa = f(a)
    if fa * f(b) > 0:
        raise ValueError("No real root in [a,b]")
    c = 0.5 * (a + b)
    for it in range(1, maxiter + 1):
        if b - a <= tol:
            break
        if it == maxiter:
            raise RuntimeError("Max iteration exceeded")
        c = 0.5 * (a + b)
        fc = f(c)
        if fc == 0:
            break
        if fa * fc > 0:
            a, fa = c, fc
        else:
            b = c
    return c


def placebo_sentinels(angle: float, shift: float, X, npoints: int) -> np.ndarray:
    """placebo_sentinels(angle, shift, X, npoints) (src/placebo_geometry.jl:60-101): npoints equally spaced
    between the two intersections of the placebo line with the data's bounding box."""
    X = np.asarray(X)
    meanx, meany = float(np.mean(X[0])), float(np.mean(X[1]))
    minx, maxx = float(X[0].min()), float(X[0].max())
    miny, maxy = float(X[1].min()), float(X[1].max())
    dydx = _tand(angle)
    direction = _placebo_direction(angle)
    cx = meanx + shift * _cosd(angle + 90.0) * direction
    cy = meany + shift * _sind(angle + 90.0) * direction
    ext = []
    y0, y1 = cy + dydx * (minx - cx), cy + dydx * (maxx - cx)
    if miny < y0 < maxy:
        ext.append((minx, y0))
    if miny < y1 < maxy:
        ext.append((maxx, y1))
    if dydx != 0.0:                       # the reference divides unconditionally: +-Inf / NaN fail both comparisons
        x0, x1 = cx + (miny - cy) / dydx, cx + (maxy - cy) / dydx
        if minx < x0 < maxx:
            ext.append((x0, miny))
        if minx < x1 < maxx:
            ext.append((x1, maxy))
    assert len(ext) == 2, "placebo line must cross the bounding box twice"
    return fmat(np.vstack([np.linspace(ext[0][0], ext[1][0], npoints), np.linspace(ext[0][1], ext[1][1], npoints)]))


def _placebo_split(angle, X, Y, kern, mean, logNoise, ctx=None):
    X = fmat(X)
    Y = np.asarray(Y, dtype=np.float64)
    shift = shift_for_even_split(angle, X)
    left = left_points(angle, shift, X)
    gl, gr = GPE.fit_many([X[:, left], X[:, ~left]], [Y[left], Y[~left]], mean, kern, logNoise, ctx=ctx)
    return shift, gl, gr


def placebo_chi(angle: float, X, Y, kern, mean, logNoise: float, nsim: int, ctx=None, **kw) -> float:
    """placebo_chi (src/hypotest_chi2.jl:62-76)."""
    shift, gl, gr = _placebo_split(angle, X, Y, kern, mean, logNoise, ctx)
    Xb = placebo_sentinels(angle, shift, X, 100)
    return boot_chi2test(gl, gr, Xb, nsim, **kw)


def placebo_invvar(angle: float, X, Y, kern, mean, logNoise: float, ctx=None) -> float:
    """placebo_invvar (src/hypotest_invvar.jl:162-175)."""
    shift, gl, gr = _placebo_split(angle, X, Y, kern, mean, logNoise, ctx)
    Xb = placebo_sentinels(angle, shift, X, 100)
    return pval_invvar_calib(gl, gr, Xb)


def placebo_mll(angle: float, X, Y, kern, mean, logNoise: float, nsim: int, ctx=None, **kw) -> float:
    """placebo_mll (src/hypotest_mll.jl:47-60)."""
    _, gl, gr = _placebo_split(angle, X, Y, kern, mean, logNoise, ctx)
    return boot_mlltest(gl, gr, nsim, **kw)


def placebo_sweep(which: str, angles, X, Y, kern, mean, logNoise: float, nsim: int = 0, workers: int = 1, **kw) -> np.ndarray:
    """The placebo sweep of the notebooks (`[placebo_chi(angle, ...) for angle in 1:2:180]`,
    notebooks/NYC_analysis-2015.ipynb:1452-1555): one p-value per angle.
      * Under torch.distributed the angles are dealt to the ranks (parallel.map_sharded), one all-gather of the results.
      * workers > 1: the angles of this process are additionally spread over `workers` host threads, each with its OWN
        context (stream, scheduling scratch, memory pool) on the same GPU.  A half-district fit uses a few SMs, so several
        independent (fit + test) problems then run side by side on the chip; the ABI calls release the GIL.
        Results do not depend on `workers` (every problem is computed by the same kernels)."""
    from . import parallel
    import threading
    from concurrent.futures import ThreadPoolExecutor
    if which not in ("chi", "mll", "invvar"):
        raise ValueError("which must be 'chi', 'mll' or 'invvar'")
    base = default_context()
    local = threading.local()

    def one(a):
        ctx = getattr(local, "ctx", None)
        if ctx is None:
            ctx = local.ctx = base if workers <= 1 else Context(base.device)
        if which == "chi":
            return placebo_chi(a, X, Y, kern, mean, logNoise, nsim, ctx=ctx, **kw)
        if which == "mll":
            return placebo_mll(a, X, Y, kern, mean, logNoise, nsim, ctx=ctx, **kw)
        return placebo_invvar(a, X, Y, kern, mean, logNoise, ctx=ctx)

    if workers <= 1:
        return parallel.map_sharded(one, list(angles))
    with ThreadPoolExecutor(max_workers=int(workers)) as pool:
        return parallel.map_sharded(one, list(angles), mapper=lambda f, items: list(pool.map(f, items)))


def adjacent_sweep(gps, borders: Dict, nugget: float = 1e-10):
    """The loop over `adjacent_pairs(rd_dict, buffer)` of the notebooks (src/regiondata.jl:64-68,
    notebooks/NYC_analysis-2015.ipynb cell 74): for every pair of regions that share a border -- `borders` maps
    (key_i, key_j) to that border's 2 x nb sentinels; finding the borders is LibGEOS work and stays with the caller --
    the cliff face, the inverse-variance LATE and the calibrated p-value, from the already fitted per-region GPs
    (`gps`: GPRealisations or a dict key -> GPE).  Returns (tau_post_pairs, pval_pairs) filled in both orientations like the
    notebook's dictionaries.  Under torch.distributed the pairs are dealt to the ranks (every rank holds all fits; one
    all-gather of three numbers per pair)."""
    from . import parallel
    pairs = list(borders.keys())
    done: Dict[int, tuple] = {}

    def pair_results(p):
        if p not in done:
            ki, kj = pairs[p]
            a, b = gps[ki], gps[kj]
            Xb = fmat(borders[(ki, kj)])
            mu, Sigma = cliff_face(a, b, Xb)
            tp = inverse_variance(mu, Sigma, nugget, ctx=a.ctx)
            done[p] = (tp.mu, tp.sigma, pval_invvar_calib(a, b, Xb, nugget=nugget))
        return done[p]

    # every call deals the same pairs to the same rank, so each pair is evaluated once; one all-gather per output
    cols = [parallel.map_sharded(lambda p, c=c: pair_results(p)[c], range(len(pairs))) for c in range(3)]
    tau_post, pvals = {}, {}
    for p, (ki, kj) in enumerate(pairs):
        m, s, pv = (float(col[p]) for col in cols)
        tau_post[(ki, kj)] = Normal(m, s)
        tau_post[(kj, ki)] = Normal(-m, s)
        pvals[(ki, kj)] = pvals[(kj, ki)] = pv
    return tau_post, pvals
