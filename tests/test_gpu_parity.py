"""GPU parity tests: the CUDA path, called through the C ABI (via the host mirror geordd.jl_b200/geordd.py),
against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (BASELINE north_star): posterior means / covariances / alpha / mll within 1e-8 RELATIVE
(normwise, SURVEY §7 hard part 1); null statistics within 1e-8 given identical standard-normal draws;
exceedance counts and p-values bit-exact.  Bit-exact counts are certified by the minimum relative margin
min_s |stat_s - obs| / |obs| being far above the 1e-8 agreement level (SURVEY §7 hard part 2).
"""
import math
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _kern(G, kind, ell, sigma2, const_s2):
    cls = {0: G.SEIso, 1: G.Mat12Iso, 2: G.Mat32Iso, 3: G.Mat52Iso}[kind]
    k = cls(math.log(ell), 0.5 * math.log(sigma2))
    if const_s2 > 0:
        k = k + G.Const(0.5 * math.log(const_s2))
    return k


def _pair(G, O, m=150, nb=40, kind=2, seed=1, ragged=0, const_s2=400.0, mean=(0.0, 0.0), ln=math.log(0.3)):
    d = O.synth_two_region(m, nb, seed=seed)
    mT = m + ragged
    XT, XC, YT, YC = d["X"][:, :mT], d["X"][:, mT:], d["Y"][:mT], d["Y"][mT:]
    spT = O.KernelSpec(kind, 0.3, 1.0, const_s2, mean[0])
    spC = O.KernelSpec(kind, 0.3, 1.0, const_s2, mean[1])
    oT, oC = O.gp_fit(spT, XT, YT, ln), O.gp_fit(spC, XC, YC, ln)
    k = _kern(G, kind, 0.3, 1.0, const_s2)
    gT = G.GPE(XT, YT, G.MeanConst(mean[0]) if mean[0] else G.MeanZero(), k, ln)
    gC = G.GPE(XC, YC, G.MeanConst(mean[1]) if mean[1] else G.MeanZero(), k, ln)
    return d, (oT, oC), (gT, gC)


# ---------------------------------------------------------------------------------------------
# K1 covariance assembly
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 63, 64, 129, 300])
def test_cov_sym_and_cross(G, O, ctx, kind, n):
    rng = np.random.default_rng(n + kind)
    X = rng.random((2, n))
    X2 = rng.random((2, 77))
    spec = O.KernelSpec(kind, 0.41, 1.7, 0.3)
    k = _kern(G, kind, 0.41, 1.7, 0.3)
    K = G.cov(k, X, diag_add=0.09)
    Ko = O.cov_sym(spec, X, 0.09)
    assert np.max(np.abs(K - Ko)) < 1e-13 * np.max(np.abs(Ko))
    assert np.array_equal(K, K.T)                          # exactly symmetric
    Kc = G.cov(k, X, X2)
    assert np.max(np.abs(Kc - O.cov_cross(spec, X, X2))) < 1e-13 * 2.0


def test_cov_3d_coordinates(G, O, ctx):
    rng = np.random.default_rng(0)
    X = rng.random((3, 50))
    K = G.cov(G.SEIso(math.log(0.5), 0.0), X)
    assert relerr(K, O.cov_sym(O.KernelSpec(0, 0.5, 1.0), X)) < 1e-14


# ---------------------------------------------------------------------------------------------
# a1 / a9: GPE fit, update_alpha!, mll
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 128, 129, 700])
@pytest.mark.parametrize("kind", [0, 2])
def test_gp_fit_matches_oracle(G, O, ctx, n, kind):
    rng = np.random.default_rng(n)
    X = rng.random((2, n))
    y = np.sin(3 * X[0]) + rng.standard_normal(n) * 0.3
    spec = O.KernelSpec(kind, 0.3, 1.0, 4.0, 0.7)
    o = O.gp_fit(spec, X, y, math.log(0.3))
    g = G.GPE(X, y, G.MeanConst(0.7), _kern(G, kind, 0.3, 1.0, 4.0), math.log(0.3), want_chol=True)
    assert g.jitter_rounds == o.jitter_rounds == 0
    assert relerr(g.alpha, o.alpha) < TOL
    assert abs(g.mll - o.mll) < TOL * abs(o.mll)
    assert relerr(np.triu(g.cholU), np.triu(o.U)) < 1e-10          # upper factor, U'U = K
    assert relerr(g.cK_mat(), o.K) < 1e-13
    # update_alpha! with a new y and the frozen factor (src/gp_utils.jl:15-22)
    y2 = rng.standard_normal(n)
    g.y[:] = y2
    G.update_alpha(g)
    o.y = y2
    O.update_alpha(o)
    assert relerr(g.alpha, o.alpha) < TOL and abs(G.mll(g) - O.gp_mll(o)) < TOL * abs(O.gp_mll(o))
    mu = G.predict_mu(g, X[:, :5])
    assert relerr(mu, O.predict_mu(o, X[:, :5], O.cov_cross(spec, X[:, :5], X))) < TOL


def test_batched_fits_equal_single_fits_bit_for_bit(G, O, ctx):
    """geordd_gp_fit_batch (GPRealisations, placebo halves): one merged dataflow Cholesky launch, per-GP tails on side
    streams -- alpha, mll, log-determinant identical to one geordd_gp_fit per region, incl. one-tile and ragged sizes;
    a batch containing a matrix that needs make_posdef! jitter falls back to the per-matrix protocol."""
    rng = np.random.default_rng(11)
    sizes = [100, 129, 300, 513, 128, 700, 64]
    xs = [np.asfortranarray(rng.random((2, n))) for n in sizes]
    ys = [rng.standard_normal(n) for n in sizes]
    k = G.Mat52Iso(math.log(0.25), 0.1) + G.Const(math.log(3.0))
    ln = math.log(0.2)
    batch = G.GPE.fit_many(xs, ys, G.MeanConst(0.3), k, ln)
    for g, x, y in zip(batch, xs, ys):
        one = G.GPE(x, y, G.MeanConst(0.3), k, ln)
        assert g.mll == one.mll and np.array_equal(g.alpha, one.alpha) and g.jitter_rounds == one.jitter_rounds == 0
        oo = O.gp_fit(O.KernelSpec(3, 0.25, math.exp(0.2), 9.0, 0.3), x, y, ln)
        assert abs(g.mll - oo.mll) < TOL * abs(oo.mll) and relerr(g.alpha, oo.alpha) < TOL
    # duplicate units make the second matrix numerically singular without noise: jitter rounds, identical to the single fit
    xd = np.asfortranarray(np.hstack([xs[1], xs[1][:, :40]]))
    yd = rng.standard_normal(xd.shape[1])
    kk = G.SEIso(math.log(0.3), 0.0)
    pair = G.GPE.fit_many([xs[2], xd], [ys[2], yd], G.MeanZero(), kk, -30.0)
    ref = G.GPE(xd, yd, G.MeanZero(), kk, -30.0)
    assert pair[1].jitter_rounds == ref.jitter_rounds >= 1 and pair[1].mll == ref.mll and np.array_equal(pair[1].alpha, ref.alpha)
    assert pair[0].mll == G.GPE(xs[2], ys[2], G.MeanZero(), kk, -30.0).mll


def test_make_posdef_jitter_and_failure(G, O, ctx):
    """Duplicate points + tiny noise: Cholesky needs the jitter loop; rounds and results must match the
    oracle's restatement of make_posdef! (App. A.5).  A hopeless matrix raises PosDefException(info)."""
    rng = np.random.default_rng(3)
    X = rng.random((2, 40))
    X = np.concatenate([X, X, X], axis=1)                  # rank-deficient kernel matrix
    y = rng.standard_normal(120)
    # Const(1e4) makes the rounding noise of the Schur complements (~1e-12) swamp the eps-level noise term, so the
    # un-jittered factorisation fails for certain in BOTH implementations; one jitter round (1e-6 tr/n ~ 1e-2) fixes it.
    spec = O.KernelSpec(0, 0.8, 1.0, 1e4)
    ln = -40.0                                             # noise ~ 1e-35 + eps
    o = O.gp_fit(spec, X, y, ln)
    g = G.GPE(X, y, G.MeanZero(), G.SEIso(math.log(0.8), 0.0) + G.Const(math.log(100.0)), ln)
    assert o.jitter_rounds == 1
    assert g.jitter_rounds == o.jitter_rounds
    assert relerr(np.diag(g.cK_mat()), np.diag(o.K)) < 1e-14   # same eps = 1e-6 tr(M)/n on the diagonal
    assert relerr(g.alpha, o.alpha) < 1e-4 and abs(g.mll - o.mll) < 1e-6 * abs(o.mll)
    with pytest.raises(ValueError):
        G.GPE(X, y, G.MeanZero(), G.LinIso(0.0), 0.0)          # ArgumentError: unsupported kernel


def test_abi_argument_errors_and_edge_sizes(G, O, ctx):
    """Error behaviour at the ABI (rc -1 -> ArgumentError): empty inputs, zero draws, missing sentinels, NULL handles,
    incompatible GP pairs, bad borders; and the smallest legal sizes (one unit, one sentinel, one draw)."""
    import ctypes as C
    lib = ctx.lib
    k = G.SEIso(math.log(0.3), 0.0)
    rng = np.random.default_rng(0)
    X, y = rng.random((2, 5)), rng.standard_normal(5)
    with pytest.raises(ValueError):
        G.GPE(np.zeros((2, 0)), np.zeros(0), G.MeanZero(), k, 0.0)              # no units
    assert lib.geordd_cliff_face(None, None, None, 0, None, None) == -1          # NULL handles never dereferenced
    assert lib.geordd_null_mll(None, 10, 0, 0, None, 0.0, None, None, None, None, None) == -1
    g1 = G.GPE(X[:, :1], y[:1], G.MeanZero(), k, 0.0)                            # a single unit
    g4 = G.GPE(X[:, 1:], y[1:], G.MeanZero(), k, 0.0)
    assert abs(g1.mll - O.gp_fit(O.KernelSpec(0, 0.3, 1.0, 0.0), X[:, :1], y[:1], 0.0).mll) < 1e-12
    mu, S = G.cliff_face(g1, g4, X[:, :1])                                       # one sentinel
    assert mu.shape == (1,) and S.shape == (1, 1) and S[0, 0] > 0
    gN = G.make_null(g1, g4)
    with pytest.raises(ValueError):
        G.nsim_logP(g1, g4, gN, 0)                                               # zero draws
    cnt = C.c_longlong()
    assert lib.geordd_null_chi2(gN._h, 5, 0, 0, None, 0.0, None, C.byref(cnt)) == -1   # no sentinels attached
    one = G.nsim_chi(g1, g4, gN, X[:, :2], 1, seed=1)                            # one draw, two sentinels
    assert one.shape == (1,) and np.isfinite(one[0]) and one[0] >= 0
    g3d = G.GPE(rng.random((3, 4)), y[:4], G.MeanZero(), k, 0.0)
    with pytest.raises(ValueError):
        G.make_null(g1, g3d)                                                     # 2-d and 3-d GPs do not pool
    with pytest.raises(ValueError):
        G.projection_points_all(X, np.array([[0.0, 0.0]]), ctx=ctx)              # a border needs a segment
    with pytest.raises(ValueError):
        G.projection_points_all(X, np.array([[0.0, 0.0], [1.0, 1.0]]), part_start=(1,), ctx=ctx)   # ... a real one
    assert "segment" in lib.geordd_last_error(ctx._h).decode()


def test_plain_c_consumer_matches_python_path(G, O, ctx, tmp_path):
    """examples/c_driver.c (gcc -std=c99, links only libgeordd_b200.so) reproduces the Python path's numbers."""
    import subprocess
    from test_abi_and_host import _build_c_driver
    exe = _build_c_driver(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.array([float(v) for v in r.stdout.split()])
    M, s = 60, 12345
    XT, XC = np.zeros((2, M), order="F"), np.zeros((2, M), order="F")
    def lcg():
        nonlocal s
        s = (s * 1664525 + 1013904223) % 2 ** 32
        return (s >> 8) / 16777216.0
    for i in range(M):
        XT[0, i] = 0.5 * lcg(); XT[1, i] = lcg(); XC[0, i] = 0.5 + 0.5 * lcg(); XC[1, i] = lcg()
    YT = np.sin(3.0 * XT[0]) + XT[1] + 0.4
    YC = np.sin(3.0 * XC[0]) + XC[1]
    Xb = np.asfortranarray(np.array([[0.5, 0.5, 0.5], [0.25, 0.5, 0.75]]))
    k = G.SEIso(math.log(0.3), 0.0)
    gT, gC = G.GPE(XT, YT, G.MeanZero(), k, math.log(0.1)), G.GPE(XC, YC, G.MeanZero(), k, math.log(0.1))
    mu, S = G.cliff_face(gT, gC, Xb)
    chi = G.chistat(gT, gC, Xb)
    p = G.boot_chi2test(gT, gC, Xb, 2000, seed=1)
    want = np.array([mu[0], mu[1], mu[2], S[0, 0], chi, p])
    assert np.allclose(got, want, rtol=1e-12, atol=0) and got[5] == want[5]


def test_prior_rand(G, O, ctx):
    d, (oT, oC), (gT, gC) = _pair(G, O, m=90, nb=10, mean=(0.4, 0.0))
    Z = np.random.default_rng(1).standard_normal((90, 5))
    Y = G.prior_rand(gT, Z=Z)
    Yo = np.stack([O.prior_rand(oT, Z[:, s]) for s in range(5)], axis=1)
    assert relerr(Y, Yo) < 1e-12


# ---------------------------------------------------------------------------------------------
# a7: cliff face; LATE reducers; observed statistics
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,nb", [(2, 1), (2, 40), (2, 130), (3, 64), (1, 100), (0, 12)])
def test_cliff_face(G, O, ctx, kind, nb):
    d, (oT, oC), (gT, gC) = _pair(G, O, m=200, nb=nb, kind=kind, ragged=17, mean=(0.3, -0.2))
    mu, S = G.cliff_face(gT, gC, d["Xb"])
    mo, So = O.cliff_face(oT, oC, d["Xb"])
    assert relerr(mu, mo) < TOL
    assert relerr(S, So) < TOL
    assert np.array_equal(S, S.T)


def test_observed_statistics_and_late(G, O, ctx):
    """Observed statistics = LU solves against the RAW cliff covariance (src/hypotest_chi2.jl:9, src/hypotest_invvar.jl:2-3,
    98-101).  Two correct LU implementations of a system with condition number kappa agree to ~kappa * eps, so the gate is
    max(1e-8, 50 kappa eps) with kappa measured -- 1e-8 itself whenever the problem allows it."""
    d, (oT, oC), (gT, gC) = _pair(G, O, m=200, nb=60, kind=2)
    Xb = d["Xb"]
    kappa = float(np.linalg.cond(O.cliff_face(oT, oC, Xb)[1]))
    tol = max(TOL, 50.0 * kappa * np.finfo(float).eps)
    assert tol < 1e-6, kappa
    assert abs(G.chistat(gT, gC, Xb) - O.chistat_obs(oT, oC, Xb)) < tol * O.chistat_obs(oT, oC, Xb)
    po = O.pval_invvar_uncalib_obs(oT, oC, Xb)
    assert abs(G.pval_invvar_uncalib(gT, gC, Xb) - po) < tol * po
    pc = O.pval_invvar_calib(oT, oC, Xb)
    assert abs(G.pval_invvar_calib(gT, gC, Xb) - pc) < tol * pc
    # a well-conditioned sentinel set (12 sentinels): the plain 1e-8 gate
    Xs = np.asfortranarray(Xb[:, ::5])
    assert np.linalg.cond(O.cliff_face(oT, oC, Xs)[1]) < 1e5
    assert abs(G.chistat(gT, gC, Xs) - O.chistat_obs(oT, oC, Xs)) < TOL * O.chistat_obs(oT, oC, Xs)
    assert abs(G.pval_invvar_uncalib(gT, gC, Xs) - O.pval_invvar_uncalib_obs(oT, oC, Xs)) < TOL * O.pval_invvar_uncalib_obs(oT, oC, Xs)
    assert abs(G.pval_invvar_calib(gT, gC, Xs) - O.pval_invvar_calib(oT, oC, Xs)) < TOL * O.pval_invvar_calib(oT, oC, Xs)
    mu, S = O.cliff_face(oT, oC, Xb)
    t, to = G.inverse_variance(mu, S), O.inverse_variance(mu, S)
    assert abs(t.mu - to.mu) < 1e-8 * abs(to.mu) and abs(t.sigma - to.sigma) < 1e-8 * to.sigma
    w = np.linspace(1.0, 2.0, 60)
    assert relerr(G.weight_at_units(gT, Xb, w), O.weight_at_units(oT, Xb, w)) < TOL
    u, uo = G.unweighted_mean(mu, S), O.unweighted_mean(mu, S)
    assert u.mu == uo.mu and abs(u.sigma - uo.sigma) < 1e-15
    x = G.generic_solve(S + 1e-10 * np.eye(60), np.ones(60))
    assert relerr(x, O.generic_solve(S + 1e-10 * np.eye(60), np.ones(60))) < 1e-6


# ---------------------------------------------------------------------------------------------
# a8 / a10 / a11 / a13: Monte-Carlo sharp-null machinery with identical host normals
# ---------------------------------------------------------------------------------------------
def _margin(sims, obs):
    return float(np.min(np.abs(sims - obs)) / abs(obs))


@pytest.mark.parametrize("kind,ragged,mean", [(2, 0, (0.0, 0.0)), (2, 23, (0.25, -0.1)), (3, -9, (0.0, 0.0))])
def test_null_chi2_invvar_mll_match_oracle(G, O, ctx, kind, ragged, mean):
    S = 200
    d, (oT, oC), (gT, gC) = _pair(G, O, m=160, nb=48, kind=kind, ragged=ragged, mean=mean)
    Xb = d["Xb"]
    Z = np.asfortranarray(np.random.default_rng(2).standard_normal((320, S)))
    # chi2
    p_o, c_o, obs_o, sims_o = O.boot_chi2test(oT, oC, Xb, Z)
    gN = G.make_null(gT, gC)
    oN_mll0 = gN.mll                                       # the pooled fit's mll: nsim_logP overwrites gN.mll below
    obs = G.chistat(gT, gC, Xb)
    sims, cnt = G.nsim_chi(gT, gC, gN, Xb, S, Z=Z, chi_obs=obs_o, return_count=True)
    assert relerr(sims, sims_o) < TOL and np.max(np.abs(sims - sims_o) / sims_o) < 1e-7
    assert _margin(sims_o, obs_o) > 1e-6                   # certificate for the exact count
    assert cnt == c_o and cnt / S == p_o
    # the library reports the same certificate itself (geordd_ctx_last_min_margin): min |chi2_s - chi_obs| over the draws
    mm = ctx.last_min_margin()
    assert abs(mm[0] - np.min(np.abs(sims - obs_o))) <= 1e-12 * obs_o and mm[2] == math.inf      # (mll not simulated)
    assert abs(obs - obs_o) < 1e-7 * obs_o
    assert gN.jitter_rounds == O.null_setup(oT, oC, Xb, 0.0).jitter_rounds
    # mll (re-uses the same handle: sentinels not needed)
    pm_o, cm_o, dobs_o, dsim_o = O.boot_mlltest(oT, oC, Z)
    simsm, cntm, mnull, malt = G.nsim_logP(gT, gC, gN, S, Z=Z, dmll_obs=dobs_o, return_count=True)
    assert relerr(malt - mnull, dsim_o) < TOL
    assert _margin(dsim_o, dobs_o) > 1e-6 and cntm == cm_o
    assert abs(ctx.last_min_margin()[2] - np.min(np.abs((malt - mnull) - dobs_o))) <= 1e-9 * abs(dobs_o)
    assert abs((gT.mll + gC.mll - oN_mll0) - dobs_o) < 1e-8 * (abs(gT.mll) + abs(gC.mll) + abs(oN_mll0))
    # nsim_logP leaves gpNull at the last draw (src/hypotest_mll.jl:6,10,14)
    oN = O.make_null(oT, oC)
    O.nsim_logP(oT, oC, oN, Z[:, -1:])
    assert relerr(gN.y, oN.y) < 1e-10 and relerr(gN.alpha, oN.alpha) < TOL and abs(gN.mll - oN.mll) < TOL * abs(oN.mll)
    # inverse variance (nugget 1e-10 before make_posdef!)
    pi_o, ci_o, pobs_o, psim_o = O.boot_invvar(oT, oC, Xb, Z)
    psim, cnti = G.nsim_invvar_pval_uncalib(gT, gC, Xb, S, Z=Z, pval_obs=pobs_o, return_count=True)
    assert np.max(np.abs(psim - psim_o)) < 1e-8
    assert _margin(psim_o, pobs_o) > 1e-6 and cnti == ci_o


def test_boot_drivers_end_to_end(G, O, ctx):
    """boot_chi2test / boot_mlltest / boot_invvar: p-values equal to the oracle's bit for bit."""
    S = 128
    d, (oT, oC), (gT, gC) = _pair(G, O, m=130, nb=30, kind=2)
    Z = np.asfortranarray(np.random.default_rng(9).standard_normal((260, S)))
    assert G.boot_chi2test(gT, gC, d["Xb"], S, Z=Z) == O.boot_chi2test(oT, oC, d["Xb"], Z)[0]
    assert G.boot_mlltest(gT, gC, S, Z=Z) == O.boot_mlltest(oT, oC, Z)[0]
    assert G.boot_invvar(gT, gC, d["Xb"], S, Z=Z) == O.boot_invvar(oT, oC, d["Xb"], Z)[0]


def test_device_rng_matches_oracle_and_is_shard_invariant(G, O, ctx):
    """Throughput mode: Philox normals generated on the device == the oracle's generator (few ulp), and
    statistics depend only on (seed, draw index): batch size and shard boundaries do not change a bit."""
    n, S = 301, 300
    Z = np.zeros((n, S), order="F")
    ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 77, 5, S, n, G.capi.dptr(Z)))
    Zo = O.philox_normals(77, 5, S, n)
    assert np.max(np.abs(Z - Zo)) < 1e-13
    d, (oT, oC), (gT, gC) = _pair(G, O, m=150, nb=32, kind=2, ragged=1)
    assert gT.nobs + gC.nobs == 300
    gN = G.make_null(gT, gC)
    Zd = np.zeros((300, S), order="F")
    ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 123, 0, S, 300, G.capi.dptr(Zd)))
    a = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=123)
    b = G.nsim_chi(gT, gC, gN, d["Xb"], S, Z=Zd)
    assert np.array_equal(a, b)                                        # device RNG == same normals from the host
    ctx.set_max_batch(64)
    try:
        c = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=123)
    finally:
        ctx.set_max_batch(0)
    assert np.array_equal(a, c)                                        # batch-size invariant
    parts = []
    for r in range(3):                                                 # 3 "ranks"
        lo, cnt = G.parallel.shard_draws(S, r, 3)
        parts.append(G.nsim_chi(gT, gC, gN, d["Xb"], cnt, seed=123, draw0=lo))
    assert np.array_equal(a, np.concatenate(parts))                    # GPU-count invariant
    sims_o = O.nsim_chi(oT, oC, O.make_null(oT, oC), d["Xb"], Zd)
    assert relerr(a, sims_o) < TOL


def test_nsim_power_matches_oracle(G, O, ctx):
    """power.jl: five p-values per alternative draw; corrected and as-shipped analytic calibration."""
    d, (oT, oC), (gT, gC) = _pair(G, O, m=120, nb=36, kind=2, ragged=5)
    Xb = d["Xb"]
    rng = np.random.default_rng(4)
    Zn = np.asfortranarray(rng.standard_normal((240, 300)))
    Zp = np.asfortranarray(rng.standard_normal((240, 90)))
    chi_null = O.nsim_chi(oT, oC, O.make_null(oT, oC), Xb, Zn)
    mnull, malt = O.nsim_logP(oT, oC, O.make_null(oT, oC), Zn)
    pinv_null = O.nsim_invvar_pval_uncalib(oT, oC, Xb, Zn)
    for bug in (False, True):
        po = O.nsim_power(oT, oC, 0.35, Xb, chi_null, malt - mnull, pinv_null, Zp, as_shipped_calib=bug)
        pg = G.nsim_power(gT, gC, 0.35, Xb, chi_null, malt - mnull, pinv_null, 90, Z=Zp, compat_calib_bug=bug)
        assert np.array_equal(pg[:, [0, 1, 3]], po[:, [0, 1, 3]])      # frequencies: exact
        assert np.max(np.abs(pg[:, 2] - po[:, 2])) < 1e-8
        assert np.max(np.abs(pg[:, 4] - po[:, 4])) < 1e-7


@pytest.mark.parametrize("kind,nb", [(2, 24), (3, 60)])
def test_nsim_invvar_calib(G, O, ctx, kind, nb):
    """nsim_invvar_calib (src/hypotest_invvar.jl:138-160): per draw the 3-argument pval_invvar_calib, i.e. LU solves
    against the un-jittered Sigma_b + 1e-10 I.  The GPU forms the draw-independent LU / covariance term once, by the same
    solves, so the p-values agree with the per-draw oracle to the parity tolerance."""
    d, (oT, oC), (gT, gC) = _pair(G, O, m=110, nb=nb, kind=kind)
    Z = np.asfortranarray(np.random.default_rng(8).standard_normal((220, 200)))
    po = O.nsim_invvar_calib(oT, oC, d["Xb"], Z)
    pg = G.nsim_invvar_calib(gT, gC, d["Xb"], 200, Z=Z)
    assert np.max(np.abs(pg - po)) < TOL
    assert 0.0 < pg.min() and pg.max() <= 1.0
    # the observed value (same code path, one draw = the data)
    assert abs(G.pval_invvar_calib(gT, gC, d["Xb"]) - O.pval_invvar_calib(oT, oC, d["Xb"])) < TOL


def test_golden_fixtures(G, golden):
    """GPU path against the committed oracle outputs (tests/golden/oracle_small.npz)."""
    g = golden["small"]
    for tag in ("mat32", "se"):
        kind, ell, s2, c2, mc, ln = g[f"{tag}_spec"]
        k = _kern(G, int(kind), ell, s2, c2)
        gT = G.GPE(g[f"{tag}_XT"], g[f"{tag}_YT"], G.MeanZero(), k, ln)
        gC = G.GPE(g[f"{tag}_XC"], g[f"{tag}_YC"], G.MeanZero(), k, ln)
        Xb, Z = g[f"{tag}_Xb"], np.asfortranarray(g[f"{tag}_Z"])
        assert relerr(gT.alpha, g[f"{tag}_alphaT"]) < TOL
        assert abs(gT.mll - float(g[f"{tag}_mllT"])) < TOL * abs(float(g[f"{tag}_mllT"]))
        mu, S = G.cliff_face(gT, gC, Xb)
        assert relerr(mu, g[f"{tag}_mu"]) < TOL and relerr(S, g[f"{tag}_Sigma"]) < TOL
        gN = G.make_null(gT, gC)
        _, cnt, mnull, malt = G.nsim_logP(gT, gC, gN, Z.shape[1], Z=Z, dmll_obs=float(g[f"{tag}_dmll_obs"]),
                                          return_count=True)
        assert relerr(malt - mnull, g[f"{tag}_dmll_sims"]) < TOL and cnt == int(g[f"{tag}_mll_cnt"])
        if tag == "mat32":   # well-conditioned cliff covariance: derived nb x nb quantities are gated (SURVEY §7.1 iii)
            sims, c2_ = G.nsim_chi(gT, gC, gN, Xb, Z.shape[1], Z=Z, chi_obs=float(g[f"{tag}_chi_obs"]), return_count=True)
            assert relerr(sims, g[f"{tag}_chi_sims"]) < 1e-7 and c2_ == int(g[f"{tag}_chi_cnt"])
            ps, c3 = G.nsim_invvar_pval_uncalib(gT, gC, Xb, Z.shape[1], Z=Z, pval_obs=float(g[f"{tag}_pinv_obs"]),
                                                return_count=True)
            assert np.max(np.abs(ps - g[f"{tag}_pinv_sims"])) < 1e-8 and c3 == int(g[f"{tag}_inv_cnt"])
            assert abs(G.pval_invvar_calib(gT, gC, Xb) - float(g[f"{tag}_calib"])) < 1e-6 * float(g[f"{tag}_calib"])


def test_mississippi_fixture_ill_conditioned(G, O, golden, ctx):
    """The reference's own fixture (SE, l = 100 km, Const(100), 200 sentinels): cond(Sigma_cliff) ~ 4e14, plain
    Cholesky fails and make_posdef! jitters.  Gate mu/Sigma normwise and the jitter protocol; chi2 draws are
    gated loosely (they sit on an eps-level-jittered matrix); the published size table is reproduced
    statistically with 20 000 device draws (notebook cell 28: 0.05 / 0.05 / 0.09 / 0.05 / 0.05)."""
    ms = golden["mississippi"]
    spec = O.KernelSpec(0, 100e3, 1.0, 1e4)
    k = G.SEIso(math.log(100e3), 0.0) + G.Const(math.log(100.0))
    nT, nC = 82, 64
    rng = np.random.default_rng(4)
    oT0, oC0 = O.gp_fit(spec, ms["X_MS"], np.zeros(nT), 0.0), O.gp_fit(spec, ms["X_LA"], np.zeros(nC), 0.0)
    y = O.prior_rand(O.make_null(oT0, oC0), rng.standard_normal(nT + nC))
    y -= y.mean()
    oT, oC = O.gp_fit(spec, ms["X_MS"], y[:nT], 0.0), O.gp_fit(spec, ms["X_LA"], y[nT:], 0.0)
    gT, gC = G.GPE(ms["X_MS"], y[:nT], G.MeanZero(), k, 0.0), G.GPE(ms["X_LA"], y[nT:], G.MeanZero(), k, 0.0)
    Xb = ms["sentinels"]
    mu, S = G.cliff_face(gT, gC, Xb)
    mo, So = O.cliff_face(oT, oC, Xb)
    assert relerr(mu, mo) < 1e-7 and relerr(S, So) < TOL
    gN = G.NullGPE(gT, gC, Xb, 0.0)
    ns = O.null_setup(oT, oC, Xb, 0.0)
    assert gN.jitter_rounds == ns.jitter_rounds >= 1
    # size table
    S_null, S_pow = 20000, 20000
    chi_null = G.nsim_chi(gT, gC, gN, Xb, S_null, seed=1)
    _, _, mnull, malt = G.nsim_logP(gT, gC, gN, S_null, seed=1, return_count=True)
    pinv_null = G.nsim_invvar_pval_uncalib(gT, gC, Xb, S_null, seed=1)
    out = G.nsim_power(gT, gC, 0.0, Xb, chi_null, malt - mnull, pinv_null, S_pow, seed=2)
    size = (out < 0.05).mean(axis=0)
    se = 4.0 * math.sqrt(0.05 * 0.95 / S_pow) * 2
    assert abs(size[0] - 0.05) < se and abs(size[1] - 0.05) < se and abs(size[3] - 0.05) < se and abs(size[4] - 0.05) < se
    assert 0.07 < size[2] < 0.11
    powr = (G.nsim_power(gT, gC, 1.2, Xb, chi_null, malt - mnull, pinv_null, S_pow, seed=3) < 0.05).mean(axis=0)
    published = np.array([0.66, 0.63, 0.87, 0.80, 0.80])
    assert np.all(np.abs(powr - published) < 0.03), powr


# ---------------------------------------------------------------------------------------------
# a3 / a4 / a5 / a6: joint fit with covariates
# ---------------------------------------------------------------------------------------------
def _mgp(G, O, n_half=50, kind=0, seed=1, const=400.0):
    d = O.synth_two_region(n_half, 10, seed=seed)
    n = 2 * n_half
    rdict = {True: G.RegionData(d["XT"], d["YT"], d["D"][:, :n_half]), False: G.RegionData(d["XC"], d["YC"], d["D"][:, n_half:])}
    return d, rdict, [0, n_half, n]


def test_multigp_gradient_reference_test(G, O, ctx):
    """test/test_gprealisations.jl:12-40: n = 100, SEIso(log .5, log 1) + Const(log 20), LinIso(log 1),
    logNoise 1.0, domean = false; analytic gradient ~ finite differences of mll (atol 1e-3) -- and both equal
    to the oracle."""
    d, rdict, off = _mgp(G, O)
    mg = G.MultiGPCovars(rdict, G.SEIso(math.log(0.5), 0.0) + G.Const(math.log(20.0)), G.LinIso(0.0), 1.0,
                         groupKeys=[True, False])
    kw = dict(noise=True, domean=False, kern=True, beta=True)
    grad = G.update_mll_and_dmll(mg, **kw).copy()
    x0 = G.mgp_get_params(mg, **kw)
    assert len(grad) == len(x0) == 5

    def f(x):
        G.mgp_set_params(mg, x, **kw)
        return G.update_mll(mg)

    num = np.array([(f(x0 + 1e-5 * e) - f(x0 - 1e-5 * e)) / 2e-5 for e in np.eye(5)])
    G.mgp_set_params(mg, x0, **kw)
    assert np.allclose(grad, num, atol=1e-3)
    om = O.MultiGP(D=d["D"], y=d["Y"], X=d["X"], group_off=off, spec=O.KernelSpec(0, 0.5, 1.0, 400.0), lin_ell2=1.0,
                   log_noise=1.0)
    go = O.multigp_update_mll_and_dmll(om, domean=False)
    assert relerr(grad, go) < 1e-7
    assert abs(G.update_mll(mg) - om.mll) < TOL * abs(om.mll) and relerr(mg.alpha, om.alpha) < TOL
    assert relerr(G.postmean_beta(mg), O.multigp_postmean_beta(om)) < 1e-7


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_multigp_matern_three_groups(G, O, ctx, kind):
    d = O.synth_two_region(90, 10, seed=5)
    off = [0, 50, 130, 180]
    rdict = {i: G.RegionData(d["X"][:, a:b], d["Y"][a:b], d["D"][:, a:b]) for i, (a, b) in enumerate(zip(off, off[1:]))}
    mg = G.MultiGPCovars(rdict, _kern(G, kind, 0.3, 1.5, 25.0), G.LinIso(math.log(2.0)), math.log(0.5), mean=G.MeanConst(0.2),
                         groupKeys=[0, 1, 2])
    om = O.MultiGP(D=d["D"], y=d["Y"], X=d["X"], group_off=off, spec=O.KernelSpec(kind, 0.3, 1.5, 25.0, 0.2), lin_ell2=4.0,
                   log_noise=math.log(0.5))
    go = O.multigp_update_mll_and_dmll(om)
    gg = G.update_mll_and_dmll(mg)
    assert len(gg) == len(go) == 6 and relerr(gg, go) < 1e-7
    assert abs(mg.mll - om.mll) < TOL * abs(om.mll)
    assert relerr(G.postmean_beta(mg), O.multigp_postmean_beta(om)) < 1e-7


def test_simulated_end_to_end(G, O, ctx):
    """test/test_simulated.jl:11-71 through the GPU path (n = 400): optimise the joint model, posterior mean
    of beta, residual GPs, cliff face, LATE inside its 99.8 % interval, calibrated p < 0.05."""
    tau = 0.3
    d = O.synth_two_region(200, 100, seed=1, tau=tau)
    rdict = {True: G.RegionData(d["XT"], d["YT"], d["D"][:, :200]), False: G.RegionData(d["XC"], d["YC"], d["D"][:, 200:])}
    mg = G.MultiGPCovars(rdict, G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0)), G.LinIso(0.0), 1.0,
                         groupKeys=[True, False])
    mll0 = mg.mll
    G.optimize(mg, noise=True, kern=True, domean=False, beta=True, options={"maxiter": 40})
    assert mg.mll > mll0
    beta = G.postmean_beta(mg)
    res = {k: G.residuals_data(rd, beta) for k, rd in rdict.items()}
    gpr = G.GPRealisations(res, mg.kernel, mg.logNoise, groupKeys=[True, False])
    assert abs(gpr.mll - (gpr[True].mll + gpr[False].mll)) < 1e-12
    mu, S = G.cliff_face(gpr[True], gpr[False], d["Xb"])
    t = G.inverse_variance(mu, S)
    assert t.quantile(0.001) < tau < t.quantile(0.999)
    # test/test_simulated.jl:55-63: the projected estimator with maxdist = 2 ell along the quarter-circle border
    ell = math.exp(mg.kernel.kleft.ll) if hasattr(mg.kernel, "kleft") else 0.3
    r = math.hypot(d["Xb"][0, 0], d["Xb"][1, 0])
    th = np.linspace(0.0, math.pi / 2, 1000)
    border = np.column_stack([r * np.cos(th), r * np.sin(th)])
    tp = G.proj_estimator(gpr[True], gpr[False], border, 2 * ell)
    assert tp.quantile(0.001) < tau < tp.quantile(0.999)
    assert G.pval_invvar_calib(gpr[True], gpr[False], d["Xb"]) < 0.05


@pytest.mark.parametrize("angle", [1.0, 33.0, 45.0, 90.0, 135.0, 179.0])
def test_placebo_drivers_match_oracle(G, O, ctx, angle):
    """placebo_chi / placebo_mll / placebo_invvar (src/hypotest_chi2.jl:62-76, src/hypotest_mll.jl:47-60,
    src/hypotest_invvar.jl:162-175) against the oracle's restatement, incl. the exact-degree angles 45 / 90 / 135 of the
    notebook sweep: identical split, identical bootstrap p-values for identical normals."""
    d = O.synth_two_region(150, 10, seed=2, tau=0.0)
    X, Y = np.asfortranarray(d["X"][:, :260]), d["Y"][:260].copy()
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(5.0))
    spec = O.KernelSpec(O.MAT32, 0.3, 1.0, 25.0)
    ln = math.log(0.3)
    Z = np.asfortranarray(np.random.default_rng(int(angle)).standard_normal((260, 200)))
    sh = O.shift_for_even_split(angle, X)
    assert G.shift_for_even_split(angle, X) == sh
    assert np.array_equal(G.left_points(angle, sh, X), O.left_points(angle, sh, X))
    assert G.placebo_chi(angle, X, Y, k, G.MeanZero(), ln, 200, Z=Z) == O.placebo_chi(angle, X, Y, spec, ln, Z)
    assert G.placebo_mll(angle, X, Y, k, G.MeanZero(), ln, 200, Z=Z) == O.placebo_mll(angle, X, Y, spec, ln, Z)
    pg, po = G.placebo_invvar(angle, X, Y, k, G.MeanZero(), ln), O.placebo_invvar(angle, X, Y, spec, ln)
    assert abs(pg - po) < 1e-7 * max(po, 1e-3)


@pytest.mark.parametrize("n,nrhs", [(1, 1), (5, 3), (31, 1), (32, 2), (33, 9), (100, 1), (200, 2), (200, 202), (256, 5), (257, 3),
                                    (604, 1), (640, 7), (641, 2)])
def test_blocked_lu_equals_unblocked_bit_for_bit(G, O, ctx, n, nrhs):
    """The pivoted LU solve behind Julia's `A \\ B` on a dense Matrix (observed chi2 / inverse-variance statistics,
    src/hypotest_chi2.jl:10, src/average_treatment_effect.jl:5-8; postmean_beta's p x p system, src/GPrealisationsCovars.jl:342):
    the blocked kernel performs the unblocked kernel's operations in the same order -- identical bits, pivots included --
    and both agree with the oracle's LAPACK getrf/getrs.  Matrices: a well-conditioned random one (every column needs a row
    exchange) and a cliff-face-like SPD one with a nugget (no exchanges)."""
    rng = np.random.default_rng(100 * n + nrhs)
    A1 = rng.standard_normal((n, n))
    t = np.linspace(0.0, 1.0, n)
    A2 = np.exp(-0.5 * ((t[:, None] - t[None, :]) / 0.3) ** 2) + 1e-6 * np.eye(n)
    for A, gate in ((A1, None), (A2, None)):
        B = rng.standard_normal((n, nrhs))
        lib = ctx.lib
        try:
            lib.geordd_dbg_set_lu_blocked(0)
            x0 = G.generic_solve(A, B)
        finally:
            lib.geordd_dbg_set_lu_blocked(1)
        x1 = G.generic_solve(A, B)
        assert np.array_equal(x0, x1)
        xo = O.generic_solve(A, B)
        cond = np.linalg.cond(A)
        assert relerr(x1, xo) < max(TOL, 50 * cond * np.finfo(float).eps)


def test_observed_statistics_from_the_null_handle_are_bit_identical(G, O, ctx):
    """geordd_null_chistat_obs / geordd_null_invvar_obs reuse the cliff face the handle evaluated when its sentinels were set;
    they must return exactly what chistat / pval_invvar_uncalib (src/hypotest_chi2.jl:4-11, src/hypotest_invvar.jl:1-8)
    return for the same sentinels and nugget, and refuse what they cannot serve."""
    d = O.synth_two_region(300, 40, seed=12)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
    gT, gC = G.GPE.fit_many([d["XT"], d["XC"]], [d["YT"], d["YC"]], G.MeanZero(), k, ln)
    n0 = G.NullGPE(gT, gC, d["Xb"], 0.0)
    assert n0.chistat_obs() == G.chistat(gT, gC, d["Xb"])
    n1 = G.NullGPE(gT, gC, d["Xb"], 1e-10)
    assert n1.invvar_obs()[2] == G.pval_invvar_uncalib(gT, gC, d["Xb"], 1e-10)
    with pytest.raises(ValueError):
        n1.chistat_obs()                       # chistat has no nugget
    with pytest.raises(ValueError):
        G.NullGPE(gT, gC).chistat_obs()        # no sentinels
    # and the bootstrap drivers built on them still equal the explicit three-step form
    Z = np.asfortranarray(np.random.default_rng(3).standard_normal((600, 256)))
    p = G.boot_chi2test(gT, gC, d["Xb"], 256, Z=Z)
    sims = G.nsim_chi(gT, gC, G.make_null(gT, gC), d["Xb"], 256, Z=Z)
    assert p == float(np.mean(sims > G.chistat(gT, gC, d["Xb"])))
    p = G.boot_invvar(gT, gC, d["Xb"], 256, Z=Z)
    pv = G.nsim_invvar_pval_uncalib(gT, gC, d["Xb"], 256, Z=Z)
    assert p == float(np.mean(pv < G.pval_invvar_uncalib(gT, gC, d["Xb"])))


def test_adjacent_sweep_matches_oracle(G, O, ctx):
    """The adjacent-pairs loop of the notebooks (src/regiondata.jl:64-68, NYC_analysis-2015.ipynb cell 74): three strips of
    the unit square, two shared borders; cliff face -> inverse-variance LATE -> calibrated p-value per pair against the
    oracle, both orientations filled like the notebook's dictionaries."""
    rng = np.random.default_rng(8)
    X = np.asfortranarray(rng.random((2, 420)))
    Y = np.sin(3.0 * X[0]) + 0.5 * (X[0] > 1 / 3) - 0.2 * (X[0] > 2 / 3) + 0.3 * rng.standard_normal(420)
    strip = np.minimum((X[0] * 3).astype(int), 2)
    rd = {name: G.RegionData(X[:, strip == i], Y[strip == i], np.empty((0, int((strip == i).sum())))) for i, name in enumerate("ABC")}
    k = G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0)); ln = math.log(0.3)
    spec = O.KernelSpec(O.SE_ISO, 0.3, 1.0, 400.0)
    gpr = G.GPRealisations(rd, k, ln)
    yy = np.linspace(0.02, 0.98, 40)
    borders = {("A", "B"): np.asfortranarray(np.vstack([np.full(40, 1 / 3), yy])),
               ("B", "C"): np.asfortranarray(np.vstack([np.full(40, 2 / 3), yy]))}
    tau, pv = G.adjacent_sweep(gpr, borders)
    assert set(tau) == {("A", "B"), ("B", "A"), ("B", "C"), ("C", "B")} and set(pv) == set(tau)
    og = {name: O.gp_fit(spec, rd[name].x, rd[name].y, ln) for name in "ABC"}
    for (ki, kj), Xb in borders.items():
        mu, Sig = O.cliff_face(og[ki], og[kj], Xb)
        tp = O.inverse_variance(mu, Sig)
        po = O.pval_invvar_calib(og[ki], og[kj], Xb)
        cond = np.linalg.cond(Sig + 1e-10 * np.eye(40))
        gate = max(TOL, 10 * cond * np.finfo(float).eps)
        assert abs(tau[(ki, kj)].mu - tp.mu) < gate * max(abs(tp.mu), tp.sigma) and abs(tau[(ki, kj)].sigma - tp.sigma) < gate * tp.sigma
        assert tau[(kj, ki)].mu == -tau[(ki, kj)].mu and tau[(kj, ki)].sigma == tau[(ki, kj)].sigma
        assert abs(pv[(ki, kj)] - po) < gate * max(po, 1e-3) and pv[(kj, ki)] == pv[(ki, kj)]


# ---------------------------------------------------------------------------------------------
# Full-size (BASELINE configs) size-independent properties
# ---------------------------------------------------------------------------------------------
def test_placebo_sweep_single_process(G, O, ctx):
    """placebo_sweep == the per-angle drivers (single process: plain loop; the rank-sharded path is covered by the
    gloo test)."""
    d = O.synth_two_region(60, 8, seed=4)
    k = G.SEIso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    X, Y = d["X"][:, :110], d["Y"][:110]
    angles = [10.0, 55.0, 120.0]
    ps = G.placebo_sweep("invvar", angles, X, Y, k, G.MeanZero(), math.log(0.3))
    assert np.array_equal(ps, np.array([G.placebo_invvar(a, X, Y, k, G.MeanZero(), math.log(0.3)) for a in angles]))
    pc = G.placebo_sweep("chi", angles[:2], X, Y, k, G.MeanZero(), math.log(0.3), 200, seed=3)
    assert np.all((pc >= 0.0) & (pc <= 1.0))
    # several independent (fit + test) problems side by side on the GPU: one context per host thread, same answers
    many = [float(a) for a in range(1, 180, 12)]
    for which in ("chi", "mll"):
        p1 = G.placebo_sweep(which, many, X, Y, k, G.MeanZero(), math.log(0.3), 300, seed=5)
        p4 = G.placebo_sweep(which, many, X, Y, k, G.MeanZero(), math.log(0.3), 300, seed=5, workers=4)
        assert np.array_equal(p1, p4)


def test_full_size_identities_c4(G, O, ctx):
    """C4 shape (4 000 units/side, Matern-3/2 + Const, 100 sentinels): y = L_N z  =>  (y-m)' K_N^-1 (y-m) = z'z,
    i.e. mll_null = -z'z/2 - logdet/2 - n log(2 pi)/2 to 1e-8 for device draws; shard invariance at size."""
    m = 4000
    d = O.synth_two_region(m, 100, seed=1)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    ln = math.log(0.3)
    gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln)
    gN = G.NullGPE(gT, gC, d["Xb"], 0.0)
    S = 256
    _, _, mnull, malt = G.nsim_logP(gT, gC, gN, S, seed=11, return_count=True)
    Z = np.zeros((2 * m, S), order="F")
    ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 11, 0, S, 2 * m, G.capi.dptr(Z)))
    zz = np.sum(Z * Z, axis=0)
    # logdet of the null factor from the fitted mll: mll_fit = -y'a/2 - logdet/2 - n log(2pi)/2
    const = np.median(mnull + zz / 2.0)
    assert np.max(np.abs(mnull + zz / 2.0 - const)) < 1e-8 * abs(const)
    chi = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=11)
    chi_b = np.concatenate([G.nsim_chi(gT, gC, gN, d["Xb"], 128, seed=11, draw0=0),
                            G.nsim_chi(gT, gC, gN, d["Xb"], 128, seed=11, draw0=128)])
    assert np.array_equal(chi, chi_b)
    assert np.all(chi > 0) and np.all(np.isfinite(malt))
    # sanity against the null distribution: E[chi2] = rank-ish, well below nb + 6 sd
    assert 1.0 < chi.mean() < 100 + 6 * math.sqrt(200)


# ---------------------------------------------------------------------------------------------
# Stress: the batched dataflow solves with MANY right-hand sides (more jobs than resident CTAs, real flag
# waits, persistent CTAs running many jobs).  Every column is checked -- this is the regime in which a ring fed
# by bulk async copies produced rare wrong tiles during round 1 (scripts/gpu_stress_solve.py).
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nrhs", [(1000, 20000), (700, 5632), (2000, 8192)])
def test_many_rhs_triangular_ops_every_column(G, O, ctx, n, nrhs):
    import ctypes as C
    import scipy.linalg as sl
    rng = np.random.default_rng(n)
    X = rng.random((2, n))
    K = O.cov_sym(O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0), X, 0.09)
    L = np.asfortranarray(np.linalg.cholesky(K))
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    refs = {2: L @ B, 0: sl.solve_triangular(L, B, lower=True), 1: sl.solve_triangular(L.T, B, lower=False)}
    refs[3], refs[4] = refs[0], refs[1]          # same solves with the right-hand sides in slab-blocked (ZB) storage
    for mode in (2, 0, 1, 3, 4):
        for rep in range(2):
            out = B.copy(order="F")
            ctx.check(ctx.lib.geordd_dbg_trsm(ctx._h, mode, G.capi.dptr(L), n, G.capi.dptr(out), nrhs, None))
            err = np.abs(out - refs[mode]).max(axis=0) / np.abs(refs[mode]).max()
            assert int((err > 1e-10).sum()) == 0, (mode, rep, float(err.max()))


@pytest.mark.parametrize("n,nv", [(1, 2), (257, 3), (3000, 1025), (5000, 2600)])
def test_projection_points_bit_exact(G, O, ctx, n, nv):
    """K7: nearest border point and distance for every unit == the GEOS restatement of the oracle BIT FOR BIT (each
    operation separately rounded), so the set of units within maxdist is identical; tiles of 1024 vertices, a
    MultiLineString break, repeated vertices, units on the border."""
    rng = np.random.default_rng(n + nv)
    th = np.linspace(0.0, np.pi / 2, nv)
    border = np.column_stack([0.7 * np.cos(th), 0.7 * np.sin(th)]) + 1e-3 * rng.standard_normal((nv, 2))
    part_start = ()
    if nv >= 1025:
        border[500] = border[499]                   # zero-length segment
        part_start = (700, 1024)                    # lines start inside a tile and exactly at a tile boundary
    X = rng.random((2, n)) * 1.1
    X[:, 0] = border[0]
    Xp, d = G.projection_points_all(X, border, part_start, ctx=ctx)
    Xo, do = O.projection_points_all(X, border, part_start)
    assert np.array_equal(d, do) and np.array_equal(Xp, Xo)
    for maxdist in (0.05, 0.3):
        assert np.array_equal(G.projection_points(X, border, maxdist, part_start, ctx=ctx), O.projection_points(X, border, maxdist, part_start))


def test_proj_estimator_and_weights_proj(G, O, ctx):
    """proj_estimator (src/border_projection.jl:24-31, used by test/test_simulated.jl:58) and weights_proj /
    weights_rho (src/weight_at_units.jl:44-53,80-86) against the oracle; the estimator equals sum(w_units * y)."""
    d, (oT, oC), (gT, gC) = _pair(G, O, m=300, nb=40, kind=0)
    th = np.linspace(0.0, np.pi / 2, 1000)
    r = np.hypot(d["Xb"][0, 0], d["Xb"][1, 0])
    border = np.column_stack([r * np.cos(th), r * np.sin(th)])
    maxdist = 0.15
    est = G.proj_estimator(gT, gC, border, maxdist)
    ref = O.proj_estimator(oT, oC, border, maxdist)
    assert abs(est.mu - ref.mu) < 1e-9 * max(1.0, abs(ref.mu)) and abs(est.sigma - ref.sigma) < 1e-9 * ref.sigma
    Xb, wb, wu = G.weights_proj(gT, gC, border, maxdist)
    assert Xb.shape[1] == wb.shape[0] > 10 and wu.shape[0] == gT.nobs + gC.nobs
    y = np.concatenate([gT.y, gC.y])
    assert abs(float(wu @ y) - est.mu) < 1e-8 * max(1.0, abs(est.mu))     # MeanZero: the estimator is linear in y
    Xs = G.sentinels(border, 25)
    assert np.allclose(Xs, O.polyline_sentinels(border, 25), rtol=0, atol=1e-14)
    Xr, wr, wur = G.weights_rho(gT, gC, lambda a, b: 1.0 + a, 25, border)
    wo = np.concatenate([O.weight_at_units(oT, Xs, wr), -O.weight_at_units(oC, Xs, wr)])
    assert relerr(wur, wo) < 1e-8


def test_points_in_region_and_infinite_proj(G, O, ctx):
    """Device within(point, region) == the oracle's ray-crossing restatement bit for bit (holes, several polygons,
    ring starts inside and at tile boundaries), and infinite_proj_sentinels / infinite_proj_estim / weights_geo match."""
    rng = np.random.default_rng(3)
    th = np.linspace(0.0, 2 * np.pi, 1301)
    shell = np.column_stack([1.0 + 0.9 * np.cos(th), 1.0 + 0.9 * np.sin(th) * (1 + 0.2 * np.sin(5 * th))]); shell[-1] = shell[0]
    th2 = np.linspace(0.0, 2 * np.pi, 200)
    hole = np.column_stack([1.0 + 0.2 * np.cos(-th2), 1.2 + 0.2 * np.sin(-th2)]); hole[-1] = hole[0]
    tri = np.array([[2.5, 0.0], [3.5, 0.0], [3.0, 1.0], [2.5, 0.0]])
    rings = np.vstack([shell, hole, tri])
    starts = (len(shell), len(shell) + len(hole))
    P = rng.random((2, 20000)) * np.array([[4.0], [2.5]]) - 0.2
    P[:, 0] = shell[10]                                   # on a vertex: boundary, not within
    got = G.points_in_region(P, rings, starts, ctx=ctx)
    want = O.points_in_region(P, rings, starts)
    assert np.array_equal(got, want) and 0.2 < got.mean() < 0.6 and not got[0]
    d, (oT, oC), (gT, gC) = _pair(G, O, m=200, nb=20, kind=0)
    r = np.hypot(d["Xb"][0, 0], d["Xb"][1, 0])
    tb = np.linspace(0.0, np.pi / 2, 400)
    border = np.column_stack([r * np.cos(tb), r * np.sin(tb)])
    region = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    Xb, w = G.infinite_proj_sentinels(gT, gC, border, region, 0.1, 0.04)
    Xo, wo = O.infinite_proj_sentinels(border, region, 0.1, 0.04)
    assert np.array_equal(Xb, Xo) and np.array_equal(w, wo) and Xb.shape[1] > 20
    est = G.infinite_proj_estim(gT, gC, border, region, 0.1, 0.04, density=lambda a, b: 1.0 + b)
    ref = O.infinite_proj_estim(oT, oC, border, region, 0.1, 0.04, density=lambda a, b: 1.0 + b)
    assert abs(est.mu - ref.mu) < 1e-9 * max(1.0, abs(ref.mu)) and abs(est.sigma - ref.sigma) < 1e-9 * ref.sigma
    Xg, wg, wu = G.weights_geo(gT, gC, border, region, 0.1, 0.04)
    y = np.concatenate([gT.y, gC.y])
    eg = G.infinite_proj_estim(gT, gC, border, region, 0.1, 0.04)
    assert abs(float(wu @ y) - eg.mu) < 1e-8 * max(1.0, abs(eg.mu))


@pytest.mark.parametrize("mT,mC", [(300, 280), (128, 200), (100, 150), (640, 1)])
def test_null_fit_reuses_treated_factor_bit_exact(G, O, ctx, mT, mC):
    """make_null: the pooled matrix starts with the treated block, so the tiles of the treated GP's factor inside that
    block are copied instead of recomputed (potrf prefill).  The pooled fit must equal a from-scratch fit of the
    pooled data BIT FOR BIT."""
    d = O.synth_two_region(max(mT, mC), 8, seed=6)
    XT, YT = d["XT"][:, :mT], d["YT"][:mT]
    XC, YC = d["XC"][:, :mC], d["YC"][:mC]
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    ln = math.log(0.3)
    gT, gC = G.GPE(XT, YT, G.MeanZero(), k, ln), G.GPE(XC, YC, G.MeanZero(), k, ln)
    gN = G.make_null(gT, gC)
    scratch = G.GPE(np.hstack([XT, XC]), np.concatenate([YT, YC]), G.MeanZero(), k, ln)
    assert gN.mll == scratch.mll                     # -(y'alpha + logdet + n log 2pi)/2: every tile of the factor enters
    oN = O.gp_fit(O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0), np.hstack([XT, XC]), np.concatenate([YT, YC]), ln)
    assert abs(gN.mll - oN.mll) < 1e-9 * abs(oN.mll)


def test_reference_notebook_outputs_through_the_gpu_path(G, O, golden, julia_replay, ctx):
    """The reference notebook's printed bootstrap results (see tests/test_oracle.py::
    test_reference_notebook_outputs_pin_the_oracle) through the CUDA path, with the notebook's own normals passed as
    host Z: boot_mlltest count 354 / 10 000 exactly; both chi2 counts (754, 759) for one common observed statistic
    within the (ill-conditioned, 2 %) noise of the device's own chistat; statistics equal to the oracle's."""
    ms, jn = golden["mississippi"], golden["julia_notebook"]
    k = G.SEIso(math.log(100e3), 0.0) + G.Const(math.log(100.0))
    y, Xb = julia_replay["Ysim"], ms["sentinels"]
    gT = G.GPE(ms["X_MS"], y[:82].copy(), G.MeanZero(), k, 0.0)
    gC = G.GPE(ms["X_LA"], y[82:].copy(), G.MeanZero(), k, 0.0)
    assert G.boot_mlltest(gT, gC, 10000, Z=julia_replay["Z_mll"]) == float(jn["boot_mll"]) == 0.0354
    gN = G.make_null(gT, gC)
    sa = np.sort(G.nsim_chi(gT, gC, gN, Xb, 0, Z=julia_replay["Z_chi_a"]))[::-1]
    sb = np.sort(G.nsim_chi(gT, gC, gN, Xb, 0, Z=julia_replay["Z_chi_b"]))[::-1]
    lo, hi = max(sa[754], sb[759]), min(sa[753], sb[758])
    assert lo < hi
    # the observed statistic is an LU solve against the raw cliff covariance (cond 1.8e14): Julia's value must have been
    # ~15.7495, LAPACK gives 15.767, the device's LU 15.88 -- each within the noise of that solve
    assert abs(0.5 * (lo + hi) - G.chistat(gT, gC, Xb)) < 2e-2 * G.chistat(gT, gC, Xb)
    spec = O.KernelSpec(O.SE_ISO, 100e3, 1.0, 100.0 ** 2, 0.0)
    oT = O.gp_fit(spec, np.asfortranarray(ms["X_MS"]), y[:82].copy(), 0.0)
    oC = O.gp_fit(spec, np.asfortranarray(ms["X_LA"]), y[82:].copy(), 0.0)
    so = O.nsim_chi(oT, oC, O.make_null(oT, oC), Xb, julia_replay["Z_chi_a"][:, :64])
    sg = G.nsim_chi(gT, gC, gN, Xb, 0, Z=julia_replay["Z_chi_a"][:, :64])
    assert relerr(sg, so) < 1e-6        # ill-conditioned sentinel system (cond 1e14 before the make_posdef! jitter)


def test_reference_notebook_table_through_the_gpu_path():
    """scripts/gpu_replay_mississippi_table.py: the published size / power table of the reference notebook (ten
    rejection rates over 10 000 power draws each, against 3 x 100 000 null statistics) through the CUDA path with the
    notebook's own normals -- equal to the notebook's table and, to all printed digits, to the oracle's replay."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gpu_replay_mississippi_table.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GPU TABLE == NOTEBOOK TABLE" in r.stdout, r.stdout
    lines = {ln.split()[0]: ln[ln.index("{"):ln.rindex("}") + 1] for ln in r.stdout.splitlines()
             if ln.split() and ln.split()[0] in ("GPU", "oracle") and "{" in ln}
    assert lines["GPU"] == lines["oracle"]


def test_mll_forward_only_mode_matches_stepwise(G, O, ctx):
    """Opt-in |L^-1 (y-m)|^2 / z'z evaluation of the mll quadratic forms == the stepwise reference evaluation to
    rounding, same exceedance count, also with non-zero means and unequal region sizes."""
    d, (oT, oC), (gT, gC) = _pair(G, O, m=260, nb=16, kind=2, ragged=17, mean=(0.3, -0.2))
    gN = G.make_null(gT, gC)
    S = 700
    d_obs = gT.mll + gC.mll - gN.mll
    _, cnt0, n0, a0 = G.nsim_logP(gT, gC, gN, S, seed=9, dmll_obs=d_obs, return_count=True)
    ctx.set_mll_forward_only(True)
    try:
        _, cnt1, n1, a1 = G.nsim_logP(gT, gC, gN, S, seed=9, dmll_obs=d_obs, return_count=True)
        ctx.set_max_batch(128)
        _, cnt2, n2, a2 = G.nsim_logP(gT, gC, gN, S, seed=9, dmll_obs=d_obs, return_count=True)
    finally:
        ctx.set_mll_forward_only(False)
        ctx.set_max_batch(0)
    assert relerr(n1, n0) < 1e-11 and relerr(a1, a0) < 1e-11
    assert cnt1 == cnt0
    assert np.array_equal(n1, n2) and np.array_equal(a1, a2) and cnt2 == cnt1


@pytest.mark.parametrize("n,nrhs", [(129, 1), (300, 8), (700, 9), (1000, 100), (2000, 128), (2000, 1), (4000, 3)])
def test_few_rhs_triangular_solves(G, O, ctx, n, nrhs):
    """<= 128 right-hand sides take the latency-optimised 8-column-slab kernel (solve_narrow.cu): forward and
    backward substitution against SciPy, repeated (persistent CTAs run several jobs, flags carry the epoch)."""
    import scipy.linalg as sl
    rng = np.random.default_rng(n + nrhs)
    X = rng.random((2, n))
    K = O.cov_sym(O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0), X, 0.09)
    L = np.asfortranarray(np.linalg.cholesky(K))
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    refs = {0: sl.solve_triangular(L, B, lower=True), 1: sl.solve_triangular(L.T, B, lower=False)}
    for mode in (0, 1):
        for rep in range(3):
            out = B.copy(order="F")
            ctx.check(ctx.lib.geordd_dbg_trsm(ctx._h, mode, G.capi.dptr(L), n, G.capi.dptr(out), nrhs, None))
            err = np.abs(out - refs[mode]).max(axis=0) / np.abs(refs[mode]).max()
            assert int((err > 1e-10).sum()) == 0, (mode, rep, float(err.max()))


@pytest.mark.parametrize("m", [1000, 2000])
def test_null_identities_every_draw_at_scale(G, O, ctx, m):
    """20 000 device draws: y = L_N z  =>  mll_null + z'z/2 is the same constant for EVERY draw (1e-9), the chi2
    statistics are finite and distributed like a null (no outliers from corrupted tiles), and repeated runs are
    bit-identical."""
    d = O.synth_two_region(m, 100, seed=1)
    k = G.Mat32Iso(math.log(0.3), 0.0) + G.Const(math.log(20.0))
    ln = math.log(0.3)
    gT, gC = G.GPE(d["XT"], d["YT"], G.MeanZero(), k, ln), G.GPE(d["XC"], d["YC"], G.MeanZero(), k, ln)
    gN = G.NullGPE(gT, gC, d["Xb"], 0.0)
    S = 20000
    _, _, mnull, malt = G.nsim_logP(gT, gC, gN, S, seed=5, return_count=True)
    Z = np.zeros((2 * m, S), order="F")
    ctx.check(ctx.lib.geordd_dbg_normals(ctx._h, 5, 0, S, 2 * m, G.capi.dptr(Z)))
    v = mnull + np.sum(Z * Z, axis=0) / 2.0
    const = np.median(v)
    assert np.max(np.abs(v - const)) < 1e-9 * abs(const)
    chi = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=5)
    chi2_ = G.nsim_chi(gT, gC, gN, d["Xb"], S, seed=5)
    assert np.array_equal(chi, chi2_)
    _, _, mnull2, malt2 = G.nsim_logP(gT, gC, gN, S, seed=5, return_count=True)
    assert np.array_equal(mnull, mnull2) and np.array_equal(malt, malt2)
    assert np.all(np.isfinite(chi)) and chi.max() < 20 * chi.mean() and 5.0 < chi.mean() < 60.0
    # 16 draws against the CPU oracle
    spec = O.KernelSpec(O.MAT32, 0.3, 1.0, 400.0)
    oT, oC = O.gp_fit(spec, d["XT"], d["YT"], ln), O.gp_fit(spec, d["XC"], d["YC"], ln)
    idx = np.array([0, 1, 63, 64, 4095, 4096, 9999, 10000, 12345, 15000, 16383, 16384, 19000, 19935, 19998, 19999])
    so = O.nsim_chi(oT, oC, O.make_null(oT, oC), d["Xb"], Z[:, idx])
    assert relerr(chi[idx], so) < 1e-8
