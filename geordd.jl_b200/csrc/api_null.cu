// C ABI, part 2: cliff face, observed statistics, Monte-Carlo sharp-null machinery, power study.
#include "api_internal.h"
#include <algorithm>

namespace geordd {

static const double kLog2Pi = 1.8378770664093453;   // log(2*pi)

// ---- cliff face on device (src/cliff_face.jl:3-9 + predict_f full_cov) --------------
void cliff_face_dev(geordd_gp* T, geordd_gp* C, const double* dXb, int nb, int nbp, double* dMuSlab, double* dSigma) {
    geordd_ctx* c = T->ctx;
    geordd_gp* sides[2] = {T, C};
    // The two sides are independent chains of small launches (covariance -> K_xb' alpha -> whitening solve, a chain of T
    // dependent row blocks -> V'V).  With few sentinels (the latency-optimised few-column solve, private job counter / flag
    // region per lane) the control side runs on a side stream beside the treated side; its two contributions are subtracted
    // on the main stream afterwards, in the order and with the roundings of the sequential form below (alpha = -1, beta = 1:
    // c - acc either way), so mu and Sigma are bit-identical to it.
    const bool two_streams = nb <= NARROW_MAX_COLS && T->npad > TILE && C->npad > TILE;
    double *KxbC = nullptr, *dotC = nullptr, *GC = nullptr;
    if (two_streams) {
        ensure_side_streams(c);
        const int mp = C->npad;
        c->tmp4.ensure((size_t)mp * nbp, c->stream, false);
        c->tmp5.ensure((size_t)nbp * 64 + (size_t)nbp * nbp, c->stream, false);
        const int sk_mu = pick_splitk(c, nbp, 64, mp), sk_g = pick_splitk(c, nbp, nbp, mp);
        c->ws2.ensure((size_t)std::max(sk_mu * 64, sk_g * nbp) * nbp, c->stream, false);
        KxbC = c->tmp4.p; dotC = c->tmp5.p; GC = dotC + (size_t)nbp * 64;
        cudaStream_t ss = c->side[0];
        GD_CUDA(cudaEventRecord(c->ev_fork, c->stream));   // after everything queued so far (the fits, the sentinels' upload)
        GD_CUDA(cudaStreamWaitEvent(ss, c->ev_fork, 0));
        launch_cov(ss, to_dev(C->k), C->dX, C->n, dXb, nb, C->dim, false, false, 0.0, KxbC, mp, mp, nbp, false);
        launch_gemm(ss, false, false, nbp, 64, mp, 1.0, KxbC, mp, C->dAlpha, mp, 0.0, dotC, nbp, sk_mu, c->ws2.p);
        launch_solve_narrow(ss, c->sched, SOLVE_FWD, C->dL, mp, KxbC, mp, nb, 1);
        launch_gemm(ss, false, false, nbp, nbp, mp, 1.0, KxbC, mp, KxbC, mp, 0.0, GC, nbp, sk_g, c->ws2.p);
        GD_CUDA(cudaEventRecord(c->ev_join[0], ss));
    }
    // Sigma = K_bb(T) + K_bb(C)  (each side's own kernel), then -= V'V per side
    for (int s = 0; s < 2; ++s) {
        ProfScope ps(c, PC_COV);
        launch_cov(c->stream, to_dev(sides[s]->k), dXb, nb, dXb, nb, 2, true, true, 0.0, dSigma, nbp, nbp, nbp, s == 1);
    }
    for (int s = 0; s < (two_streams ? 1 : 2); ++s) {
        geordd_gp* gp = sides[s];
        const int m = gp->n, mp = gp->npad;
        const double sign = s == 0 ? 1.0 : -1.0;
        c->tmp0.ensure((size_t)mp * nbp, c->stream, false);
        double* Kxb = c->tmp0.p;   // mp x nbp, ld mp : cov(kernel, X, Xb)
        {
            ProfScope ps(c, PC_COV);
            launch_cov(c->stream, to_dev(gp->k), gp->dX, m, dXb, nb, gp->dim, false, false, 0.0, Kxb, mp, mp, nbp, false);
        }
        // mu += sign * (m(Xb) + Kxb' alpha)
        gemm(c, false, false, nbp, 64, mp, sign, Kxb, mp, gp->dAlpha, mp, s == 0 ? 0.0 : 1.0, dMuSlab, nbp, 0);
        launch_add_const(c->stream, dMuSlab, nbp, nb, 1, sign * gp->k.mean_const);
        // V = L^-1 Kxb  (whiten!), Sigma -= V'V
        {
            ProfScope ps(c, PC_FWD);
            SolveEpilogue none;
            launch_solve(c->stream, c->sched, SOLVE_FWD, gp->dL, mp, Kxb, mp, nbp, none, nb);
        }
        gemm(c, false, false, nbp, nbp, mp, -1.0, Kxb, mp, Kxb, mp, 1.0, dSigma, nbp, 0);
    }
    if (two_streams) {
        GD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
        launch_axpby(c->stream, (long)nbp * 64, -1.0, dotC, 1.0, dMuSlab);
        launch_add_const(c->stream, dMuSlab, nbp, nb, 1, -C->k.mean_const);
        launch_axpby(c->stream, (long)nbp * nbp, -1.0, GC, 1.0, dSigma);
    }
    launch_mirror_lower(c->stream, dSigma, nbp, nbp);   // copytri!: exactly symmetric
}

void predict_mu_dev(geordd_gp* gp, const double* dKb, const double* dXb, int nb, int nbp, double* dMuSlab) {
    geordd_ctx* c = gp->ctx;
    const int mp = gp->npad;
    const double* Kb = dKb;
    if (!Kb) {
        c->tmp1.ensure((size_t)nbp * mp, c->stream, false);
        ProfScope ps(c, PC_COV);
        launch_cov(c->stream, to_dev(gp->k), dXb, nb, gp->dX, gp->n, gp->dim, false, false, 0.0, c->tmp1.p, nbp, nbp, mp,
                   false);
        Kb = c->tmp1.p;
    }
    gemm(c, true, false, nbp, 64, mp, 1.0, Kb, nbp, gp->dAlpha, mp, 0.0, dMuSlab, nbp, 0);
    launch_add_const(c->stream, dMuSlab, nbp, nb, 1, gp->k.mean_const);
}

static void check_pair(geordd_gp* T, geordd_gp* C) {
    if (!T || !C) throw ApiError(GEORDD_ERR_ARG, "gpT / gpC is NULL");
    if (T->ctx != C->ctx) throw ApiError(GEORDD_ERR_ARG, "gpT and gpC belong to different contexts");
    if (T->dim != 2 || C->dim != 2) throw ApiError(GEORDD_ERR_ARG, "cliff face needs 2-d coordinates");
}

// scratch for one cliff-face evaluation; caller frees
struct CliffTmp {
    double *dXb = nullptr, *dMu = nullptr, *dSig = nullptr;
    int nb = 0, nbp = 0;
    ~CliffTmp() { dfree(dXb); dfree(dMu); dfree(dSig); }
};
static void cliff_tmp(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, CliffTmp& t) {
    geordd_ctx* c = T->ctx;
    if (!Xb || nb <= 0) throw ApiError(GEORDD_ERR_ARG, "sentinels missing");
    t.nb = nb; t.nbp = round_up(nb, TILE);
    t.dXb = dalloc((size_t)2 * nb, c->stream, false);
    t.dMu = dalloc((size_t)t.nbp * 64, c->stream, true);
    t.dSig = dalloc((size_t)t.nbp * t.nbp, c->stream, true);
    GD_CUDA(cudaMemcpyAsync(t.dXb, Xb, (size_t)2 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    cliff_face_dev(T, C, t.dXb, nb, t.nbp, t.dMu, t.dSig);
}

// Solve (Sigma + nugget I) X = RHS by LU on device.  dSig is consumed.  RHS: nbp-strided, nrhs columns.
static void lu_solve_dev(geordd_ctx* c, double* dSig, int nb, int nbp, double nugget, double* dRhs, int nrhs) {
    if (nugget != 0.0) launch_add_diag(c->stream, dSig, nbp, nb, nugget);
    ProfScope ps(c, PC_LU);
    launch_lu_solve(c->stream, dSig, nbp, nb, dRhs, nbp, nrhs, c->d_info);
}

static double normal_two_sided_at_zero(double tau, double var) {
    double sd = std::sqrt(var);
    double z = (0.0 - tau) / sd;
    double cdf = std::erfc(-z * 0.7071067811865476) / 2.0;
    double ccdf = std::erfc(z * 0.7071067811865476) / 2.0;
    return 2.0 * std::min(cdf, ccdf);
}

// ---- analytic calibration, draw-independent part (src/hypotest_invvar.jl:78-99 / :111-131) ------------
// cov_mu_delta = WT KTT WT' + WC KCC WC' - WC KCT WT' - WT KCT' WC'   (as_shipped: first term only)
// returns into dOut (nbp x nbp, ld nbp)
static void calib_cov_mu_delta(geordd_gp* T, geordd_gp* C, const double* dXb, int nb, int nbp, bool as_shipped,
                               double* dOut) {
    geordd_ctx* c = T->ctx;
    const int mT = T->n, mTp = T->npad, mC = C->n, mCp = C->npad;
    gp_ensure_full_K(T);
    gp_ensure_full_K(C);
    // WT_c = KTT \ KTb (mTp x nbp), WC_c = KCC \ KCb
    double* WT = dalloc((size_t)mTp * nbp, c->stream, false);
    double* WC = nullptr; double* P = nullptr; double* KCT = nullptr; double* E = nullptr;
    try {
        SolveEpilogue none;
        {
            ProfScope ps(c, PC_COV);
            launch_cov(c->stream, to_dev(T->k), T->dX, mT, dXb, nb, 2, false, false, 0.0, WT, mTp, mTp, nbp, false);
        }
        pd_solve(c, T->dL, mTp, WT, mTp, nbp, none, nb);
        // P also serves as the nbp x nbp transpose buffer of E below: with more sentinels than units per side (the
        // Mississippi fixture: 82 / 64 units, 200 sentinels) max(mTp, mCp) * nbp alone is too small -- an out-of-bounds write
        // that stayed silent inside the caching pool until a trimmed pool left the neighbouring pages unmapped (round 2)
        P = dalloc((size_t)std::max(std::max(mTp, mCp), nbp) * nbp, c->stream, false);
        // P = KTT * WT_c ; out = WT_c' P
        gemm(c, true, false, mTp, nbp, mTp, 1.0, T->dK, mTp, WT, mTp, 0.0, P, mTp, 0);
        gemm(c, false, false, nbp, nbp, mTp, 1.0, WT, mTp, P, mTp, 0.0, dOut, nbp, 0);
        if (!as_shipped) {
            WC = dalloc((size_t)mCp * nbp, c->stream, false);
            {
                ProfScope ps(c, PC_COV);
                launch_cov(c->stream, to_dev(C->k), C->dX, mC, dXb, nb, 2, false, false, 0.0, WC, mCp, mCp, nbp, false);
            }
            pd_solve(c, C->dL, mCp, WC, mCp, nbp, none, nb);
            gemm(c, true, false, mCp, nbp, mCp, 1.0, C->dK, mCp, WC, mCp, 0.0, P, mCp, 0);
            gemm(c, false, false, nbp, nbp, mCp, 1.0, WC, mCp, P, mCp, 1.0, dOut, nbp, 0);
            // KCT = cov(kernel_C, X_C, X_T) (mCp x mTp); P = KCT * WT_c (mCp x nbp); E = WC_c' P ; out -= E + E'
            KCT = dalloc((size_t)mCp * mTp, c->stream, false);
            {
                ProfScope ps(c, PC_COV);
                launch_cov(c->stream, to_dev(C->k), C->dX, mC, T->dX, mT, 2, false, false, 0.0, KCT, mCp, mCp, mTp, false);
            }
            gemm(c, true, false, mCp, nbp, mTp, 1.0, KCT, mCp, WT, mTp, 0.0, P, mCp, 0);
            E = dalloc((size_t)nbp * nbp, c->stream, false);
            gemm(c, false, false, nbp, nbp, mCp, 1.0, WC, mCp, P, mCp, 0.0, E, nbp, 0);
            launch_axpby(c->stream, (long)nbp * nbp, -1.0, E, 1.0, dOut);
            double* Et = P;   // reuse: transpose of E
            launch_copy2d(c->stream, Et, nbp, E, nbp, nbp, nbp, true);
            launch_axpby(c->stream, (long)nbp * nbp, -1.0, Et, 1.0, dOut);
        }
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(WT); dfree(WC); dfree(P); dfree(KCT); dfree(E);
        throw;
    }
    dfree(WT); dfree(WC); dfree(P); dfree(KCT); dfree(E);
}

// ---- null handle ---------------------------------------------------------------------------------------
static void null_free(geordd_null* h) {
    if (!h) return;
    if (h->N) gp_free(h->N);
    dfree(h->dXb); dfree(h->dKbT); dfree(h->dKbC); dfree(h->dSig); dfree(h->dSigRaw); dfree(h->dLSig); dfree(h->dOnes); dfree(h->dMuObs);
    h->oNumLU.release();
    h->Z.release(); h->RN.release(); h->AN.release(); h->AT.release(); h->AC.release(); h->M.release(); h->W.release();
    h->partN.release(); h->partT.release(); h->partC.release(); h->partChi.release(); h->partNum.release();
    h->oChi.release(); h->oPval.release(); h->oNum.release(); h->oNull.release(); h->oAlt.release();
    delete h;
}

struct SimRequest {
    bool want_cliff = false, want_mll = false;
    bool want_lunum = false;   // also sum(lu(Sigma + nugget I) \ mu_b) per draw (sim_invvar_calib!, src/hypotest_invvar.jl:101)
    double tau = 0.0;
    long nsim = 0;
    unsigned long long seed = 0, draw0 = 0;
    const double* Z = nullptr;
    double chi_obs = 0.0, pval_obs = 0.0, dmll_obs = 0.0;
    // host outputs (nullable)
    double *chi = nullptr, *pval = nullptr, *num = nullptr, *numlu = nullptr, *mll_null = nullptr, *mll_alt = nullptr;
    double *last_y = nullptr, *last_alpha = nullptr;
    // device-resident outputs are left in h->oChi / oPval / oNum / oNull / oAlt as well
    long long n_chi = 0, n_inv = 0, n_mll = 0;
};

static long pick_batch(geordd_null* h, long nsim) {
    geordd_ctx* c = h->ctx;
    const long need = (nsim + 63) / 64 * 64;
    long want = need;
    if (c->max_batch > 0) want = std::min(want, std::max<long>(64, c->max_batch / 64 * 64));
    if (want <= h->cap_cols) return want;   // buffers of an earlier call suffice: no driver query on the hot path
    const long npN = h->N->npad, mT = h->T->npad, mC = h->C->npad, nbp = h->nbp;
    const long per_col = 8 * ((3 * npN + mT + mC) * 5 / 4 + 2 * nbp + (npN + mT + mC + 2 * nbp) / TILE + 8);   // ZB storage: 20/16
    // small requests (C3-sized calls, placebo halves) skip the driver query: cudaMemGetInfo costs up to milliseconds on a busy
    // host, and a request of under 1 GiB either fits or fails in dalloc with GEORDD_ERR_NOMEM
    if ((size_t)want * (size_t)per_col <= ((size_t)1 << 30)) return want;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    long budget = (long)std::min<size_t>((size_t)24 << 30, free_b / 2);
    long cap = std::max<long>(64, budget / per_col / 64 * 64);
    return std::min(cap, want);
}

static void run_sims(geordd_null* h, SimRequest& rq) {
    geordd_ctx* c = h->ctx;
    cudaStream_t st = c->stream;
    geordd_gp *T = h->T, *C = h->C, *N = h->N;
    const int n = N->n, npN = N->npad, mT = T->n, mTp = T->npad, mC = C->n, mCp = C->npad, nbp = h->nbp, nb = h->nb;
    if (rq.nsim <= 0) throw ApiError(GEORDD_ERR_ARG, "nsim must be positive");
    if (rq.want_cliff && nb <= 0) throw ApiError(GEORDD_ERR_ARG, "null handle was prepared without sentinels");
    // opt-in forward-only evaluation of the mll quadratic forms (include/geordd_b200.h, geordd_ctx_set_mll_forward_only)
    const bool fwd_only = c->mll_forward_only && rq.want_mll && !rq.want_cliff && !rq.last_y && !rq.last_alpha;
    const long cap = pick_batch(h, rq.nsim);
    const int TN = npN / TILE, TT = mTp / TILE, TC = mCp / TILE, TS = nbp > 0 ? nbp / TILE : 0;
    // (re)allocate simulation buffers for `cap` columns; padded regions must be zero, so a fresh
    // allocation is zero-filled and only live regions are ever written afterwards.
    if (cap > h->cap_cols) {
        h->Z.release(); h->RN.release(); h->AN.release(); h->AT.release(); h->AC.release(); h->M.release(); h->W.release();
        h->partN.release(); h->partT.release(); h->partC.release(); h->partChi.release(); h->partNum.release();
        h->cap_cols = 0;
    }
    // Z and every buffer the solves work on in place are in ZB storage (kernels.h)
    h->Z.ensure(zb_elems(npN, cap), st, true);
    h->AT.ensure(zb_elems(mTp, cap), st, true);
    h->AC.ensure(zb_elems(mCp, cap), st, true);
    if (rq.want_mll) {
        if (!fwd_only) {
            h->RN.ensure(zb_elems(npN, cap), st, true);
            h->AN.ensure(zb_elems(npN, cap), st, true);
        }
        h->partN.ensure((size_t)TN * cap, st, true);
        h->partT.ensure((size_t)TT * cap, st, true);
        h->partC.ensure((size_t)TC * cap, st, true);
    }
    if (rq.want_cliff) {
        h->M.ensure((size_t)nbp * cap, st, true);
        h->W.ensure((size_t)nbp * cap, st, true);
        h->partChi.ensure((size_t)TS * cap, st, true);
        h->partNum.ensure((size_t)TS * cap, st, true);
        h->oChi.ensure(rq.nsim, st, false);
        h->oPval.ensure(rq.nsim, st, false);
        h->oNum.ensure(rq.nsim, st, false);
        if (rq.want_lunum) h->oNumLU.ensure(rq.nsim, st, false);
    }
    if (rq.want_mll) {
        h->oNull.ensure(rq.nsim, st, false);
        h->oAlt.ensure(rq.nsim, st, false);
    }
    h->cap_cols = std::max(h->cap_cols, cap);
    GD_CUDA(cudaMemsetAsync(c->d_tally, 0, 4 * sizeof(unsigned long long), st));
    GD_CUDA(cudaMemsetAsync(c->d_tally + 4, 0xff, 4 * sizeof(unsigned long long), st));   // minimum margins: [4] chi2, [5] invvar, [6] mll

    const double mN = N->k.mean_const, mTc = T->k.mean_const, mCc = C->k.mean_const;
    for (long done = 0; done < rq.nsim; done += cap) {
        const long live = std::min(cap, rq.nsim - done);
        const int ncols = (int)((live + 63) / 64 * 64);
        // 1. standard normals (column s = draw done+s)
        if (rq.Z) {   // host normals: stage compactly, then re-block on the device
            c->tmp2.ensure((size_t)n * live, st, false);
            GD_CUDA(cudaMemcpyAsync(c->tmp2.p, rq.Z + (size_t)done * n, (size_t)n * live * sizeof(double), cudaMemcpyHostToDevice, st));
            ProfScope ps(c, PC_MISC);
            launch_zb_pack(st, c->tmp2.p, n, n, live, h->Z.p, npN, ncols);
        } else {
            ProfScope ps(c, PC_RNG);
            launch_philox_normals(st, rq.seed, rq.draw0 + (unsigned long long)done, ncols, n, h->Z.p, 0, npN, true);
        }
        // 2. prior_rand: y = m_N + L_N z (+ tau on treated), scattered into residual buffers
        {
            SolveEpilogue e;
            e.xb = true;
            e.nrows = n; e.split = mT; e.add_const = mN; e.add_T = rq.tau;
            if (rq.want_mll && !fwd_only) {
                e.out0 = h->RN.p; e.ld0 = npN; e.sub_0 = mN;
                e.out1 = h->AN.p; e.ld1 = npN; e.sub_1 = mN;
            }
            e.outT = h->AT.p; e.ldT = mTp; e.sub_T = mTc;
            e.outC = h->AC.p; e.ldC = mCp; e.sub_C = mCc;
            ProfScope ps(c, PC_TRMM);
            launch_solve(st, c->sched, SOLVE_TRMM, N->dL, npN, h->Z.p, npN, ncols, e);
        }
        // 3. update_alpha! on T, C (and N), with the (y - m)' alpha quadratic forms fused into the solves
        if (fwd_only) {
            // (y-m)'K^-1(y-m) = |L^-1 (y-m)|^2: forward solves only, sum of squares fused; pooled null: z'z
            launch_zb_colsumsq(st, h->Z.p, npN, ncols, h->partN.p);
            SolveEpilogue eT, eC;
            eT.xb = eC.xb = true;
            eT.dot_self = eC.dot_self = true;
            eT.dot0 = h->partT.p; eC.dot0 = h->partC.p;
            SolveSeqItem it[2] = {SolveSeqItem{SOLVE_FWD, T->dL, mTp, h->AT.p, eT, -1, gp_diag_inverse(T)},
                                  SolveSeqItem{SOLVE_FWD, C->dL, mCp, h->AC.p, eC, -1, gp_diag_inverse(C)}};
            ProfScope ps(c, PC_SEQ);
            launch_solve_seq(st, c->sched, it, 2, ncols);
        } else {
            SolveEpilogue eT, eC;
            eT.xb = eC.xb = true;
            if (rq.want_mll) {
                eT.wmat = h->RN.p; eT.ldw = npN; eT.wshift = mN - mTc; eT.wrows = mT; eT.wlive = mT; eT.dot0 = h->partT.p;
                eC.wmat = h->RN.p; eC.wrow0 = mT; eC.ldw = npN; eC.wshift = mN - mCc; eC.wrows = mC; eC.wlive = mC; eC.dot0 = h->partC.p;
            }
            // forward and backward solves of all GPs as ONE persistent kernel (solve.cu, fused solve sequence)
            SolveEpilogue none;
            none.xb = true;
            // order: the pooled (largest) solve first in each phase, so that the kernel's final tail comes from the
            // shortest jobs
            SolveSeqItem it[SOLVE_SEQ_MAX];
            int ns = 0, iN = -1;
            if (rq.want_mll) { iN = ns; it[ns++] = SolveSeqItem{SOLVE_FWD, N->dL, npN, h->AN.p, none, -1, gp_diag_inverse(N)}; }
            const int iT = ns; it[ns++] = SolveSeqItem{SOLVE_FWD, T->dL, mTp, h->AT.p, none, -1, gp_diag_inverse(T)};
            const int iC = ns; it[ns++] = SolveSeqItem{SOLVE_FWD, C->dL, mCp, h->AC.p, none, -1, gp_diag_inverse(C)};
            if (rq.want_mll) {
                SolveEpilogue eN;
                eN.xb = true;
                eN.wmat = h->RN.p; eN.ldw = npN; eN.dot0 = h->partN.p;
                it[ns++] = SolveSeqItem{SOLVE_BWD, N->dL, npN, h->AN.p, eN, iN, gp_diag_inverse(N)};
            }
            it[ns++] = SolveSeqItem{SOLVE_BWD, T->dL, mTp, h->AT.p, eT, iT, gp_diag_inverse(T)};
            it[ns++] = SolveSeqItem{SOLVE_BWD, C->dL, mCp, h->AC.p, eC, iC, gp_diag_inverse(C)};
            ProfScope ps(c, PC_SEQ);
            launch_solve_seq(st, c->sched, it, ns, ncols);
        }
        // Partials of dead columns in the last slab are garbage-free (zero inputs) and ignored.
        if (rq.want_mll) {
            FinishMll f{};
            f.partN = h->partN.p; f.partT = h->partT.p; f.partC = h->partC.p;
            f.nbN = fwd_only ? 1 : TN; f.nbT = TT; f.nbC = TC; f.ncols = ncols; f.ndraws = live;
            f.halflogdetN = N->logdet / 2.0; f.halflogdetT = T->logdet / 2.0; f.halflogdetC = C->logdet / 2.0;
            f.constN = (double)n * kLog2Pi / 2.0; f.constT = (double)mT * kLog2Pi / 2.0; f.constC = (double)mC * kLog2Pi / 2.0;
            f.dmll_obs = rq.dmll_obs; f.out_null = h->oNull.p; f.out_alt = h->oAlt.p; f.out_off = done;
            f.tally = c->d_tally + 2;
            f.margin = c->d_tally + 6;
            ProfScope ps(c, PC_FINISH);
            launch_finish_mll(st, f);
        }
        // 4. mu_b = (m_T - m_C) + cK_T alpha_T - cK_C alpha_C ; Sigma \ mu ; chi2 and inverse-variance
        if (rq.want_cliff) {
            gemm(c, true, false, nbp, ncols, mTp, 1.0, h->dKbT, nbp, h->AT.p, mTp, 0.0, h->M.p, nbp, 1, /*b_xb=*/true);
            gemm(c, true, false, nbp, ncols, mCp, -1.0, h->dKbC, nbp, h->AC.p, mCp, 1.0, h->M.p, nbp, 1, /*b_xb=*/true);
            launch_add_const(st, h->M.p, nbp, nb, ncols, mTc - mCc);
            GD_CUDA(cudaMemcpyAsync(h->W.p, h->M.p, (size_t)nbp * ncols * sizeof(double), cudaMemcpyDeviceToDevice, st));
            if (rq.want_lunum) {
                // the reference's per-draw `Sigma_b \ mu_b` with Sigma_b a plain Matrix = LU with partial pivoting of the
                // un-jittered Sigma + nugget I; the matrix is the same for every draw, each CTA refactors it for its chunk
                constexpr int kColsPerCta = 128;
                const size_t ctas = (size_t)((live + kColsPerCta - 1) / kColsPerCta);
                c->tmp3.ensure((size_t)nbp * ncols + ctas * nb * nb, st, false);
                GD_CUDA(cudaMemcpyAsync(c->tmp3.p, h->M.p, (size_t)nbp * ncols * sizeof(double), cudaMemcpyDeviceToDevice, st));
                {
                    ProfScope ps(c, PC_LU);
                    launch_lu_solve_multi(st, h->dSigRaw, nbp, nb, c->tmp3.p + (size_t)nbp * ncols, c->tmp3.p, nbp, live, kColsPerCta,
                                          c->d_info);
                }
                launch_colsum(st, c->tmp3.p, nbp, nb, live, h->oNumLU.p + done);
            }
            SolveEpilogue e;
            e.wmat = h->M.p; e.ldw = nbp; e.dot0 = h->partChi.p;
            e.wvec = h->dOnes; e.dot1 = h->partNum.p;
            pd_solve(c, h->dLSig, nbp, h->W.p, nbp, ncols, e);
            FinishCliff f{};
            f.part_chi = h->partChi.p; f.part_num = h->partNum.p; f.nb = TS; f.ncols = ncols; f.ndraws = live;
            f.chi_obs = rq.chi_obs; f.denom = h->denom; f.pval_obs = rq.pval_obs;
            f.out_chi = h->oChi.p; f.out_pval = h->oPval.p; f.out_num = h->oNum.p; f.out_off = done;
            f.tally_chi = c->d_tally + 0; f.tally_inv = c->d_tally + 1;
            f.margin_chi = c->d_tally + 4; f.margin_inv = c->d_tally + 5;
            ProfScope ps(c, PC_FINISH);
            launch_finish_cliff(st, f);
        }
        if (done + cap >= rq.nsim && rq.want_mll && (rq.last_y || rq.last_alpha)) {
            // the reference leaves gpNull.y / .alpha at the final draw (src/hypotest_mll.jl:6,10)
            c->tmp2.ensure((size_t)2 * n, st, false);
            if (rq.last_y) {
                launch_zb_unpack(st, h->RN.p, npN, c->tmp2.p, n, n, 1, live - 1);
                GD_CUDA(cudaMemcpyAsync(rq.last_y, c->tmp2.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
                GD_CUDA(cudaStreamSynchronize(st));
                for (int i = 0; i < n; ++i) rq.last_y[i] += mN;
            }
            if (rq.last_alpha) {
                launch_zb_unpack(st, h->AN.p, npN, c->tmp2.p + n, n, n, 1, live - 1);
                GD_CUDA(cudaMemcpyAsync(rq.last_alpha, c->tmp2.p + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
            }
        }
    }
    const size_t bytes = (size_t)rq.nsim * sizeof(double);
    if (rq.want_cliff) {
        if (rq.chi) GD_CUDA(cudaMemcpyAsync(rq.chi, h->oChi.p, bytes, cudaMemcpyDeviceToHost, st));
        if (rq.pval) GD_CUDA(cudaMemcpyAsync(rq.pval, h->oPval.p, bytes, cudaMemcpyDeviceToHost, st));
        if (rq.num) GD_CUDA(cudaMemcpyAsync(rq.num, h->oNum.p, bytes, cudaMemcpyDeviceToHost, st));
        if (rq.numlu && rq.want_lunum) GD_CUDA(cudaMemcpyAsync(rq.numlu, h->oNumLU.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    if (rq.want_mll) {
        if (rq.mll_null) GD_CUDA(cudaMemcpyAsync(rq.mll_null, h->oNull.p, bytes, cudaMemcpyDeviceToHost, st));
        if (rq.mll_alt) GD_CUDA(cudaMemcpyAsync(rq.mll_alt, h->oAlt.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    GD_CUDA(cudaMemcpyAsync(c->h_tally, c->d_tally, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    check_abort(c);
    rq.n_chi = (long long)c->h_tally[0];
    rq.n_inv = (long long)c->h_tally[1];
    rq.n_mll = (long long)c->h_tally[2];
    for (int k = 0; k < 3; ++k) {   // +Inf when the statistic was not simulated (or the threshold was infinite / NaN)
        double m;
        std::memcpy(&m, &c->h_tally[4 + k], sizeof(double));
        c->min_margin[k] = (c->h_tally[4 + k] == ~0ull || !(m == m)) ? INFINITY : m;
    }
}

// draw-independent calibration constant for the 7-argument pval_invvar_calib (PD solves)
static double null_cov_mu_tau(geordd_null* h, bool as_shipped) {
    const int idx = as_shipped ? 1 : 0;
    if (h->calib_ready[idx]) return h->cov_mu_tau[idx];
    geordd_ctx* c = h->ctx;
    const int nb = h->nb, nbp = h->nbp;
    double* cmd = dalloc((size_t)nbp * nbp, c->stream, true);
    double* one = nullptr;
    double out = 0.0;
    try {
        calib_cov_mu_delta(h->T, h->C, h->dXb, nb, nbp, as_shipped, cmd);
        // cov_mu_tau = sum((Sigma \ cmd) * (Sigma \ 1))
        SolveEpilogue none;
        pd_solve(c, h->dLSig, nbp, cmd, nbp, nbp, none, nb);
        one = dalloc((size_t)nbp * 64, c->stream, true);
        GD_CUDA(cudaMemcpyAsync(one, h->dOnes, (size_t)nbp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        pd_solve(c, h->dLSig, nbp, one, nbp, 64, none, 1);
        // v = (Sigma\cmd) * s  then sum(v)  == sum over all entries of (Sigma\cmd)(i,j) * s(j)
        double* v = dalloc((size_t)nbp * 64, c->stream, true);
        gemm(c, true, false, nbp, 64, nbp, 1.0, cmd, nbp, one, nbp, 0.0, v, nbp, 1);
        launch_reduce(c->stream, 1, v, nullptr, 0, nb, 0, c->d_scal + 4);
        out = read_scalar(c, c->d_scal + 4);
        dfree(v);
    } catch (...) {
        dfree(cmd); dfree(one);
        throw;
    }
    dfree(cmd); dfree(one);
    h->cov_mu_tau[idx] = out;
    h->calib_ready[idx] = true;
    return out;
}

}  // namespace geordd

using namespace geordd;

#define API_BEGIN(ctxptr)                  \
    geordd_ctx* c__ = (ctxptr);            \
    if (!c__) return GEORDD_ERR_ARG;       \
    try {                                  \
        ctx_use(c__);
#define API_END                            \
    }                                      \
    catch (const ApiError& e) {            \
        return ctx_fail(c__, e);           \
    }                                      \
    catch (const std::exception& e) {      \
        c__->err = e.what();               \
        return GEORDD_ERR_CUDA;            \
    }                                      \
    return 0;

extern "C" {

int geordd_gp_predict_mu(geordd_gp* gp, const double* Xb, int nb, double* mu) {
    if (!gp) return GEORDD_ERR_ARG;
    API_BEGIN(gp->ctx)
    if (!Xb || !mu || nb <= 0) throw ApiError(GEORDD_ERR_ARG, "predict_mu: bad arguments");
    const int nbp = round_up(nb, TILE);
    double* dXb = dalloc((size_t)gp->dim * nb, c__->stream, false);
    double* dMu = nullptr;
    try {
        dMu = dalloc((size_t)nbp * 64, c__->stream, true);
        GD_CUDA(cudaMemcpyAsync(dXb, Xb, (size_t)gp->dim * nb * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        predict_mu_dev(gp, nullptr, dXb, nb, nbp, dMu);
        GD_CUDA(cudaMemcpyAsync(mu, dMu, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
        GD_CUDA(cudaStreamSynchronize(c__->stream));
    } catch (...) {
        dfree(dXb); dfree(dMu);
        throw;
    }
    dfree(dXb); dfree(dMu);
    API_END
}

int geordd_cliff_face(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double* mu, double* Sigma) {
    if (!T) return GEORDD_ERR_ARG;
    API_BEGIN(T->ctx)
    check_pair(T, C);
    CliffTmp t;
    cliff_tmp(T, C, Xb, nb, t);
    if (mu) GD_CUDA(cudaMemcpyAsync(mu, t.dMu, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
    if (Sigma) download2d(c__, Sigma, nb, t.dSig, t.nbp, nb, nb);
    check_abort(c__);
    API_END
}

int geordd_projection_points(geordd_ctx* c, const double* X, int n, const double* V, int nv, const int* part_start, int nparts,
                             double* Xproj, double* dist) {
    API_BEGIN(c)
    if (!X || !V || !Xproj || !dist || n <= 0 || nv < 2 || nparts < 0 || (nparts > 0 && !part_start))
        throw ApiError(GEORDD_ERR_ARG, "projection_points: bad arguments");
    std::vector<unsigned char> brk((size_t)nv, 0);
    brk[0] = 1;
    for (int k = 0; k < nparts; ++k) {
        if (part_start[k] < 0 || part_start[k] >= nv) throw ApiError(GEORDD_ERR_ARG, "projection_points: part_start out of range");
        brk[part_start[k]] = 1;
    }
    bool any_segment = false;
    for (int v = 1; v < nv; ++v) any_segment |= !brk[v];
    if (!any_segment) throw ApiError(GEORDD_ERR_ARG, "projection_points: the border has no segment");
    double *dX = nullptr, *dV = nullptr, *dP = nullptr, *dD = nullptr, *dB = nullptr;
    try {
        dX = dalloc((size_t)2 * n, c__->stream, false);
        dV = dalloc((size_t)2 * nv, c__->stream, false);
        dP = dalloc((size_t)2 * n, c__->stream, false);
        dD = dalloc((size_t)n, c__->stream, false);
        dB = dalloc((size_t)(nv + 7) / 8 + 1, c__->stream, false);
        GD_CUDA(cudaMemcpyAsync(dX, X, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        GD_CUDA(cudaMemcpyAsync(dV, V, (size_t)2 * nv * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        GD_CUDA(cudaMemcpyAsync(dB, brk.data(), (size_t)nv, cudaMemcpyHostToDevice, c__->stream));
        {
            ProfScope ps(c__, PC_MISC);
            launch_projection(c__->stream, dX, n, dV, reinterpret_cast<const unsigned char*>(dB), nv, dP, dD);
        }
        GD_CUDA(cudaMemcpyAsync(Xproj, dP, (size_t)2 * n * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
        GD_CUDA(cudaMemcpyAsync(dist, dD, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
        GD_CUDA(cudaStreamSynchronize(c__->stream));
        GD_CUDA(cudaGetLastError());
    } catch (...) {
        dfree(dX); dfree(dV); dfree(dP); dfree(dD); dfree(dB);
        throw;
    }
    dfree(dX); dfree(dV); dfree(dP); dfree(dD); dfree(dB);
    API_END
}

int geordd_points_in_region(geordd_ctx* c, const double* P, int n, const double* R, int nr, const int* ring_start, int nrings,
                            unsigned char* inside) {
    API_BEGIN(c)
    if (!P || !R || !inside || n <= 0 || nr < 4 || nrings < 0 || (nrings > 0 && !ring_start))
        throw ApiError(GEORDD_ERR_ARG, "points_in_region: bad arguments");
    std::vector<unsigned char> brk((size_t)nr, 0);
    brk[0] = 1;
    for (int k = 0; k < nrings; ++k) {
        if (ring_start[k] < 0 || ring_start[k] >= nr) throw ApiError(GEORDD_ERR_ARG, "points_in_region: ring_start out of range");
        brk[ring_start[k]] = 1;
    }
    double *dP = nullptr, *dR = nullptr, *dB = nullptr, *dO = nullptr;
    try {
        dP = dalloc((size_t)2 * n, c__->stream, false);
        dR = dalloc((size_t)2 * nr, c__->stream, false);
        dB = dalloc((size_t)(nr + 7) / 8 + 1, c__->stream, false);
        dO = dalloc((size_t)(n + 7) / 8 + 1, c__->stream, false);
        GD_CUDA(cudaMemcpyAsync(dP, P, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        GD_CUDA(cudaMemcpyAsync(dR, R, (size_t)2 * nr * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        GD_CUDA(cudaMemcpyAsync(dB, brk.data(), (size_t)nr, cudaMemcpyHostToDevice, c__->stream));
        {
            ProfScope ps(c__, PC_MISC);
            launch_in_region(c__->stream, dP, n, dR, reinterpret_cast<const unsigned char*>(dB), nr, reinterpret_cast<unsigned char*>(dO));
        }
        GD_CUDA(cudaMemcpyAsync(inside, dO, (size_t)n, cudaMemcpyDeviceToHost, c__->stream));
        GD_CUDA(cudaStreamSynchronize(c__->stream));
        GD_CUDA(cudaGetLastError());
    } catch (...) {
        dfree(dP); dfree(dR); dfree(dB); dfree(dO);
        throw;
    }
    dfree(dP); dfree(dR); dfree(dB); dfree(dO);
    API_END
}

int geordd_chistat_obs(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double* chi2) {
    if (!T) return GEORDD_ERR_ARG;
    API_BEGIN(T->ctx)
    check_pair(T, C);
    if (!chi2) throw ApiError(GEORDD_ERR_ARG, "chistat_obs: output is NULL");
    CliffTmp t;
    cliff_tmp(T, C, Xb, nb, t);
    double* x = dalloc((size_t)t.nbp, c__->stream, true);
    try {
        GD_CUDA(cudaMemcpyAsync(x, t.dMu, (size_t)t.nbp * sizeof(double), cudaMemcpyDeviceToDevice, c__->stream));
        lu_solve_dev(c__, t.dSig, nb, t.nbp, 0.0, x, 1);   // Sigma \ mu  (LU, raw Sigma)
        launch_reduce(c__->stream, 2, t.dMu, x, 0, nb, 0, c__->d_scal);
        check_abort(c__);
        *chi2 = read_scalar(c__, c__->d_scal);
    } catch (...) {
        dfree(x);
        throw;
    }
    dfree(x);
    API_END
}

// inverse_variance on device data: rhs columns [1, mu]
static void invvar_dev(geordd_ctx* c, double* dSig, const double* dMu, int nb, int nbp, double nugget, double* tau,
                       double* var) {
    double* rhs = dalloc((size_t)nbp * 2, c->stream, true);
    try {
        launch_fill(c->stream, rhs, nb, 1.0);
        GD_CUDA(cudaMemcpyAsync(rhs + nbp, dMu, (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        lu_solve_dev(c, dSig, nb, nbp, nugget, rhs, 2);
        launch_reduce(c->stream, 1, rhs, nullptr, 0, nb, 0, c->d_scal);
        launch_reduce(c->stream, 1, rhs + nbp, nullptr, 0, nb, 0, c->d_scal + 1);
        GD_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        check_abort(c);
        double denom = c->h_scal[0], num = c->h_scal[1];
        *tau = num / denom;
        *var = 1.0 / denom;
    } catch (...) {
        dfree(rhs);
        throw;
    }
    dfree(rhs);
}

int geordd_invvar_obs(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, double* tau_hat,
                      double* var_tau, double* pval) {
    if (!T) return GEORDD_ERR_ARG;
    API_BEGIN(T->ctx)
    check_pair(T, C);
    CliffTmp t;
    cliff_tmp(T, C, Xb, nb, t);
    double tau = 0, var = 0;
    invvar_dev(c__, t.dSig, t.dMu, nb, t.nbp, nugget, &tau, &var);
    if (tau_hat) *tau_hat = tau;
    if (var_tau) *var_tau = var;
    if (pval) *pval = normal_two_sided_at_zero(tau, var);
    API_END
}

// The observed statistics from a prepared null handle: geordd_null_set_sentinels has already evaluated the cliff face of the
// observed outcomes with the kernels geordd_chistat_obs / geordd_invvar_obs use, so boot_chi2test / boot_invvar need not
// evaluate it a second time.  Same operations on the same numbers: results are bit-identical to the stand-alone calls.
int geordd_null_chistat_obs(geordd_null* h, double* chi2) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    if (!chi2) throw ApiError(GEORDD_ERR_ARG, "null_chistat_obs: output is NULL");
    if (h->nb <= 0 || !h->dMuObs) throw ApiError(GEORDD_ERR_ARG, "null handle was prepared without sentinels");
    if (h->nugget != 0.0) throw ApiError(GEORDD_ERR_ARG, "null_chistat_obs: chistat uses the cliff covariance without a nugget; this handle has one");
    const int nb = h->nb, nbp = h->nbp;
    double* x = dalloc((size_t)nbp, c__->stream, true);
    double* sig = nullptr;
    try {
        sig = dalloc((size_t)nbp * nbp, c__->stream, false);
        GD_CUDA(cudaMemcpyAsync(sig, h->dSigRaw, (size_t)nbp * nbp * sizeof(double), cudaMemcpyDeviceToDevice, c__->stream));
        GD_CUDA(cudaMemcpyAsync(x, h->dMuObs, (size_t)nbp * sizeof(double), cudaMemcpyDeviceToDevice, c__->stream));
        lu_solve_dev(c__, sig, nb, nbp, 0.0, x, 1);
        launch_reduce(c__->stream, 2, h->dMuObs, x, 0, nb, 0, c__->d_scal);
        check_abort(c__);
        *chi2 = read_scalar(c__, c__->d_scal);
    } catch (...) {
        dfree(x); dfree(sig);
        throw;
    }
    dfree(x); dfree(sig);
    API_END
}

int geordd_null_invvar_obs(geordd_null* h, double* tau_hat, double* var_tau, double* pval) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    if (h->nb <= 0 || !h->dMuObs) throw ApiError(GEORDD_ERR_ARG, "null handle was prepared without sentinels");
    const int nb = h->nb, nbp = h->nbp;
    double* sig = dalloc((size_t)nbp * nbp, c__->stream, false);
    double tau = 0, var = 0;
    try {
        GD_CUDA(cudaMemcpyAsync(sig, h->dSigRaw, (size_t)nbp * nbp * sizeof(double), cudaMemcpyDeviceToDevice, c__->stream));
        invvar_dev(c__, sig, h->dMuObs, nb, nbp, 0.0 /* the handle's nugget is already inside dSigRaw */, &tau, &var);
    } catch (...) {
        dfree(sig);
        throw;
    }
    dfree(sig);
    if (tau_hat) *tau_hat = tau;
    if (var_tau) *var_tau = var;
    if (pval) *pval = normal_two_sided_at_zero(tau, var);
    API_END
}

int geordd_inverse_variance(geordd_ctx* c, const double* mu, const double* Sigma, int nb, double nugget, double* tau_hat,
                            double* var_tau) {
    API_BEGIN(c)
    if (!mu || !Sigma || nb <= 0 || !tau_hat || !var_tau) throw ApiError(GEORDD_ERR_ARG, "inverse_variance: bad arguments");
    const int nbp = round_up(nb, TILE);
    double* dS = dalloc((size_t)nbp * nbp, c->stream, true);
    double* dM = nullptr;
    try {
        dM = dalloc((size_t)nbp, c->stream, true);
        upload2d(c, dS, nbp, Sigma, nb, nb, nb);
        GD_CUDA(cudaMemcpyAsync(dM, mu, (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        invvar_dev(c, dS, dM, nb, nbp, nugget, tau_hat, var_tau);
    } catch (...) {
        dfree(dS); dfree(dM);
        throw;
    }
    dfree(dS); dfree(dM);
    API_END
}

int geordd_solve_general(geordd_ctx* c, const double* A, int n, double* B, int nrhs) {
    API_BEGIN(c)
    if (!A || !B || n <= 0 || nrhs <= 0) throw ApiError(GEORDD_ERR_ARG, "solve_general: bad arguments");
    double* dA = dalloc((size_t)n * n, c->stream, false);
    double* dB = nullptr;
    try {
        dB = dalloc((size_t)n * nrhs, c->stream, false);
        GD_CUDA(cudaMemcpyAsync(dA, A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GD_CUDA(cudaMemcpyAsync(dB, B, (size_t)n * nrhs * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        {
            ProfScope ps(c, PC_LU);
            launch_lu_solve(c->stream, dA, n, n, dB, n, nrhs, c->d_info);
        }
        GD_CUDA(cudaMemcpyAsync(B, dB, (size_t)n * nrhs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaMemcpyAsync(c->h_info + 1, c->d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
        if (c->h_info[1] != 0) throw ApiError(GEORDD_ERR_ARG, "solve_general: matrix is singular");   // SingularException
    } catch (...) {
        dfree(dA); dfree(dB);
        throw;
    }
    dfree(dA); dfree(dB);
    API_END
}

int geordd_weight_at_units(geordd_gp* gp, const double* Xb, int nb, const double* w_border, double* w_unit) {
    if (!gp) return GEORDD_ERR_ARG;
    API_BEGIN(gp->ctx)
    if (!Xb || !w_border || !w_unit || nb <= 0) throw ApiError(GEORDD_ERR_ARG, "weight_at_units: bad arguments");
    const int nbp = round_up(nb, TILE), mp = gp->npad;
    double *dXb = nullptr, *KXb = nullptr, *w = nullptr, *out = nullptr;
    try {
        dXb = dalloc((size_t)gp->dim * nb, c__->stream, false);
        KXb = dalloc((size_t)mp * nbp, c__->stream, false);
        w = dalloc((size_t)nbp * 64, c__->stream, true);
        out = dalloc((size_t)mp * 64, c__->stream, true);
        GD_CUDA(cudaMemcpyAsync(dXb, Xb, (size_t)gp->dim * nb * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        double sw = 0.0;
        for (int i = 0; i < nb; ++i) sw += w_border[i];
        GD_CUDA(cudaMemcpyAsync(w, w_border, (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
        {
            ProfScope ps(c__, PC_COV);
            launch_cov(c__->stream, to_dev(gp->k), gp->dX, gp->n, dXb, nb, gp->dim, false, false, 0.0, KXb, mp, mp, nbp, false);
        }
        SolveEpilogue none;
        pd_solve(c__, gp->dL, mp, KXb, mp, nbp, none, nb);      // K_XX \ K_Xb
        gemm(c__, true, false, mp, 64, nbp, 1.0 / sw, KXb, mp, w, nbp, 0.0, out, mp, 1);
        GD_CUDA(cudaMemcpyAsync(w_unit, out, (size_t)gp->n * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
        check_abort(c__);
    } catch (...) {
        dfree(dXb); dfree(KXb); dfree(w); dfree(out);
        throw;
    }
    dfree(dXb); dfree(KXb); dfree(w); dfree(out);
    API_END
}

// The 3-argument pval_invvar_calib (src/hypotest_invvar.jl:74-105) on device data: Sigma_b = dSig + nugget I solved by LU
// (generic `\\` on a Matrix), cov_mu_tau = sum((Sigma_b \\ cov_mu_delta) * (Sigma_b \\ ones)), numerator sum(Sigma_b \\ mu_b).
// dSig is consumed.  dMu may be NULL (numerator not wanted).
static void invvar_calib_lu(geordd_gp* T, geordd_gp* C, const double* dXb, int nb, int nbp, double* dSig, double nugget,
                            const double* dMu, double* cov_mu_tau, double* numer) {
    geordd_ctx* c = T->ctx;
    // rhs = [cov_mu_delta (nb cols) | ones | mu_b]  solved with LU of (Sigma_b + nugget I)
    double* rhs = dalloc((size_t)nbp * (nbp + 2), c->stream, true);
    try {
        calib_cov_mu_delta(T, C, dXb, nb, nbp, false, rhs);
        launch_fill(c->stream, rhs + (size_t)nbp * nb, nb, 1.0);
        if (dMu) GD_CUDA(cudaMemcpyAsync(rhs + (size_t)nbp * (nb + 1), dMu, (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice,
                                         c->stream));
        lu_solve_dev(c, dSig, nb, nbp, nugget, rhs, nb + 2);
        std::vector<double> hr((size_t)nb * (nb + 2));
        download2d(c, hr.data(), nb, rhs, nbp, nb, nb + 2);
        check_abort(c);
        // (tiny: nb^2 flops) row sums of (Sigma\\cmd) * (Sigma\\1), then the total
        const double* s1 = hr.data() + (size_t)nb * nb;
        const double* sm = hr.data() + (size_t)nb * (nb + 1);
        double cov_mt = 0.0, num = 0.0;
        for (int i = 0; i < nb; ++i) {
            double row = 0.0;
            for (int j = 0; j < nb; ++j) row += hr[(size_t)j * nb + i] * s1[j];
            cov_mt += row;
        }
        for (int i = 0; i < nb; ++i) num += sm[i];
        *cov_mu_tau = cov_mt;
        if (numer) *numer = num;
    } catch (...) {
        dfree(rhs);
        throw;
    }
    dfree(rhs);
}

int geordd_invvar_calib(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, double* pval) {
    if (!T) return GEORDD_ERR_ARG;
    API_BEGIN(T->ctx)
    check_pair(T, C);
    if (!pval) throw ApiError(GEORDD_ERR_ARG, "invvar_calib: output is NULL");
    CliffTmp t;
    cliff_tmp(T, C, Xb, nb, t);
    double cov_mt = 0.0, num = 0.0;
    invvar_calib_lu(T, C, t.dXb, nb, t.nbp, t.dSig, nugget, t.dMu, &cov_mt, &num);
    double z = std::fabs(num) / std::sqrt(cov_mt);
    *pval = 2.0 * (std::erfc(z * 0.7071067811865476) / 2.0);
    API_END
}

static int null_set_sentinels(geordd_null* h, const double* Xb, int nb, double nugget) {
    geordd_ctx* c = h->ctx;
    geordd_gp *T = h->T, *C = h->C;
    check_pair(T, C);
    if (!Xb || nb <= 0) throw ApiError(GEORDD_ERR_ARG, "sentinels missing");
    dfree(h->dXb); dfree(h->dKbT); dfree(h->dKbC); dfree(h->dSig); dfree(h->dSigRaw); dfree(h->dLSig); dfree(h->dOnes); dfree(h->dMuObs);
    h->dXb = h->dKbT = h->dKbC = h->dSig = h->dSigRaw = h->dLSig = h->dOnes = h->dMuObs = nullptr;
    h->calib_lu_ready = false;
    if (round_up(nb, TILE) != h->nbp) {   // sentinel-sized simulation buffers must be re-made
        h->M.release(); h->W.release(); h->partChi.release(); h->partNum.release();
        h->cap_cols = 0;
    }
    h->nb = 0;
    h->calib_ready[0] = h->calib_ready[1] = false;
    h->nugget = nugget;
    const int nbp = round_up(nb, TILE);
    h->nbp = nbp;
    h->dXb = dalloc((size_t)2 * nb, c->stream, false);
    GD_CUDA(cudaMemcpyAsync(h->dXb, Xb, (size_t)2 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    h->dSig = dalloc((size_t)nbp * nbp, c->stream, true);
    h->dLSig = dalloc(factor_elems(nbp), c->stream, true);
    // the cliff face of the observed outcomes: Sigma_b for the simulations, mu_b kept for the observed statistics
    h->dMuObs = dalloc((size_t)nbp * 64, c->stream, true);
    cliff_face_dev(T, C, h->dXb, nb, nbp, h->dMuObs, h->dSig);
    if (nugget != 0.0) launch_add_diag(c->stream, h->dSig, nbp, nb, nugget);
    // keep Sigma + nugget I as it is BEFORE make_posdef! may jitter it: the analytic calibration solves against it by LU
    h->dSigRaw = dalloc((size_t)nbp * nbp, c->stream, false);
    GD_CUDA(cudaMemcpyAsync(h->dSigRaw, h->dSig, (size_t)nbp * nbp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    int info = make_posdef(c, h->dSig, h->dLSig, nbp, nb, nbp, &h->jitter_rounds);
    if (info != 0) {
        c->err = "cliff-face covariance is not positive definite after 10 jitter rounds";
        return info;
    }
    h->dKbT = dalloc((size_t)nbp * T->npad, c->stream, false);
    h->dKbC = dalloc((size_t)nbp * C->npad, c->stream, false);
    {
        ProfScope ps(c, PC_COV);
        launch_cov(c->stream, to_dev(T->k), h->dXb, nb, T->dX, T->n, 2, false, false, 0.0, h->dKbT, nbp, nbp, T->npad, false);
        launch_cov(c->stream, to_dev(C->k), h->dXb, nb, C->dX, C->n, 2, false, false, 0.0, h->dKbC, nbp, nbp, C->npad, false);
    }
    h->dOnes = dalloc((size_t)nbp, c->stream, true);
    launch_fill(c->stream, h->dOnes, nb, 1.0);
    // denom = sum(PDcliff \ ones)
    double* slab = dalloc((size_t)nbp * 64, c->stream, true);
    try {
        GD_CUDA(cudaMemcpyAsync(slab, h->dOnes, (size_t)nbp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        SolveEpilogue none;
        pd_solve(c, h->dLSig, nbp, slab, nbp, 64, none, 1);
        launch_reduce(c->stream, 1, slab, nullptr, 0, nb, 0, c->d_scal + 2);
        check_abort(c);
        h->denom = read_scalar(c, c->d_scal + 2);
    } catch (...) {
        dfree(slab);
        throw;
    }
    dfree(slab);
    h->nb = nb;
    return 0;
}

int geordd_null_set_sentinels(geordd_null* h, const double* Xb, int nb, double nugget, int* jitter_rounds) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    int info = null_set_sentinels(h, Xb, nb, nugget);
    if (jitter_rounds) *jitter_rounds = h->jitter_rounds;
    if (info != 0) return info;
    API_END
}

int geordd_null_prepare(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, geordd_null** out,
                        int* jitter_rounds, double* mll_null) {
    if (out) *out = nullptr;
    if (!T) return GEORDD_ERR_ARG;
    API_BEGIN(T->ctx)
    if (!out) throw ApiError(GEORDD_ERR_ARG, "null_prepare: out is NULL");
    if (!C || T->ctx != C->ctx || T->dim != C->dim) throw ApiError(GEORDD_ERR_ARG, "null_prepare: incompatible gpT / gpC");
    geordd_null* h = new geordd_null();
    int info = 0;
    try {
        h->ctx = c__; h->T = T; h->C = C; h->nugget = nugget;
        // make_null: pooled GP with T's kernel, mean and noise (src/hypotest.jl:1-12)
        const int n = T->n + C->n, dim = T->dim;
        std::vector<double> X((size_t)dim * n), y(n);
        std::copy(T->X.begin(), T->X.end(), X.begin());
        std::copy(C->X.begin(), C->X.end(), X.begin() + (size_t)dim * T->n);
        std::copy(T->y.begin(), T->y.end(), y.begin());
        std::copy(C->y.begin(), C->y.end(), y.begin() + T->n);
        h->N = gp_create(c__, &T->k, X.data(), dim, n, y.data(), T->log_noise, &info, T);
        if (!h->N) {
            null_free(h);
            c__->err = "pooled null covariance is not positive definite";
            return info;
        }
        if (mll_null) *mll_null = h->N->mll;
        if (Xb && nb > 0) {
            info = null_set_sentinels(h, Xb, nb, nugget);
            if (jitter_rounds) *jitter_rounds = h->jitter_rounds;
            if (info != 0) {
                null_free(h);
                return info;
            }
        } else if (jitter_rounds) {
            *jitter_rounds = 0;
        }
        GD_CUDA(cudaStreamSynchronize(c__->stream));
    } catch (...) {
        null_free(h);
        throw;
    }
    *out = h;
    API_END
}

// Access to the pooled null GP of a handle (alpha / mll / y as fitted by make_null).
int geordd_null_gp(geordd_null* h, geordd_gp** gp) {
    if (!h || !gp) return GEORDD_ERR_ARG;
    *gp = h->N;
    return 0;
}

// prior_rand(gp) for nsim draws (src/hypotest.jl:13-18): Y (n x nsim) = m + L Z.
int geordd_gp_prior_rand(geordd_gp* gp, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                         double* Y) {
    if (!gp) return GEORDD_ERR_ARG;
    API_BEGIN(gp->ctx)
    if (!Y || nsim <= 0) throw ApiError(GEORDD_ERR_ARG, "prior_rand: bad arguments");
    const int n = gp->n, np = gp->npad;
    const int ncols = (int)((nsim + 63) / 64 * 64);
    double* dZ = dalloc(zb_elems(np, ncols), c__->stream, true);
    double* dY = nullptr;
    try {
        dY = dalloc((size_t)np * ncols, c__->stream, true);
        if (Z) {
            c__->tmp2.ensure((size_t)n * nsim, c__->stream, false);
            GD_CUDA(cudaMemcpyAsync(c__->tmp2.p, Z, (size_t)n * nsim * sizeof(double), cudaMemcpyHostToDevice, c__->stream));
            launch_zb_pack(c__->stream, c__->tmp2.p, n, n, nsim, dZ, np, ncols);
        } else {
            ProfScope ps(c__, PC_RNG);
            launch_philox_normals(c__->stream, seed, draw0, ncols, n, dZ, 0, np, true);
        }
        SolveEpilogue e;
        e.nrows = n; e.split = n; e.add_const = gp->k.mean_const; e.out0 = dY; e.ld0 = np;
        {
            ProfScope ps(c__, PC_TRMM);
            launch_solve(c__->stream, c__->sched, SOLVE_TRMM, gp->dL, np, dZ, np, ncols, e);
        }
        download2d(c__, Y, n, dY, np, n, (int)nsim);
        check_abort(c__);
    } catch (...) {
        dfree(dZ); dfree(dY);
        throw;
    }
    dfree(dZ); dfree(dY);
    API_END
}

int geordd_null_chi2(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                     double chi_obs, double* stats, long long* n_exceed) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    SimRequest rq;
    rq.want_cliff = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z; rq.chi_obs = chi_obs;
    rq.pval_obs = -1.0; rq.chi = stats;
    run_sims(h, rq);
    if (n_exceed) *n_exceed = rq.n_chi;
    API_END
}

int geordd_null_invvar(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                       double pval_obs, double* pvals, long long* n_below) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    SimRequest rq;
    rq.want_cliff = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z; rq.pval_obs = pval_obs;
    rq.chi_obs = INFINITY; rq.pval = pvals;
    run_sims(h, rq);
    if (n_below) *n_below = rq.n_inv;
    API_END
}

int geordd_null_mll(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                    double dmll_obs, double* mll_null, double* mll_alt, long long* n_exceed, double* last_y,
                    double* last_alpha) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    SimRequest rq;
    rq.want_mll = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z; rq.dmll_obs = dmll_obs;
    rq.mll_null = mll_null; rq.mll_alt = mll_alt; rq.last_y = last_y; rq.last_alpha = last_alpha;
    run_sims(h, rq);
    if (n_exceed) *n_exceed = rq.n_mll;
    API_END
}

int geordd_null_all(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                    double tau, double* chi, double* pvals, double* mll_null, double* mll_alt) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    SimRequest rq;
    rq.want_cliff = h->nb > 0; rq.want_mll = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z; rq.tau = tau;
    rq.chi_obs = INFINITY; rq.pval_obs = -1.0; rq.dmll_obs = INFINITY;
    rq.chi = chi; rq.pval = pvals; rq.mll_null = mll_null; rq.mll_alt = mll_alt;
    run_sims(h, rq);
    API_END
}

int geordd_power(geordd_null* h, double tau, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                 const double* chi_null, const double* mll_null, const double* pinv_null, long nnull, int compat_calib_bug,
                 double* out5) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    if (!chi_null || !mll_null || !pinv_null || nnull <= 0 || !out5) throw ApiError(GEORDD_ERR_ARG, "power: bad arguments");
    if (h->nb <= 0) throw ApiError(GEORDD_ERR_ARG, "power: null handle prepared without sentinels");
    cudaStream_t st = c__->stream;
    SimRequest rq;
    rq.want_cliff = true; rq.want_mll = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z; rq.tau = tau;
    rq.chi_obs = INFINITY; rq.pval_obs = -1.0; rq.dmll_obs = INFINITY;
    std::vector<double> num(nsim), pobs(nsim);
    rq.num = num.data(); rq.pval = pobs.data();
    run_sims(h, rq);
    const double cov_mt = null_cov_mu_tau(h, compat_calib_bug != 0);
    // device-side p-values against the precomputed null vectors (src/power.jl:30-40)
    double* ref = dalloc((size_t)3 * nnull, st, false);
    double* dm = nullptr; double* o = nullptr;
    try {
        dm = dalloc((size_t)nsim, st, false);
        o = dalloc((size_t)3 * nsim, st, false);
        GD_CUDA(cudaMemcpyAsync(ref, mll_null, (size_t)nnull * sizeof(double), cudaMemcpyHostToDevice, st));
        GD_CUDA(cudaMemcpyAsync(ref + nnull, chi_null, (size_t)nnull * sizeof(double), cudaMemcpyHostToDevice, st));
        GD_CUDA(cudaMemcpyAsync(ref + 2 * nnull, pinv_null, (size_t)nnull * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_axpby(st, nsim, 1.0, h->oAlt.p, 0.0, dm);
        launch_axpby(st, nsim, -1.0, h->oNull.p, 1.0, dm);                   // dmll = alt - null
        launch_frac_compare(st, ref, nnull, dm, nsim, true, o);              // mean(mll_null .> dmll)
        launch_frac_compare(st, ref + nnull, nnull, h->oChi.p, nsim, true, o + nsim);       // mean(chi_null .> chi2)
        launch_frac_compare(st, ref + 2 * nnull, nnull, h->oPval.p, nsim, false, o + 2 * nsim);   // mean(pnull .< pobs)
        std::vector<double> ho((size_t)3 * nsim);
        GD_CUDA(cudaMemcpyAsync(ho.data(), o, (size_t)3 * nsim * sizeof(double), cudaMemcpyDeviceToHost, st));
        check_abort(c__);
        const double sd = std::sqrt(cov_mt);
        for (long s = 0; s < nsim; ++s) {
            out5[5 * s + 0] = ho[s];
            out5[5 * s + 1] = ho[nsim + s];
            out5[5 * s + 2] = pobs[s];
            out5[5 * s + 3] = ho[2 * nsim + s];
            double z = std::fabs(num[s]) / sd;
            out5[5 * s + 4] = 2.0 * (std::erfc(z * 0.7071067811865476) / 2.0);
        }
    } catch (...) {
        dfree(ref); dfree(dm); dfree(o);
        throw;
    }
    dfree(ref); dfree(dm); dfree(o);
    API_END
}

int geordd_null_invvar_calib(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                             double* pvals) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    if (!pvals) throw ApiError(GEORDD_ERR_ARG, "null_invvar_calib: output is NULL");
    if (h->nb <= 0) throw ApiError(GEORDD_ERR_ARG, "null_invvar_calib: null handle prepared without sentinels");
    // Per draw the reference evaluates the 3-argument pval_invvar_calib(gpT, gpC, Xb) (src/hypotest_invvar.jl:138-147):
    // only mu_b changes between draws; Sigma_b, its LU and cov_mu_tau are recomputed there to the same values every time
    // and are computed ONCE here, by the same LU solves.
    if (!h->calib_lu_ready) {
        const size_t sq = (size_t)h->nbp * h->nbp;
        double* sig = dalloc(sq, c__->stream, false);
        try {
            GD_CUDA(cudaMemcpyAsync(sig, h->dSigRaw, sq * sizeof(double), cudaMemcpyDeviceToDevice, c__->stream));
            invvar_calib_lu(h->T, h->C, h->dXb, h->nb, h->nbp, sig, 0.0 /* nugget already inside dSigRaw */, nullptr,
                            &h->cov_mu_tau_lu, nullptr);
        } catch (...) {
            dfree(sig);
            throw;
        }
        dfree(sig);
        h->calib_lu_ready = true;
    }
    SimRequest rq;
    rq.want_cliff = true; rq.want_lunum = true; rq.nsim = nsim; rq.seed = seed; rq.draw0 = draw0; rq.Z = Z;
    rq.chi_obs = INFINITY; rq.pval_obs = -1.0;
    std::vector<double> num(nsim);
    rq.numlu = num.data();
    run_sims(h, rq);
    const double sd = std::sqrt(h->cov_mu_tau_lu);
    for (long s = 0; s < nsim; ++s) pvals[s] = 2.0 * (std::erfc(std::fabs(num[s]) / sd * 0.7071067811865476) / 2.0);
    API_END
}

int geordd_null_destroy(geordd_null* h) {
    if (!h) return 0;
    try { geordd::ctx_use(h->ctx); } catch (...) { return GEORDD_ERR_CUDA; }
    cudaStreamSynchronize(h->ctx->stream);
    geordd::null_free(h);
    return 0;
}

}  // extern "C"
