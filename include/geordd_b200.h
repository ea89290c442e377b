/* geordd_b200.h -- C ABI of libgeordd_b200.so: the B200-native (sm_100a) implementation of the
 * GeoRDD.jl data-parallel hot path.
 *
 * The reference (maximerischard/GeoRDD.jl) has no FFI boundary: the hot path is ordinary Julia
 * calls into GaussianProcesses.jl / PDMats / LinearAlgebra.  This header is the boundary the
 * reference-side wrappers `ccall` into (see INTEGRATION.md and julia/GeoRDDB200.jl); each entry
 * point cites the reference function whose body it replaces.
 *
 * Conventions
 *   - All array arguments are caller-owned HOST memory, FP64, column-major (Julia layout):
 *     coordinates are dim x n (unit i = column i), sentinels 2 x nb, D is p x n.
 *     The library copies in/out and never retains a host pointer after return.
 *   - Device state lives in opaque handles freed by the matching *_destroy.
 *   - Return value: 0 ok; > 0 LAPACK-style index of the first non-positive pivot once the
 *     make_posdef! jitter budget (10 rounds) is exhausted (wrapper throws PosDefException(info),
 *     keeping src/GPrealisationsCovars.jl:239-241,261-263 working); < 0 argument / CUDA error, text
 *     from geordd_last_error().
 *   - One context = one GPU = one CUDA stream; a context is not re-entrant.  Multi-GPU runs use one
 *     process (context) per GPU; null draws are sharded by draw index (`draw0`), results depend only
 *     on (seed, draw index) and the integer tallies are summed by the caller (NCCL all-reduce).
 *   - There is NO CPU fallback: every entry point fails with GEORDD_ERR_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef GEORDD_B200_H
#define GEORDD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GEORDD_API __attribute__((visibility("default")))
#else
#define GEORDD_API
#endif

#define GEORDD_ERR_ARG (-1)
#define GEORDD_ERR_CUDA (-2)
#define GEORDD_ERR_WATCHDOG (-3)
#define GEORDD_ERR_NOMEM (-4)
#define GEORDD_ERR_NCCL (-5)

typedef struct geordd_ctx geordd_ctx;
typedef struct geordd_gp geordd_gp;
typedef struct geordd_null geordd_null;
typedef struct geordd_multigp geordd_multigp;
typedef struct geordd_group geordd_group;   /* several GPUs driven by one host process */
typedef struct geordd_ggp geordd_ggp;       /* a fitted GP replicated on every device of a group */
typedef struct geordd_gnull geordd_gnull;   /* a null handle replicated on every device of a group */

/* Isotropic spatial kernel (+ optional Const kernel) and constant mean.
 * kind: 0 SEIso, 1 Mat12Iso, 2 Mat32Iso, 3 Mat52Iso (GaussianProcesses.jl kernels);
 * ell: length scale (sqrt(l2) for SEIso); sigma2: signal variance;
 * const_sigma2: variance of a summed Const kernel (0 = none); mean_const: MeanConst (0 = MeanZero). */
typedef struct {
    int kind;
    double ell, sigma2, const_sigma2, mean_const;
} geordd_kernel;

GEORDD_API int geordd_version(void);
/* Total number of CUDA kernels this library has launched in this process. */
GEORDD_API long long geordd_launch_count(void);

/* ---- context ------------------------------------------------------------------------------- */
GEORDD_API int geordd_ctx_create(geordd_ctx** out, int device);
GEORDD_API int geordd_ctx_destroy(geordd_ctx* ctx);
GEORDD_API const char* geordd_last_error(const geordd_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (e.g. the caller's current torch stream)
 * so that the caller's CUDA events bracket it.  NULL restores the context's own stream. */
GEORDD_API int geordd_ctx_set_stream(geordd_ctx* ctx, void* cuda_stream);
GEORDD_API int geordd_ctx_sync(geordd_ctx* ctx);
/* Per-kernel-class CUDA-event timing (on the launching stream).  enable != 0 starts/clears it.
 * geordd_ctx_profile_get returns accumulated milliseconds and launch count for a kernel class
 * ("cov", "potrf", "solve_fwd", "solve_bwd", "solve_seq", "trmm", "gemm", "rng", "finish", "lu", "misc"),
 * or for "*" the total over all classes. */
GEORDD_API int geordd_ctx_profile(geordd_ctx* ctx, int enable);
GEORDD_API int geordd_ctx_profile_get(geordd_ctx* ctx, const char* name, double* ms, long long* launches);
/* Upper bound on null draws processed per internal batch (0 = automatic). */
GEORDD_API int geordd_ctx_set_max_batch(geordd_ctx* ctx, long max_batch);
/* Opt-in (default 0): evaluate the quadratic forms of geordd_null_mll without the backward solves,
 *   (y-m)' K^-1 (y-m) = |L^-1 (y-m)|^2   and, for the pooled null GP whose draw is y = m + L z,   = z'z,
 * instead of the reference's stepwise alpha = cK \ (y-m); dot(y-m, alpha) (src/gp_utils.jl:15-22,
 * src/hypotest_mll.jl:1-19).  Same statistic in exact arithmetic (rounding-level differences, ~1e-12 relative), 6 m^2
 * instead of 16 m^2 flop per draw.  Ignored when the caller asks for the final draw's y / alpha. */
GEORDD_API int geordd_ctx_set_mll_forward_only(geordd_ctx* ctx, int enable);
/* Certificate for the exact exceedance counts of the LAST geordd_null_* / geordd_power call on this context: the smallest
 * |simulated statistic - observed threshold| over all draws, for (chi2, inverse-variance p-value, delta-mll); +Inf where the
 * statistic was not simulated or the threshold was infinite.  A count can only differ from the reference's if the
 * statistics disagree by more than this margin (SURVEY §5, §7 hard part 2). */
GEORDD_API int geordd_ctx_last_min_margin(const geordd_ctx* ctx, double* margin3);

/* ---- K1 covariance assembly: cov(kernel, X, X) + diag_add*I and cov(kernel, X1, X2) ----------
 * replaces GaussianProcesses.cov / cov! (call sites src/hypotest_chi2.jl:38-39, src/power.jl:60-62,
 * src/hypotest_invvar.jl:80-86) and addcov!/add_diag! (src/gp_utils.jl:24-44).  K: n x n / n1 x n2. */
GEORDD_API int geordd_cov_sym(geordd_ctx* ctx, const geordd_kernel* k, const double* X, int dim, int n, double diag_add,
                   double* K);
GEORDD_API int geordd_cov_cross(geordd_ctx* ctx, const geordd_kernel* k, const double* X1, int n1, const double* X2, int n2,
                     int dim, double* K);

/* ---- a1: GPE(x, y, mean, kernel, logNoise) (src/region_gp_fit.jl:12,28; src/hypotest.jl:4) ----
 * K = k(X,X) + (exp(2 logNoise) + eps) I; make_posdef! (jitter retries); alpha = K \ (y - m);
 * mll (src/gp_utils.jl:19-22).  alpha (n), mll (1), cholU (n x n upper factor, nullable),
 * jitter_rounds (nullable) are outputs. */
GEORDD_API int geordd_gp_fit(geordd_ctx* ctx, const geordd_kernel* k, const double* X, int dim, int n, const double* y,
                  double log_noise, geordd_gp** out, double* alpha, double* mll, double* cholU, int* jitter_rounds);
/* a6: the fits of GPRealisations(rdict, groupKeys, kernel, logNoise, mean) (src/GPrealisations.jl:10-19, reached from
 * src/region_gp_fit.jl:23-37): `nfit` independent GPE fits in one call (also the two halves of a placebo split,
 * src/hypotest_chi2.jl:66-69).  All covariance matrices are assembled, then factored by ONE dataflow launch in which the
 * critical paths of the factorisations interleave; results are identical to nfit geordd_gp_fit calls, bit for bit.
 * X[f]: dim x n[f]; y[f]: n[f]; k, log_noise: nfit entries; outputs out[f], alpha[f] (nullable, n[f]), mll[f],
 * jitter_rounds[f], info[f] (nullable).  If any fit fails, every handle is released and the first LAPACK info is returned. */
GEORDD_API int geordd_gp_fit_batch(geordd_ctx* ctx, int nfit, const geordd_kernel* k, const double* const* X, int dim, const int* n,
                        const double* const* y, const double* log_noise, geordd_gp** out, double* const* alpha, double* mll,
                        int* jitter_rounds, int* info);
/* a9: gp.y = y; update_alpha!(gp); mll(gp) with the frozen factor (src/gp_utils.jl:15-22). */
GEORDD_API int geordd_gp_set_y(geordd_gp* gp, const double* y, double* alpha, double* mll);
/* predict_mu(gp, Xb, cK) = m(Xb) + cov(kernel, Xb, gp.x) * alpha (src/gp_utils.jl:1-5). */
GEORDD_API int geordd_gp_predict_mu(geordd_gp* gp, const double* Xb, int nb, double* mu);
/* cK.mat (jittered covariance incl. noise), n x n. */
GEORDD_API int geordd_gp_get_K(geordd_gp* gp, double* K);
GEORDD_API int geordd_gp_nobs(const geordd_gp* gp);
GEORDD_API int geordd_gp_destroy(geordd_gp* gp);

/* ---- a7: cliff_face(gpT, gpC, sentinels) (src/cliff_face.jl:3-9): mu (nb), Sigma (nb x nb) ---- */
GEORDD_API int geordd_cliff_face(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double* mu, double* Sigma);

/* ---- a11 (LATE): inverse_variance(mu, Sigma; nugget) (src/average_treatment_effect.jl:3-11),
 * Sigma a plain Matrix => + nugget*I and LU solves.  Also unweighted/weighted means (:13-29). */
GEORDD_API int geordd_inverse_variance(geordd_ctx* ctx, const double* mu, const double* Sigma, int nb, double nugget,
                            double* tau_hat, double* var_tau);
/* Julia's generic `A \\ B` on a dense square Matrix = lu(A) \\ B (partial pivoting; hit at
 * src/weight_at_units.jl:36, src/hypotest_chi2.jl:9).  A is n x n, B (n x nrhs) is overwritten by the solution. */
GEORDD_API int geordd_solve_general(geordd_ctx* ctx, const double* A, int n, double* B, int nrhs);
/* ---- K7 projection of units onto the border: nearest point on a (Multi)LineString and its distance ----------
 * replaces the per-unit LibGEOS calls distance(border, point) / nearestPoints(border, point)[1] of projection_points
 * (src/border_projection.jl:4-22), which feed proj_estimator (:24-31) and weights_proj (src/weight_at_units.jl:44-53).
 * X: 2 x n unit coordinates; V: 2 x nv border vertices; part_start[nparts]: vertex indices at which a new line of a
 * MultiLineString starts (vertex 0 always does; nparts may be 0).  Outputs: Xproj 2 x n, dist n; the caller keeps the
 * columns with dist <= maxdist.  GEOS 3.8 arithmetic restated operation by operation (first minimum wins). */
GEORDD_API int geordd_projection_points(geordd_ctx* ctx, const double* X, int n, const double* V, int nv, const int* part_start,
                                        int nparts, double* Xproj, double* dist);
/* LibGEOS.within(point, region) for n points (the grid filter of infinite_proj_sentinels, src/border_projection.jl:44-102;
 * the reference notes it is "stupidly slow").  R: 2 x nr vertices of all rings of the (Multi)Polygon, each ring closed
 * (first vertex repeated); ring_start[nrings]: vertex indices at which further rings start.  inside[i] = 1 / 0.
 * GEOS RayCrossingCounter restated with the orientation test in plain FP64; even-odd over all rings. */
GEORDD_API int geordd_points_in_region(geordd_ctx* ctx, const double* P, int n, const double* R, int nr, const int* ring_start,
                                       int nrings, unsigned char* inside);

/* weight_at_units(gp, sentinels, weights) (src/weight_at_units.jl:12-18): w_unit (n). */
GEORDD_API int geordd_weight_at_units(geordd_gp* gp, const double* Xb, int nb, const double* w_border, double* w_unit);

/* ---- observed statistics (LU on the raw cliff-face covariance) --------------------------------
 * chistat(gpT,gpC,Xb) (src/hypotest_chi2.jl:4-10); pval_invvar_uncalib(gpT,gpC,Xb;nugget)
 * (src/hypotest_invvar.jl:1-6); pval_invvar_calib(gpT,gpC,Xb;nugget) (src/hypotest_invvar.jl:74-105). */
GEORDD_API int geordd_chistat_obs(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double* chi2);
GEORDD_API int geordd_invvar_obs(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, double* tau_hat,
                      double* var_tau, double* pval);
GEORDD_API int geordd_invvar_calib(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, double* pval);

/* ---- a8/a10/a11/a13/a14: Monte-Carlo sharp-null machinery -------------------------------------
 * geordd_null_prepare = the draw-independent setup of nsim_chi (src/hypotest_chi2.jl:33-48),
 * nsim_invvar_pval_uncalib (src/hypotest_invvar.jl:48-59) and nsim_power (src/power.jl:53-66):
 * make_null (pooled fit, src/hypotest.jl:1-12), cliff face, make_posdef!(Sigma + nugget I),
 * cK_T/cK_C = cov(kernel, Xb, gp.x).  Xb may be NULL with nb = 0 for the mll test alone.
 * T and C must outlive the handle.
 *
 * Draws: Z == NULL  -> device Philox-4x32-10 / Box-Muller normals keyed by (seed, draw0 + s, element);
 *        Z != NULL  -> host normals, n x nsim column-major, column s = the randn(n) of draw s
 *                      (exactly the order rand(MvNormal) consumes them, src/hypotest.jl:13-18).
 * Every draw: y = m + L_N z (prior_rand); alpha_T, alpha_C (update_alpha!); statistic. */
GEORDD_API int geordd_null_prepare(geordd_gp* T, geordd_gp* C, const double* Xb, int nb, double nugget, geordd_null** out,
                        int* jitter_rounds, double* mll_null);
/* Re-run only the sentinel-dependent part of the setup on an existing handle (new Xb and/or nugget). */
GEORDD_API int geordd_null_set_sentinels(geordd_null* h, const double* Xb, int nb, double nugget, int* jitter_rounds);
/* chistat(gpT, gpC, Xb) (src/hypotest_chi2.jl:4-11) and pval_invvar_uncalib(gpT, gpC, Xb; nugget) (src/hypotest_invvar.jl:1-8)
 * of the OBSERVED outcomes from a handle whose sentinels are set: the cliff face evaluated there is reused instead of
 * being evaluated again (boot_chi2test :54-59, boot_invvar :65-69 call both).  Bit-identical to geordd_chistat_obs /
 * geordd_invvar_obs with the handle's sentinels and nugget; geordd_null_chistat_obs requires nugget == 0 (chistat has none). */
GEORDD_API int geordd_null_chistat_obs(geordd_null* h, double* chi2);
GEORDD_API int geordd_null_invvar_obs(geordd_null* h, double* tau_hat, double* var_tau, double* pval);
/* Borrow the pooled null GP (make_null result; owned by the handle). */
GEORDD_API int geordd_null_gp(geordd_null* h, geordd_gp** gp);
/* a8: prior_rand(gp) for nsim draws (src/hypotest.jl:13-18): Y (n x nsim) = m + L z. Z as below. */
GEORDD_API int geordd_gp_prior_rand(geordd_gp* gp, long nsim, unsigned long long seed, unsigned long long draw0,
                         const double* Z, double* Y);
/* nsim_chi / boot_chi2test (src/hypotest_chi2.jl:19-59): stats (nsim, nullable), n_exceed = #{chi > chi_obs}. */
GEORDD_API int geordd_null_chi2(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                     double chi_obs, double* stats, long long* n_exceed);
/* nsim_invvar_pval_uncalib / boot_invvar (src/hypotest_invvar.jl:11-69): pvals, n_below = #{p < pval_obs}. */
GEORDD_API int geordd_null_invvar(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                       double pval_obs, double* pvals, long long* n_below);
/* sim_invvar_calib! / nsim_invvar_calib (src/hypotest_invvar.jl:138-160): the 3-argument pval_invvar_calib (:74-105) of
 * every null draw, 2 ccdf(Normal(0, sqrt(cov_mu_tau)), |sum(Sigma_b \ mu_b)|), with Sigma_b the un-jittered cliff covariance
 * + nugget solved by LU exactly as the reference's generic `\` does.  Only mu_b depends on the draw: the LU of Sigma_b is
 * formed once per CTA and cov_mu_tau (:88-99) once per handle (the reference recomputes both per draw, to the same values).
 * Prepare the handle with nugget = 1e-10 (the reference's default). */
GEORDD_API int geordd_null_invvar_calib(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0,
                             const double* Z, double* pvals);
/* sim_logP! / nsim_logP / boot_mlltest (src/hypotest_mll.jl:1-45): mll_null, mll_alt (nsim, nullable),
 * n_exceed = #{(alt - null) > dmll_obs}.  last_y / last_alpha (n, nullable) receive the pooled GP's y and
 * alpha after the final draw (the reference mutates gpNull, src/hypotest_mll.jl:6,10,14). */
GEORDD_API int geordd_null_mll(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                    double dmll_obs, double* mll_null, double* mll_alt, long long* n_exceed, double* last_y,
                    double* last_alpha);
/* All three statistics from the same draws in one pass (used by power.jl-style studies and the bench). */
GEORDD_API int geordd_null_all(geordd_null* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                    double tau, double* chi, double* pvals, double* mll_null, double* mll_alt);
/* nsim_power (src/power.jl:1-73): out5 is 5 x nsim column-major: (pval_mll, pval_chi2, pval_invvar_obs,
 * invvar_bootcalib, invvar_calib) per draw.  compat_calib_bug != 0 reproduces the as-shipped
 * 7-argument pval_invvar_calib (missing '+', src/hypotest_invvar.jl:125-128). */
GEORDD_API int geordd_power(geordd_null* h, double tau, long nsim, unsigned long long seed, unsigned long long draw0,
                 const double* Z, const double* chi_null, const double* mll_null, const double* pinv_null, long nnull,
                 int compat_calib_bug, double* out5);
GEORDD_API int geordd_null_destroy(geordd_null* h);

/* ---- a3/a4/a5: MultiGPCovars numerics (src/GPrealisationsCovars.jl) ---------------------------
 * create: uploads D (p x n), X (dim x n, groups contiguous), group offsets (ngroups+1), y; builds
 * D'D once (KernelData(LinIso, D, D), :42).  mll = update_mll! (:110-125): cK = D'D/lin_ell2 +
 * blockdiag(K_g) + max(exp(2 logNoise),1e-8) I -> make_posdef! -> alpha -> mll.
 * mll_grad = update_mll_and_dmll! (:145-193); dmll order [logNoise][mean][log ell, log sigma,
 * (log sigma_const)][LinIso ll], entries selected by the noise/domean/kern/beta flags.
 * postmean_beta = postmean_beta (:335-344). */
GEORDD_API int geordd_multigp_create(geordd_ctx* ctx, const double* D, int p, const double* X, int dim, const int* group_off,
                          int ngroups, const double* y, geordd_multigp** out);
GEORDD_API int geordd_multigp_mll(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise, double* alpha,
                       double* mll, int* jitter_rounds);
GEORDD_API int geordd_multigp_mll_grad(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise, int noise,
                            int domean, int kern, int beta, double* mll, double* dmll, int* ndmll);
GEORDD_API int geordd_multigp_postmean_beta(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise,
                                 double* beta_hat);
GEORDD_API int geordd_multigp_destroy(geordd_multigp* h);

/* ---- multi-GPU inside the library (App. D geordd_ctx_create(ndev, devs)) -----------------------------------
 * Only the embarrassingly parallel draw loops shard (src/hypotest_chi2.jl:49-50, src/hypotest_mll.jl:27-28,
 * src/hypotest_invvar.jl:60-61, src/power.jl:67-71).  A group lets ONE host process (one Julia task) drive `ndev` GPUs:
 * fits are replicated (one host thread per device, bit-identical replicas), draws are sharded contiguously by index --
 * a draw depends only on (seed, draw index), so every result equals the single-GPU one bit for bit -- per-draw statistics
 * land in the caller's arrays, and the exceedance tallies are summed with ONE ncclAllReduce(int64) over NVLink.
 * devices == NULL selects devices 0 .. ndev-1.  NCCL is bound at run time (libnccl.so.2); a group of more than one device
 * without it fails with GEORDD_ERR_NCCL (no host-side fallback). */
GEORDD_API int geordd_group_create(geordd_group** out, int ndev, const int* devices);
GEORDD_API int geordd_group_destroy(geordd_group* g);
GEORDD_API int geordd_group_size(const geordd_group* g);
GEORDD_API const char* geordd_group_last_error(const geordd_group* g);
/* Borrow the context of device i of the group (for the single-GPU entry points: fits of other models, cliff faces ...). */
GEORDD_API geordd_ctx* geordd_group_ctx(geordd_group* g, int i);
/* geordd_gp_fit on every device of the group; alpha / mll / jitter_rounds are device 0's (all replicas are identical). */
GEORDD_API int geordd_group_gp_fit(geordd_group* g, const geordd_kernel* k, const double* X, int dim, int n, const double* y,
                                   double log_noise, geordd_ggp** out, double* alpha, double* mll, int* jitter_rounds);
GEORDD_API int geordd_group_gp_destroy(geordd_ggp* h);
GEORDD_API geordd_gp* geordd_group_gp_replica(geordd_ggp* h, int i);   /* borrowed */
/* geordd_null_prepare on every device. */
GEORDD_API int geordd_group_null_prepare(geordd_ggp* T, geordd_ggp* C, const double* Xb, int nb, double nugget, geordd_gnull** out,
                                         int* jitter_rounds, double* mll_null);
GEORDD_API int geordd_group_null_destroy(geordd_gnull* h);
GEORDD_API geordd_null* geordd_group_null_replica(geordd_gnull* h, int i);   /* borrowed */
/* The sharded forms of geordd_null_mll / _chi2 / _invvar / geordd_power: same arguments and results as on one GPU. */
GEORDD_API int geordd_group_null_mll(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                                     double dmll_obs, double* mll_null, double* mll_alt, long long* n_exceed);
GEORDD_API int geordd_group_null_chi2(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                                      double chi_obs, double* stats, long long* n_exceed);
GEORDD_API int geordd_group_null_invvar(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                                        double pval_obs, double* pvals, long long* n_below);
GEORDD_API int geordd_group_power(geordd_gnull* h, double tau, long nsim, unsigned long long seed, unsigned long long draw0,
                                  const double* Z, const double* chi_null, const double* mll_null, const double* pinv_null,
                                  long nnull, int compat_calib_bug, double* out5);
/* One process per GPU (MPI-launched Julia, torchrun): the same tally reduce on a plain context.  Rank 0 calls
 * geordd_comm_unique_id and hands the 128 bytes to the other ranks by any means (a file, MPI_Bcast); every rank then calls
 * geordd_comm_init_rank; geordd_comm_allreduce_tally sums up to 8 int64 counters in place over all ranks. */
GEORDD_API int geordd_comm_unique_id(char* id128);
GEORDD_API int geordd_comm_init_rank(geordd_ctx* ctx, int nranks, int rank, const char* id128);
GEORDD_API int geordd_comm_allreduce_tally(geordd_ctx* ctx, long long* vals, int n);
GEORDD_API int geordd_comm_destroy(geordd_ctx* ctx);

/* ---- debug / test hooks (not part of the reference-facing surface) ----------------------------- */
/* GEORDD_GUARD=1 in the environment brackets every device buffer with NaN-filled guard zones (see api_core.cu): number of
 * buffers whose guards were found overwritten when they were released. */
GEORDD_API long long geordd_dbg_guard_violations(void);
/* Test hook (process-wide, default 1): 0 routes every pivoted-LU solve through the unblocked single-pass kernel, 1 through the
 * blocked one (n <= 640); both perform the same operations in the same order, the parity tests compare them bit for bit. */
GEORDD_API int geordd_dbg_set_lu_blocked(int on);
/* Device Philox normals copied to the host: Z is n x ndraws. */
GEORDD_API int geordd_dbg_normals(geordd_ctx* ctx, unsigned long long seed, unsigned long long draw0, int ndraws, int n,
                       double* Z);
/* C = alpha*op(A)*op(B) + beta*C through the DMMA mainloop (host matrices, column-major). transa/transb:
 * 0 = as stored, 1 = transposed. */
GEORDD_API int geordd_dbg_gemm(geordd_ctx* ctx, int transa, int transb, int M, int N, int K, double alpha, const double* A,
                    int lda, const double* B, int ldb, double beta, double* C, int ldc, int splitk);
/* Cholesky of a host SPD matrix (n x n) through the dataflow kernel: L lower, returns info. */
GEORDD_API int geordd_dbg_potrf(geordd_ctx* ctx, const double* A, int n, double* L, double* ms);
/* Triangular ops with a host lower factor: mode 0 B <- L^-1 B, 1 B <- L^-T B, 2 B <- L B. */
GEORDD_API int geordd_dbg_trsm(geordd_ctx* ctx, int mode, const double* L, int n, double* B, int nrhs, double* ms);
/* Device-resident micro-benchmarks (synthetic SPD matrix built on device): times in ms. */
GEORDD_API int geordd_bench_potrf(geordd_ctx* ctx, int n, int reps, double* ms_best, double* ms_mean);
/* cycles10[0..3]: the four shared-memory tile operations (potrf, right solve, forward solve, backward solve; one CTA);
   [4..6]: potrf phases (diagonal block, panel, trailing update); [7..9]: right-solve phases (panel load, substitution, update) */
GEORDD_API int geordd_bench_tileops(geordd_ctx* ctx, double* cycles10);
GEORDD_API int geordd_bench_dmma_peak(geordd_ctx* ctx, double* tflops);
GEORDD_API int geordd_bench_dmma(geordd_ctx* ctx, int threads_per_cta, int ctas, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GEORDD_B200_H */
