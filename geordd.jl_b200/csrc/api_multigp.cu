// C ABI, part 3: MultiGPCovars numerics (joint fit with linear covariates), its gradient, and the
// posterior mean of the regression coefficients -- src/GPrealisationsCovars.jl:110-193,284-344.
#include "api_internal.h"
#include <algorithm>

namespace geordd {

// ---- kernels ------------------------------------------------------------------------------------------
__global__ void scale_div_kernel(long n, const double* x, double d, double* y) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] / d;
}

// Per-group reduction of  sum_ij (alpha_i alpha_j - Kinv_ij) * dK_p(r_ij)  for the kernel log-parameters
// (p = 0: log ell, 1: log sigma, 2: log sigma_const) and  sum_ij (alpha_i alpha_j - Kinv_ij) * G_ij.
// Replaces dmll_kern! scalar loops.  Block partials are written to out[blk*4 + p]
// and summed on the host in block order (deterministic).
__global__ void __launch_bounds__(256)
grad_reduce_kernel(KernelSpecDev spec, const double* __restrict__ X, int dim, int a, int ng, const double* __restrict__ alpha,
                   const double* __restrict__ Kinv, const double* __restrict__ G, long ld, double* out) {
    // one 32x32 element tile of the group block per CTA
    const int tiles = (ng + 31) / 32;
    const int ti = blockIdx.x % tiles, tj = blockIdx.x / tiles;
    const int li = threadIdx.x % 32, lj0 = threadIdx.x / 32;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const int i = ti * 32 + li;
    if (i < ng) {
        double xi[8];
        for (int d = 0; d < dim; ++d) xi[d] = X[(long)(a + i) * dim + d];
        const double ai = alpha[a + i];
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tj * 32 + lj0 + 8 * jj;
            if (j >= ng) continue;
            double r2 = 0.0;
            for (int d = 0; d < dim; ++d) {
                double df = xi[d] - X[(long)(a + j) * dim + d];
                r2 += df * df;
            }
            const double b = ai * alpha[a + j] - Kinv[(long)(a + j) * ld + a + i];
            double k, dl;
            if (spec.kind == 0) {
                k = spec.sigma2 * exp(-0.5 * r2 / (spec.ell * spec.ell));
                dl = k * r2 / (spec.ell * spec.ell);
            } else {
                double r = sqrt(r2);
                if (spec.kind == 1) {
                    k = spec.sigma2 * exp(-r / spec.ell);
                    dl = k * r / spec.ell;
                } else if (spec.kind == 2) {
                    double sv = 1.7320508075688772 * r / spec.ell, e = exp(-sv);
                    k = spec.sigma2 * (1.0 + sv) * e;
                    dl = spec.sigma2 * sv * sv * e;
                } else {
                    double sv = 2.23606797749979 * r / spec.ell, e = exp(-sv);
                    k = spec.sigma2 * (1.0 + sv + sv * sv / 3.0) * e;
                    dl = spec.sigma2 * sv * sv * (1.0 + sv) * e / 3.0;
                }
            }
            s[0] += b * dl;
            s[1] += b * 2.0 * k;
            s[2] += b * 2.0 * spec.const_sigma2;
        }
    }
    __shared__ double red[8][4];
    for (int p = 0; p < 3; ++p) {
        double v = s[p];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][p] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        out[(long)blockIdx.x * 4 + threadIdx.x] = v;
    }
}

// sum_ij (alpha_i alpha_j - Kinv_ij) G_ij over the full n x n, plus trace terms:
// out[blk*4+0] = sum B.*G, out[blk*4+1] = partial tr(B) (diagonal entries in this block's rows)
__global__ void __launch_bounds__(256)
beta_reduce_kernel(int n, const double* __restrict__ alpha, const double* __restrict__ Kinv, const double* __restrict__ G,
                   long ld, double* out) {
    const int j = blockIdx.x;   // one column per CTA
    double s = 0.0, tr = 0.0;
    const double aj = alpha[j];
    for (int i = threadIdx.x; i < n; i += 256) {
        double b = alpha[i] * aj - Kinv[(long)j * ld + i];
        s += b * G[(long)j * ld + i];
        if (i == j) tr = b;
    }
    __shared__ double red[8][2];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        tr += __shfl_down_sync(0xffffffffu, tr, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s; red[threadIdx.x >> 5][1] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v0 = 0.0, v1 = 0.0;
        for (int w = 0; w < 8; ++w) { v0 += red[w][0]; v1 += red[w][1]; }
        out[(long)j * 4 + 0] = v0;
        out[(long)j * 4 + 1] = v1;
    }
}

static void mg_free(geordd_multigp* h) {
    if (!h) return;
    dfree(h->dD); dfree(h->dX); dfree(h->dG); dfree(h->dK); dfree(h->dL); dfree(h->dB);
    dfree(h->dY); dfree(h->dR); dfree(h->dAlpha);
    delete h;
}

static void mg_check_kernel(const geordd_kernel* k, double lin_ell2) {
    if (!k || k->kind < 0 || k->kind > 3 || !(k->ell > 0.0) || !(k->sigma2 >= 0.0) || !(k->const_sigma2 >= 0.0) ||
        !std::isfinite(k->ell) || !std::isfinite(k->sigma2) || !std::isfinite(k->const_sigma2))
        throw ApiError(GEORDD_ERR_ARG, "multigp: bad kernel");
    if (!(lin_ell2 > 0.0) || !std::isfinite(lin_ell2)) throw ApiError(GEORDD_ERR_ARG, "multigp: bad LinIso l2");
}

// blockdiag(K_g) accumulated into the lower tiles of dst (which must already hold the base matrix)
static void add_spatial_blocks(geordd_multigp* h, const geordd_kernel* k, double* dst) {
    geordd_ctx* c = h->ctx;
    for (size_t g = 0; g + 1 < h->group_off.size(); ++g) {
        const int a = h->group_off[g], ng = h->group_off[g + 1] - a;
        ProfScope ps(c, PC_COV);
        launch_cov(c->stream, to_dev(*k), h->dX + (size_t)a * h->dim, ng, h->dX + (size_t)a * h->dim, ng, h->dim, true, false,
                   0.0, dst + (size_t)a * h->npad + a, h->npad, ng, ng, true);
    }
}

// update_mll! (src/GPrealisationsCovars.jl:110-125).  Returns LAPACK info (0 ok).
static int mg_update_mll(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise, double* mll,
                         int* rounds) {
    geordd_ctx* c = h->ctx;
    const int n = h->n, np = h->npad;
    const long tot = (long)np * np;
    scale_div_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(tot, h->dG, lin_ell2, h->dK); ++g_kernel_launches;   // cov!(cK, LinIso, D)
    add_spatial_blocks(h, k, h->dK);                                                                        // addcov!
    launch_add_diag(c->stream, h->dK, np, n, std::max(std::exp(2.0 * log_noise), 1e-8));                    // add_diag!
    if (np > n) launch_add_diag(c->stream, h->dK + (size_t)n * np + n, np, np - n, 1.0);
    int nr = 0;
    int info = make_posdef(c, h->dK, h->dL, np, n, np, &nr);   // update_chol! -> make_posdef!
    if (rounds) *rounds = nr;
    if (info != 0) return info;
    launch_logdiag_sum(c->stream, h->dL, np, n, c->d_scal + 1);
    h->logdet = 2.0 * read_scalar(c, c->d_scal + 1);
    // alpha = cK \ (y - mu)
    launch_axpby(c->stream, n, 1.0, h->dY, 0.0, h->dR);
    launch_add_const(c->stream, h->dR, np, n, 1, -k->mean_const);
    GD_CUDA(cudaMemsetAsync(h->dAlpha, 0, (size_t)np * 64 * sizeof(double), c->stream));
    GD_CUDA(cudaMemcpyAsync(h->dAlpha, h->dR, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    SolveEpilogue none;
    pd_solve(c, h->dL, np, h->dAlpha, np, 64, none, 1);
    launch_reduce(c->stream, 2, h->dR, h->dAlpha, 0, n, 0, c->d_scal);
    check_abort(c);
    double quad = read_scalar(c, c->d_scal);
    *mll = -quad / 2.0 - h->logdet / 2.0 - (double)n * std::log(2.0 * M_PI) / 2.0;
    return 0;
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

int geordd_multigp_create(geordd_ctx* c, const double* D, int p, const double* X, int dim, const int* group_off,
                          int ngroups, const double* y, geordd_multigp** out) {
    if (out) *out = nullptr;
    API_BEGIN(c)
    if (!out || !D || !X || !group_off || !y || p <= 0 || dim <= 0 || dim > 8 || ngroups <= 0)
        throw ApiError(GEORDD_ERR_ARG, "multigp_create: bad arguments");
    const int n = group_off[ngroups];
    if (group_off[0] != 0 || n <= 0) throw ApiError(GEORDD_ERR_ARG, "multigp_create: group_off must start at 0");
    for (int g = 0; g < ngroups; ++g)
        if (group_off[g + 1] <= group_off[g]) throw ApiError(GEORDD_ERR_ARG, "empty group");   // throw("empty group"), :86
    geordd_multigp* h = new geordd_multigp();
    try {
        h->ctx = c; h->n = n; h->npad = round_up(n, TILE); h->p = p; h->ppad = round_up(p, BK); h->dim = dim;
        h->group_off.assign(group_off, group_off + ngroups + 1);
        const size_t np = h->npad;
        h->dD = dalloc((size_t)h->ppad * np, c->stream, true);
        h->dX = dalloc((size_t)dim * n, c->stream, false);
        h->dG = dalloc(np * np, c->stream, false);
        h->dK = dalloc(np * np, c->stream, false);
        h->dL = dalloc(factor_elems(h->npad), c->stream, false);
        h->dY = dalloc(np, c->stream, true);
        h->dR = dalloc(np, c->stream, true);
        h->dAlpha = dalloc(np * 64, c->stream, true);
        upload2d(c, h->dD, h->ppad, D, p, p, n);
        GD_CUDA(cudaMemcpyAsync(h->dX, X, (size_t)dim * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GD_CUDA(cudaMemcpyAsync(h->dY, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        // G = D'D  (KernelData(LinIso, D, D), src/GPrealisationsCovars.jl:42)
        gemm(c, false, false, h->npad, h->npad, h->ppad, 1.0, h->dD, h->ppad, h->dD, h->ppad, 0.0, h->dG, np, 1);
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
    } catch (...) {
        mg_free(h);
        throw;
    }
    *out = h;
    API_END
}

int geordd_multigp_mll(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise, double* alpha,
                       double* mll, int* jitter_rounds) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    mg_check_kernel(k, lin_ell2);
    if (!std::isfinite(log_noise)) throw ApiError(GEORDD_ERR_ARG, "multigp: non-finite logNoise");
    double m = 0.0;
    int info = mg_update_mll(h, k, lin_ell2, log_noise, &m, jitter_rounds);
    if (info != 0) {
        c__->err = "joint covariance is not positive definite";
        return info;
    }
    if (mll) *mll = m;
    if (alpha) {
        GD_CUDA(cudaMemcpyAsync(alpha, h->dAlpha, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
        GD_CUDA(cudaStreamSynchronize(c__->stream));
    }
    API_END
}

int geordd_multigp_mll_grad(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise, int noise,
                            int domean, int kern, int beta, double* mll, double* dmll, int* ndmll) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    mg_check_kernel(k, lin_ell2);
    if (!dmll) throw ApiError(GEORDD_ERR_ARG, "multigp_mll_grad: dmll is NULL");
    double m = 0.0;
    int info = mg_update_mll(h, k, lin_ell2, log_noise, &m, nullptr);
    if (info != 0) {
        c__->err = "joint covariance is not positive definite";
        return info;
    }
    if (mll) *mll = m;
    const int n = h->n, np = h->npad;
    cudaStream_t st = c__->stream;
    // get_aa_invcKI!: B = alpha alpha' - cK^-1 ; cK^-1 = cK \ I  (n right-hand sides through the batched solver)
    if (!h->dB) h->dB = dalloc((size_t)np * np, st, false);
    GD_CUDA(cudaMemsetAsync(h->dB, 0, (size_t)np * np * sizeof(double), st));
    launch_add_diag(st, h->dB, np, np, 1.0);
    SolveEpilogue none;
    pd_solve(c__, h->dL, np, h->dB, np, np, none);
    // reductions
    const int ngroups = (int)h->group_off.size() - 1;
    std::vector<int> blk_off(ngroups + 1, 0);
    for (int g = 0; g < ngroups; ++g) {
        int ng = h->group_off[g + 1] - h->group_off[g];
        int t = (ng + 31) / 32;
        blk_off[g + 1] = blk_off[g] + t * t;
    }
    const int nblk = blk_off[ngroups];
    double* part = dalloc((size_t)(nblk + n) * 4, st, true);
    std::vector<double> hp((size_t)(nblk + n) * 4);
    try {
        for (int g = 0; g < ngroups; ++g) {
            const int a = h->group_off[g], ng = h->group_off[g + 1] - a;
            grad_reduce_kernel<<<blk_off[g + 1] - blk_off[g], 256, 0, st>>>(to_dev(*k), h->dX, h->dim, a, ng, h->dAlpha, h->dB,
                                                                           h->dG, np, part + (size_t)blk_off[g] * 4); ++g_kernel_launches;
        }
        beta_reduce_kernel<<<n, 256, 0, st>>>(n, h->dAlpha, h->dB, h->dG, np, part + (size_t)nblk * 4); ++g_kernel_launches;
        launch_reduce(st, 1, h->dAlpha, nullptr, 0, n, 0, c__->d_scal + 3);
        GD_CUDA(cudaMemcpyAsync(hp.data(), part, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        check_abort(c__);
    } catch (...) {
        dfree(part);
        throw;
    }
    dfree(part);
    const double sum_alpha = read_scalar(c__, c__->d_scal + 3);
    double sk[3] = {0, 0, 0}, sG = 0.0, trB = 0.0;
    for (int b = 0; b < nblk; ++b)
        for (int q = 0; q < 3; ++q) sk[q] += hp[(size_t)b * 4 + q];
    for (int j = 0; j < n; ++j) {
        sG += hp[(size_t)(nblk + j) * 4 + 0];
        trB += hp[(size_t)(nblk + j) * 4 + 1];
    }
    int i = 0;
    if (noise) dmll[i++] = std::exp(2.0 * log_noise) * trB;               // :162
    if (domean && k->mean_const != 0.0) dmll[i++] = sum_alpha;            // dot(ones, alpha), :165-171
    if (kern) {
        dmll[i++] = 0.5 * sk[0];
        dmll[i++] = 0.5 * sk[1];
        if (k->const_sigma2 != 0.0) dmll[i++] = 0.5 * sk[2];
    }
    if (beta) dmll[i++] = 0.5 * (-2.0 / lin_ell2) * sG;                   // LinIso: dK/dll = -2 D'D / l2
    if (ndmll) *ndmll = i;
    API_END
}

int geordd_multigp_postmean_beta(geordd_multigp* h, const geordd_kernel* k, double lin_ell2, double log_noise,
                                 double* beta_hat) {
    if (!h) return GEORDD_ERR_ARG;
    API_BEGIN(h->ctx)
    mg_check_kernel(k, lin_ell2);
    if (!beta_hat) throw ApiError(GEORDD_ERR_ARG, "postmean_beta: output is NULL");
    const int n = h->n, np = h->npad, p = h->p, pp = round_up(p, TILE);
    cudaStream_t st = c__->stream;
    // Sigma_Y|beta = blockdiag(K_g) + noise I  -> PDMat (plain Cholesky, no jitter), get_SigmaYbeta :299-305
    GD_CUDA(cudaMemsetAsync(h->dK, 0, (size_t)np * np * sizeof(double), st));
    add_spatial_blocks(h, k, h->dK);
    launch_add_diag(st, h->dK, np, n, std::max(std::exp(2.0 * log_noise), 1e-8));
    if (np > n) launch_add_diag(st, h->dK + (size_t)n * np + n, np, np - n, 1.0);
    {
        ProfScope ps(c__, PC_POTRF);
        launch_potrf(st, c__->sched, h->dK, h->dL, np, np);
    }
    GD_CUDA(cudaMemcpyAsync(c__->h_info, c__->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, st));
    GD_CUDA(cudaStreamSynchronize(st));
    int info = c__->h_info[0];
    if (info != 0) {
        GD_CUDA(cudaMemsetAsync(c__->sched.abort_word, 0, sizeof(int), st));
        if (info < 0) throw ApiError(GEORDD_ERR_WATCHDOG, "potrf watchdog fired");
        c__->err = "Sigma_Y|beta is not positive definite";
        return info;
    }
    launch_mirror_factor(st, h->dL, np);
    double *W = nullptr, *prec = nullptr, *rhs = nullptr, *v = nullptr;
    try {
        // W = L^-1 D'  (n x p): X_invA_Xt(Sigma, D) = W'W   (:337)
        std::vector<double> Dt((size_t)np * pp, 0.0);
        std::vector<double> Dh((size_t)h->ppad * np);
        GD_CUDA(cudaMemcpyAsync(Dh.data(), h->dD, Dh.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        GD_CUDA(cudaStreamSynchronize(st));
        for (int j = 0; j < n; ++j)
            for (int q = 0; q < p; ++q) Dt[(size_t)q * np + j] = Dh[(size_t)j * h->ppad + q];
        W = dalloc((size_t)np * pp, st, false);
        GD_CUDA(cudaMemcpyAsync(W, Dt.data(), Dt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        {
            ProfScope ps(c__, PC_FWD);
            SolveEpilogue none;
            launch_solve(st, c__->sched, SOLVE_FWD, h->dL, np, W, np, pp, none, p);
        }
        prec = dalloc((size_t)pp * pp, st, true);
        gemm(c__, false, false, pp, pp, np, 1.0, W, np, W, np, 0.0, prec, pp, 0);
        launch_add_diag(st, prec, pp, p, lin_ell2);                                   // + l2 I  (:338-340)
        // rhs = Sigma \ (y - m)
        rhs = dalloc((size_t)np * 64, st, true);
        launch_axpby(st, n, 1.0, h->dY, 0.0, rhs);
        launch_add_const(st, rhs, np, n, 1, -k->mean_const);
        SolveEpilogue none;
        pd_solve(c__, h->dL, np, rhs, np, 64, none, 1);
        // beta = precision \ (D * rhs)    [== (precision \ D) * rhs, :342]
        v = dalloc((size_t)pp * 64, st, true);
        // D (p x n) stored ppad x np column-major: A(m,k) = D(m,k) at k*ppad + m -> a_mc with M = pp needs ld >= pp: use Dt'
        // use Dt (np x pp, ld np): A(m,k) = Dt(k,m) at m*np + k -> a_mc = false
        GD_CUDA(cudaMemcpyAsync(W, Dt.data(), Dt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        gemm(c__, false, false, pp, 64, np, 1.0, W, np, rhs, np, 0.0, v, pp, 0);
        {
            ProfScope ps(c__, PC_LU);
            launch_lu_solve(st, prec, pp, p, v, pp, 1, c__->d_info);
        }
        GD_CUDA(cudaMemcpyAsync(beta_hat, v, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, st));
        check_abort(c__);
    } catch (...) {
        dfree(W); dfree(prec); dfree(rhs); dfree(v);
        throw;
    }
    dfree(W); dfree(prec); dfree(rhs); dfree(v);
    API_END
}

int geordd_multigp_destroy(geordd_multigp* h) {
    if (!h) return 0;
    try { geordd::ctx_use(h->ctx); } catch (...) { return GEORDD_ERR_CUDA; }
    cudaStreamSynchronize(h->ctx->stream);
    geordd::mg_free(h);
    return 0;
}

}  // extern "C"
