// C ABI, part 1: context, profiling, covariance assembly, single-region GP fit.
#include "api_internal.h"
#include <algorithm>
#include <mutex>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler is attached

namespace geordd {

std::atomic<long long> g_kernel_launches{0};

const char* const kProfNames[PC_COUNT] = {"cov", "potrf", "solve_fwd", "solve_bwd", "trmm",
                                          "gemm", "rng", "finish", "lu", "misc", "solve_seq"};

// ---- profiling -------------------------------------------------------------------------------------
// GEORDD_DEBUG_SYNC=1: synchronise before and after every profiled phase and name the phase in which a CUDA error first
// shows up (errors otherwise surface at the next synchronising ABI call, far from the launch that caused them).
static const bool kDebugSync = getenv("GEORDD_DEBUG_SYNC") != nullptr;
ProfScope::ProfScope(geordd_ctx* ctx, int cls_) : c(ctx), cls(cls_) {
    if (kDebugSync) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) fprintf(stderr, "geordd_b200 DEBUG_SYNC: %s BEFORE phase '%s' (an un-scoped launch)\n", cudaGetErrorString(e), kProfNames[cls]);
    }
    nvtxRangePushA(kProfNames[cls]);   // one NVTX range per kernel class / phase (nsys / ncu --nvtx)
    if (!c->prof_on) return;
    cudaEvent_t a, b;
    if (c->ev_pool.size() >= 2) {
        a = c->ev_pool.back(); c->ev_pool.pop_back();
        b = c->ev_pool.back(); c->ev_pool.pop_back();
    } else {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    cudaEventRecord(a, c->stream);
    c->recs.push_back({cls, a, b});
    idx = (int)c->recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->recs[idx].b, c->stream);
    if (kDebugSync) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) fprintf(stderr, "geordd_b200 DEBUG_SYNC: %s after phase '%s'\n", cudaGetErrorString(e), kProfNames[cls]);
    }
    nvtxRangePop();
}

static void prof_collect(geordd_ctx* c) {
    cudaStreamSynchronize(c->stream);
    for (auto& r : c->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            c->prof_ms[r.cls] += ms;
            c->prof_n[r.cls] += 1;
        }
        c->ev_pool.push_back(r.a);
        c->ev_pool.push_back(r.b);
    }
    c->recs.clear();
}

// ---- small helpers ---------------------------------------------------------------------------------
static thread_local geordd_ctx* tl_ctx = nullptr;

void ctx_use(geordd_ctx* c) {
    GD_CUDA(cudaSetDevice(c->device));
    tl_ctx = c;
}

int ctx_fail(geordd_ctx* c, const ApiError& e) {
    if (c) c->err = e.what();
    cudaGetLastError();
    return e.code;
}

void check_abort(geordd_ctx* c) {
    GD_CUDA(cudaMemcpyAsync(c->h_info, c->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GD_CUDA(cudaStreamSynchronize(c->stream));
    GD_CUDA(cudaGetLastError());
    if (c->h_info[0] < 0) {
        GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
        throw ApiError(GEORDD_ERR_WATCHDOG, "dataflow kernel watchdog fired (scheduling stall)");
    }
}

double read_scalar(geordd_ctx* c, const double* d) {
    GD_CUDA(cudaMemcpyAsync(c->h_scal, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GD_CUDA(cudaStreamSynchronize(c->stream));
    return c->h_scal[0];
}

void upload2d(geordd_ctx* c, double* dst, long ldd, const double* src, long lds, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return;
    GD_CUDA(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), (size_t)rows * sizeof(double), cols,
                              cudaMemcpyHostToDevice, c->stream));
}
void download2d(geordd_ctx* c, double* dst, long ldd, const double* src, long lds, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return;
    GD_CUDA(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), (size_t)rows * sizeof(double), cols,
                              cudaMemcpyDeviceToHost, c->stream));
}

static const size_t kPoolCapBytes = (size_t)64 << 30;

// ---- guard mode (GEORDD_GUARD=1): our own out-of-bounds detector ----------------------------------------------------
// compute-sanitizer is closed on the B200 pool this library is developed on, so the allocator can bracket every device
// buffer with 64 KiB guard zones filled with 0xFF (= a NaN pattern as doubles): an out-of-bounds WRITE is found when the
// buffer is released (the guards are read back and compared), an out-of-bounds READ of up to 64 KiB turns into NaN in the
// results (scripts/sanitize_driver.py asserts every statistic finite).  Buffers are not pooled and not rounded up in this
// mode, so the first byte past the requested size is already guard.  geordd_dbg_guard_violations() counts the hits.
static const bool kGuard = getenv("GEORDD_GUARD") != nullptr;
static const size_t kGuardBytes = (size_t)64 << 10;
static std::atomic<long long> g_guard_violations{0};
static std::unordered_map<void*, size_t> g_guard_sizes;   // user pointer -> user bytes (guard mode only; one thread per ctx, global map locked)
static std::mutex g_guard_mutex;

static void guard_check_and_free(void* user, cudaStream_t st) {
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_guard_mutex);
        auto it = g_guard_sizes.find(user);
        if (it == g_guard_sizes.end()) return;
        bytes = it->second;
        g_guard_sizes.erase(it);
    }
    char* raw = (char*)user - kGuardBytes;
    std::vector<unsigned char> h(2 * kGuardBytes);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), raw, kGuardBytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data() + kGuardBytes, raw + kGuardBytes + bytes, kGuardBytes, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (unsigned char v : h) bad += v != 0xFF;
    if (bad) {
        g_guard_violations += 1;
        fprintf(stderr, "geordd_b200 GUARD: %lld guard bytes of a %zu-byte buffer were overwritten\n", bad, bytes);
    }
    cudaFree(raw);
}

static void pool_trim(geordd_ctx* c, size_t keep_bytes) {
    while (c->pool_cached_bytes > keep_bytes && !c->pool_free.empty()) {
        auto it = std::prev(c->pool_free.end());   // largest first
        cudaFree(it->second);
        c->pool_cached_bytes -= it->first;
        c->pool_free.erase(it);
    }
}

double* dalloc(size_t count, cudaStream_t st, bool zero) {
    geordd_ctx* c = tl_ctx;
    if (count == 0) count = 1;
    if (kGuard) {
        const size_t ub = (count * sizeof(double) + 15) / 16 * 16;
        char* raw = nullptr;
        if (cudaMalloc(&raw, ub + 2 * kGuardBytes) != cudaSuccess) {
            cudaGetLastError();
            throw ApiError(GEORDD_ERR_NOMEM, "cudaMalloc (guard mode) failed");
        }
        GD_CUDA(cudaMemsetAsync(raw, 0xFF, ub + 2 * kGuardBytes, st));   // the buffer itself starts as NaN too: reads of
        if (zero) GD_CUDA(cudaMemsetAsync(raw + kGuardBytes, 0, count * sizeof(double), st));   // never-written memory show up
        std::lock_guard<std::mutex> lk(g_guard_mutex);
        g_guard_sizes[raw + kGuardBytes] = ub;
        return (double*)(raw + kGuardBytes);
    }
    size_t bytes = (count * sizeof(double) + 511) / 512 * 512;
    void* p = nullptr;
    if (c) {
        auto it = c->pool_free.lower_bound(bytes);
        if (it != c->pool_free.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
            p = it->second;
            bytes = it->first;
            c->pool_cached_bytes -= it->first;
            c->pool_free.erase(it);
        }
    }
    if (!p) {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess && c) {   // give cached blocks back to the driver and retry once
            cudaGetLastError();
            cudaStreamSynchronize(c->stream);
            pool_trim(c, 0);
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw ApiError(GEORDD_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
    }
    if (c) c->pool_live[p] = bytes;
    if (zero) GD_CUDA(cudaMemsetAsync(p, 0, count * sizeof(double), st));
    return (double*)p;
}

// All work of a context is on one stream, so a block handed back here can be reused by later launches on that
// stream without waiting for the earlier ones.
void dfree(void* p) {
    if (!p) return;
    geordd_ctx* c = tl_ctx;
    if (kGuard) {
        guard_check_and_free(p, c ? c->stream : nullptr);
        return;
    }
    if (c) {
        auto it = c->pool_live.find(p);
        if (it != c->pool_live.end()) {
            c->pool_free.emplace(it->second, p);
            c->pool_cached_bytes += it->second;
            c->pool_live.erase(it);
            if (c->pool_cached_bytes > kPoolCapBytes) {
                cudaStreamSynchronize(c->stream);
                pool_trim(c, kPoolCapBytes / 2);
            }
            return;
        }
    }
    cudaFree(p);
}

long long g_guard_violations_get() { return g_guard_violations.load(); }

void ensure_ws(geordd_ctx* c, size_t count) { c->ws.ensure(count, c->stream, false); }

KernelSpecDev to_dev(const geordd_kernel& k) {
    KernelSpecDev d;
    d.kind = k.kind; d.ell = k.ell; d.sigma2 = k.sigma2; d.const_sigma2 = k.const_sigma2;
    return d;
}

static void check_kernel(const geordd_kernel* k) {
    if (!k) throw ApiError(GEORDD_ERR_ARG, "kernel is NULL");
    if (k->kind < 0 || k->kind > 3) throw ApiError(GEORDD_ERR_ARG, "kernel kind must be 0..3 (SE, Mat12, Mat32, Mat52)");
    if (!(k->ell > 0.0) || !(k->sigma2 >= 0.0) || !(k->const_sigma2 >= 0.0))
        throw ApiError(GEORDD_ERR_ARG, "kernel parameters must be positive and finite");
}

int pick_splitk(geordd_ctx* c, int M, int N, int K) {
    long tiles = (long)(M / TILE) * (N / 64);
    long kch = K / BK;
    if (tiles >= 2L * c->sched.num_sms || kch < 16) return 1;
    long want = (2L * c->sched.num_sms + tiles - 1) / tiles;
    long maxs = kch / 8;   // at least 8 chunks (K = 128) per split
    if (maxs < 1) maxs = 1;
    return (int)std::min(want, maxs);
}

void gemm(geordd_ctx* c, bool a_mc, bool b_nc, int M, int N, int K, double alpha, const double* A, long lda,
          const double* B, long ldb, double beta, double* C, long ldc, int splitk, bool b_xb) {
    if (splitk <= 0) splitk = pick_splitk(c, M, N, K);
    if (splitk > 1) ensure_ws(c, (size_t)splitk * ldc * N);
    ProfScope ps(c, PC_GEMM);
    launch_gemm(c->stream, a_mc, b_nc, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, c->ws.p, b_xb);
}

int make_posdef(geordd_ctx* c, double* K, double* L, long ld, int n, int npad, int* rounds, const double* Lpre, int npad_pre,
                int ptiles) {
    int nr = 0;
    for (int attempt = 0; attempt <= 10; ++attempt) {
        {
            ProfScope ps(c, PC_POTRF);
            // the known leading block is only valid for the un-jittered matrix (first attempt)
            if (attempt == 0) launch_potrf(c->stream, c->sched, K, L, ld, npad, Lpre, npad_pre, ptiles);
            else launch_potrf(c->stream, c->sched, K, L, ld, npad);
        }
        GD_CUDA(cudaMemcpyAsync(c->h_info, c->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
        int info = c->h_info[0];
        if (info != 0) GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
        if (info < 0) throw ApiError(GEORDD_ERR_WATCHDOG, "potrf watchdog fired");
        if (info == 0) {
            if (rounds) *rounds = nr;
            ProfScope ps(c, PC_MISC);
            launch_mirror_factor(c->stream, L, npad);
            return 0;
        }
        if (attempt == 10) {
            if (rounds) *rounds = nr;
            return info;
        }
        // eps = 1e-6 * tr(M) / n on the already-jittered matrix; M[i,i] += eps
        launch_reduce(c->stream, 0, K, nullptr, ld, n, 0, c->d_scal);
        double tr = read_scalar(c, c->d_scal);
        double eps = 1e-6 * tr / (double)n;
        launch_add_diag(c->stream, K, ld, n, eps);
        ++nr;
    }
    return -1;
}

void pd_solve(geordd_ctx* c, const double* L, int npad, double* B, long ldb, int ncols, const SolveEpilogue& epi_bwd,
              int live_cols) {
    SolveEpilogue none;
    none.xb = epi_bwd.xb;
    {
        ProfScope ps(c, PC_FWD);
        launch_solve(c->stream, c->sched, SOLVE_FWD, L, npad, B, ldb, ncols, none, live_cols);
    }
    {
        ProfScope ps(c, PC_BWD);
        launch_solve(c->stream, c->sched, SOLVE_BWD, L, npad, B, ldb, ncols, epi_bwd, live_cols);
    }
}

// ---- GP ----------------------------------------------------------------------------------------------
void gp_free(geordd_gp* gp) {
    if (!gp) return;
    dfree(gp->dX); dfree(gp->dK); dfree(gp->dL); dfree(gp->dDinv); dfree(gp->dR); dfree(gp->dAlpha);
    delete gp;
}

void gp_update_alpha(geordd_gp* gp) {
    geordd_ctx* c = gp->ctx;
    const int npad = gp->npad;
    // slab: column 0 = r, others zero
    GD_CUDA(cudaMemsetAsync(gp->dAlpha, 0, (size_t)npad * 64 * sizeof(double), c->stream));
    GD_CUDA(cudaMemcpyAsync(gp->dAlpha, gp->dR, (size_t)npad * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    SolveEpilogue none;
    pd_solve(c, gp->dL, npad, gp->dAlpha, npad, 64, none, 1);
    launch_reduce(c->stream, 2, gp->dR, gp->dAlpha, 0, gp->n, 0, c->d_scal);
    check_abort(c);
    double quad = read_scalar(c, c->d_scal);
    gp->mll = -quad / 2.0 - gp->logdet / 2.0 - (double)gp->n * std::log(2.0 * M_PI) / 2.0;
}

static void gp_set_resid(geordd_gp* gp, const double* y) {
    geordd_ctx* c = gp->ctx;
    gp->y.assign(y, y + gp->n);
    std::vector<double> r(gp->npad, 0.0);
    for (int i = 0; i < gp->n; ++i) r[i] = y[i] - gp->k.mean_const;
    GD_CUDA(cudaMemcpyAsync(gp->dR, r.data(), (size_t)gp->npad * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GD_CUDA(cudaStreamSynchronize(c->stream));
}

// Phase 1 of a fit: allocate, upload the coordinates, assemble K + (exp(2 logNoise) + eps) I  (K1).
static geordd_gp* gp_alloc_assemble(geordd_ctx* c, const geordd_kernel* k, const double* X, int dim, int n, double log_noise) {
    check_kernel(k);
    if (!X || n <= 0 || dim <= 0 || dim > 8) throw ApiError(GEORDD_ERR_ARG, "gp_fit: bad X / y / n / dim (1..8)");
    geordd_gp* gp = new geordd_gp();
    try {
        gp->ctx = c; gp->k = *k; gp->log_noise = log_noise; gp->n = n; gp->dim = dim;
        gp->npad = round_up(n, TILE);
        const size_t np = gp->npad;
        gp->X.assign(X, X + (size_t)dim * n);
        gp->dX = dalloc((size_t)dim * n, c->stream, false);
        gp->dK = dalloc(np * np, c->stream, false);
        gp->dL = dalloc(factor_elems(gp->npad), c->stream, false);
        gp->dR = dalloc(np, c->stream, true);
        gp->dAlpha = dalloc(np * 64, c->stream, true);
        GD_CUDA(cudaMemcpyAsync(gp->dX, X, (size_t)dim * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const double noise = std::exp(2.0 * log_noise) + 2.220446049250313e-16;   // exp(2 logNoise) + eps()
        ProfScope ps(c, PC_COV);
        launch_cov(c->stream, to_dev(*k), gp->dX, n, gp->dX, n, dim, true, false, noise, gp->dK, np, gp->npad, gp->npad, false);
    } catch (...) {
        gp_free(gp);
        throw;
    }
    return gp;
}

// Phase 3: log-determinant, alpha = cK \ (y - m), mll.
static void gp_finish(geordd_gp* gp, const double* y) {
    geordd_ctx* c = gp->ctx;
    launch_logdiag_sum(c->stream, gp->dL, gp->npad, gp->n, c->d_scal + 1);
    gp->logdet = 2.0 * read_scalar(c, c->d_scal + 1);
    gp_set_resid(gp, y);
    gp_update_alpha(gp);
}

static void gp_finish_batch(geordd_ctx* c, int nfit, geordd_gp** gps, const double* const* ys);

geordd_gp* gp_create(geordd_ctx* c, const geordd_kernel* k, const double* X, int dim, int n, const double* y,
                     double log_noise, int* info_out, const geordd_gp* lead) {
    if (!y) throw ApiError(GEORDD_ERR_ARG, "gp_fit: bad X / y / n / dim (1..8)");
    geordd_gp* gp = gp_alloc_assemble(c, k, X, dim, n, log_noise);
    try {
        // leading block already factored?  Same units, kernel and noise, and that fit needed no jitter.
        const double* Lpre = nullptr;
        int ptiles = 0;
        if (lead && lead->ctx == c && lead->dim == dim && lead->n <= n && lead->jitter_rounds == 0 && lead->log_noise == log_noise &&
            lead->k.kind == k->kind && lead->k.ell == k->ell && lead->k.sigma2 == k->sigma2 && lead->k.const_sigma2 == k->const_sigma2 &&
            std::memcmp(lead->X.data(), X, (size_t)dim * lead->n * sizeof(double)) == 0) {
            Lpre = lead->dL;
            ptiles = lead->n / TILE;   // tiles entirely inside the leading block
        }
        // Fast path (no jitter needed, the rule): factorisation and tails queued back to back, one host synchronisation (see
        // gp_create_batch).  Same kernels and arithmetic as make_posdef + gp_finish, which remain the path after a
        // non-positive pivot.
        {
            ProfScope ps(c, PC_POTRF);
            launch_potrf(c->stream, c->sched, gp->dK, gp->dL, gp->npad, gp->npad, Lpre, lead ? lead->npad : 0, ptiles);
        }
        gp_finish_batch(c, 1, &gp, &y);
        int info = 0;
        if (c->h_info[0] > 0) {
            GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
            info = make_posdef(c, gp->dK, gp->dL, gp->npad, n, gp->npad, &gp->jitter_rounds, Lpre, lead ? lead->npad : 0, ptiles);
            if (info == 0) gp_finish(gp, y);
        }
        if (info_out) *info_out = info;
        if (info != 0) {
            gp_free(gp);
            return nullptr;
        }
    } catch (...) {
        gp_free(gp);
        throw;
    }
    return gp;
}

// Tails of the fits of a batch -- mirror the factor, log-determinant, alpha = cK \ (y - m), y'alpha -- are independent chains
// of small / latency-bound launches (the few-column solve is a chain of T dependent row blocks on T SMs).  They run side by
// side on up to four side streams (own job counter and flag region per lane), forked from and joined back into the
// context's stream with events; all scalars come back with ONE copy.  Same kernels, same arithmetic as gp_finish.
void ensure_side_streams(geordd_ctx* c) {
    for (int l = 0; l < NARROW_LANES; ++l) {
        if (!c->side[l]) GD_CUDA(cudaStreamCreateWithFlags(&c->side[l], cudaStreamNonBlocking));
        if (!c->ev_join[l]) GD_CUDA(cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming));
    }
    if (!c->ev_fork) GD_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
}

static void gp_finish_batch(geordd_ctx* c, int nfit, geordd_gp** gps, const double* const* ys) {
    ensure_side_streams(c);
    constexpr int kRound = 24;   // two scalars per fit in d_scal[8 .. 8 + 2 * kRound)
    std::vector<std::vector<double>> resid(nfit);
    for (int f0 = 0; f0 < nfit; f0 += kRound) {
        const int nf = std::min(kRound, nfit - f0);
        for (int f = f0; f < f0 + nf; ++f) {   // residuals y - m (host -> device) on the main stream
            geordd_gp* gp = gps[f];
            gp->y.assign(ys[f], ys[f] + gp->n);
            resid[f].assign(gp->npad, 0.0);
            for (int i = 0; i < gp->n; ++i) resid[f][i] = ys[f][i] - gp->k.mean_const;
            GD_CUDA(cudaMemcpyAsync(gp->dR, resid[f].data(), (size_t)gp->npad * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        }
        GD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        for (int f = f0; f < f0 + nf; ++f) {
            geordd_gp* gp = gps[f];
            const int lane = (f - f0) % NARROW_LANES;
            cudaStream_t s = c->side[lane];
            const int npad = gp->npad;
            double* slot = c->d_scal + 8 + 2 * (f - f0);
            if (f - f0 < NARROW_LANES) GD_CUDA(cudaStreamWaitEvent(s, c->ev_fork, 0));
            gp->jitter_rounds = 0;
            launch_mirror_factor(s, gp->dL, npad);
            launch_logdiag_sum(s, gp->dL, npad, gp->n, slot);
            GD_CUDA(cudaMemsetAsync(gp->dAlpha, 0, (size_t)npad * 64 * sizeof(double), s));
            GD_CUDA(cudaMemcpyAsync(gp->dAlpha, gp->dR, (size_t)npad * sizeof(double), cudaMemcpyDeviceToDevice, s));
            if (npad > TILE) {
                launch_solve_narrow(s, c->sched, SOLVE_FWD, gp->dL, npad, gp->dAlpha, npad, 1, lane);
                launch_solve_narrow(s, c->sched, SOLVE_BWD, gp->dL, npad, gp->dAlpha, npad, 1, lane);
            } else {   // single tile: the wide kernel (gp_update_alpha's path for npad == TILE); serialised on the main stream below
                continue;
            }
            launch_reduce(s, 2, gp->dR, gp->dAlpha, 0, gp->n, 0, slot + 1);
        }
        for (int l = 0; l < std::min(nf, NARROW_LANES); ++l) {
            GD_CUDA(cudaEventRecord(c->ev_join[l], c->side[l]));
            GD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[l], 0));
        }
        for (int f = f0; f < f0 + nf; ++f) {   // one-tile GPs: same launches as gp_update_alpha, on the main stream
            geordd_gp* gp = gps[f];
            if (gp->npad > TILE) continue;
            SolveEpilogue none;
            pd_solve(c, gp->dL, gp->npad, gp->dAlpha, gp->npad, 64, none, 1);
            launch_reduce(c->stream, 2, gp->dR, gp->dAlpha, 0, gp->n, 0, c->d_scal + 8 + 2 * (f - f0) + 1);
        }
        GD_CUDA(cudaMemcpyAsync(c->h_scal + 8, c->d_scal + 8, (size_t)2 * nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        check_abort(c);   // synchronises the main stream (which has joined the side streams)
        for (int f = f0; f < f0 + nf; ++f) {
            geordd_gp* gp = gps[f];
            gp->logdet = 2.0 * c->h_scal[8 + 2 * (f - f0)];
            const double quad = c->h_scal[8 + 2 * (f - f0) + 1];
            gp->mll = -quad / 2.0 - gp->logdet / 2.0 - (double)gp->n * std::log(2.0 * M_PI) / 2.0;
        }
    }
}

// Several independent fits (the regions of GPRealisations, the halves of a placebo split): all covariance matrices are
// assembled first, then factored by ONE dataflow launch whose merged job list interleaves their critical paths
// (potrf.cu).  A batch in which any matrix needs make_posdef! jitter is redone matrix by matrix (rare).
// On return gps[] holds the fitted GPs; a failed fit leaves nullptr there and its LAPACK info in infos[].
void gp_create_batch(geordd_ctx* c, int nfit, const geordd_kernel* ks, const double* const* Xs, int dim, const int* ns,
                     const double* const* ys, const double* log_noise, geordd_gp** gps, int* infos) {
    for (int f = 0; f < nfit; ++f) { gps[f] = nullptr; infos[f] = 0; }
    try {
        for (int f = 0; f < nfit; ++f) {
            if (!ys[f]) throw ApiError(GEORDD_ERR_ARG, "gp_fit_batch: y is NULL");
            gps[f] = gp_alloc_assemble(c, &ks[f], Xs[f], dim, ns[f], log_noise[f]);
        }
        bool fallback = false;
        for (int f0 = 0; f0 < nfit && !fallback; f0 += POTRF_MAX_BATCH) {
            const int nm = std::min(POTRF_MAX_BATCH, nfit - f0);
            PotrfItem items[POTRF_MAX_BATCH];
            for (int m = 0; m < nm; ++m) items[m] = PotrfItem{gps[f0 + m]->dK, gps[f0 + m]->dL, gps[f0 + m]->npad, gps[f0 + m]->npad, nullptr, 0, 0};
            int rc;
            {
                ProfScope ps(c, PC_POTRF);
                rc = launch_potrf_batch(c->stream, c->sched, items, nm);
            }
            if (rc != 0) { fallback = true; break; }
            // One launch for the whole batch (the usual case): the tails are queued behind it at once and the factorisation's
            // status comes back with their scalars -- ONE host synchronisation per call (each one costs the caller a wake-up,
            // milliseconds on a busy host).  A non-positive pivot raises the abort word: the tails then bail out at their first
            // wait, and the batch is redone matrix by matrix below.
            if (nfit <= POTRF_MAX_BATCH) break;
            GD_CUDA(cudaMemcpyAsync(c->h_info, c->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            GD_CUDA(cudaStreamSynchronize(c->stream));
            GD_CUDA(cudaGetLastError());
            const int info = c->h_info[0];
            if (info != 0) GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
            if (info < 0) throw ApiError(GEORDD_ERR_WATCHDOG, "potrf watchdog fired");
            if (info > 0) fallback = true;
        }
        if (!fallback && nfit <= POTRF_MAX_BATCH) {
            gp_finish_batch(c, nfit, gps, ys);   // its check_abort leaves the abort word in h_info (and throws on a watchdog)
            if (c->h_info[0] > 0) {
                GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
                fallback = true;
            }
        }
        if (fallback) {
            for (int f = 0; f < nfit; ++f) {
                geordd_gp* gp = gps[f];
                infos[f] = make_posdef(c, gp->dK, gp->dL, gp->npad, gp->n, gp->npad, &gp->jitter_rounds);
                if (infos[f] != 0) { gp_free(gp); gps[f] = nullptr; continue; }
                gp_finish(gp, ys[f]);
            }
        } else if (nfit > POTRF_MAX_BATCH) {
            gp_finish_batch(c, nfit, gps, ys);
        }
    } catch (...) {
        for (int f = 0; f < nfit; ++f) { if (gps[f]) gp_free(gps[f]); gps[f] = nullptr; }
        throw;
    }
}

// GEORDD_SEQ_SUBST=1 keeps round 1's shared-memory substitution in the fused solve sequence (A/B measurements).
static const bool kSeqSubst = getenv("GEORDD_SEQ_SUBST") != nullptr;
const double* gp_diag_inverse(geordd_gp* gp) {
    if (kSeqSubst) return nullptr;
    if (!gp->dDinv) {
        geordd_ctx* c = gp->ctx;
        gp->dDinv = dalloc(diag_inverse_elems(gp->npad), c->stream, false);
        ProfScope ps(c, PC_MISC);
        launch_diag_inverse(c->stream, gp->dL, gp->npad, gp->dDinv);
    }
    return gp->dDinv;
}

void gp_ensure_full_K(geordd_gp* gp) {
    if (gp->k_full) return;
    geordd_ctx* c = gp->ctx;
    // mirror the lower triangle into the upper (tile transposes), keeps the jittered diagonal
    launch_mirror_lower(c->stream, gp->dK, gp->npad, gp->npad);
    gp->k_full = true;
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

int geordd_version(void) { return 100; }

long long geordd_launch_count(void) { return geordd::g_kernel_launches.load(); }

int geordd_ctx_create(geordd_ctx** out, int device) {
    if (!out) return GEORDD_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return GEORDD_ERR_CUDA;   // no GPU: there is deliberately no CPU fallback
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return GEORDD_ERR_CUDA;
    geordd_ctx* c = new geordd_ctx();
    try {
        c->device = device;
        GD_CUDA(cudaSetDevice(device));
        GD_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        c->sched.num_sms = prop.multiProcessorCount;
        c->sched.flags_cap = 4u << 20;
        c->sched.jobs_cap = 1u << 20;
        GD_CUDA(cudaMalloc(&c->sched.counter, 16 * sizeof(unsigned)));
        GD_CUDA(cudaMalloc(&c->sched.flags, c->sched.flags_cap * sizeof(unsigned)));
        GD_CUDA(cudaMalloc(&c->sched.abort_word, sizeof(int)));
        GD_CUDA(cudaMalloc(&c->sched.jobs, c->sched.jobs_cap * sizeof(int2)));
        GD_CUDA(cudaMemset(c->sched.counter, 0, 16 * sizeof(unsigned)));
        GD_CUDA(cudaMemset(c->sched.flags, 0, c->sched.flags_cap * sizeof(unsigned)));
        GD_CUDA(cudaMemset(c->sched.abort_word, 0, sizeof(int)));
        c->sched.epoch = 0;
        GD_CUDA(cudaMalloc(&c->d_scal, 64 * sizeof(double)));
        GD_CUDA(cudaMallocHost(&c->h_scal, 64 * sizeof(double)));
        GD_CUDA(cudaMalloc(&c->d_info, 4 * sizeof(int)));
        GD_CUDA(cudaMallocHost(&c->h_info, 4 * sizeof(int)));
        GD_CUDA(cudaMalloc(&c->d_tally, 8 * sizeof(unsigned long long)));
        GD_CUDA(cudaMallocHost(&c->h_tally, 8 * sizeof(unsigned long long)));
    } catch (const ApiError&) {
        delete c;
        return GEORDD_ERR_CUDA;
    }
    *out = c;
    return 0;
}

int geordd_ctx_destroy(geordd_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    geordd_comm_destroy(c);
    for (int l = 0; l < 4; ++l) {
        if (c->side[l]) cudaStreamDestroy(c->side[l]);
        if (c->ev_join[l]) cudaEventDestroy(c->ev_join[l]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaFree(c->sched.counter); cudaFree(c->sched.flags); cudaFree(c->sched.abort_word); cudaFree(c->sched.jobs);
    cudaFree(c->d_scal); cudaFreeHost(c->h_scal); cudaFree(c->d_info); cudaFreeHost(c->h_info);
    cudaFree(c->d_tally); cudaFreeHost(c->h_tally);
    tl_ctx = c;
    c->ws.release(); c->tmp0.release(); c->tmp1.release(); c->tmp2.release(); c->tmp3.release();
    c->ws2.release(); c->tmp4.release(); c->tmp5.release();
    pool_trim(c, 0);
    for (auto& kv : c->pool_live) cudaFree(kv.first);
    c->pool_live.clear();
    tl_ctx = nullptr;
    cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

const char* geordd_last_error(const geordd_ctx* c) { return c ? c->err.c_str() : "no context (no usable sm_100 GPU)"; }

int geordd_ctx_set_stream(geordd_ctx* c, void* s) {
    if (!c) return GEORDD_ERR_ARG;
    cudaStreamSynchronize(c->stream);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

int geordd_ctx_sync(geordd_ctx* c) {
    API_BEGIN(c)
    GD_CUDA(cudaStreamSynchronize(c->stream));
    API_END
}

int geordd_ctx_profile(geordd_ctx* c, int enable) {
    API_BEGIN(c)
    prof_collect(c);
    for (int i = 0; i < PC_COUNT; ++i) { c->prof_ms[i] = 0.0; c->prof_n[i] = 0; }
    c->prof_on = enable != 0;
    API_END
}

int geordd_ctx_profile_get(geordd_ctx* c, const char* name, double* ms, long long* launches) {
    API_BEGIN(c)
    prof_collect(c);
    double t = 0.0;
    long long n = 0;
    bool found = false;
    for (int i = 0; i < PC_COUNT; ++i) {
        if (!strcmp(name, "*") || !strcmp(name, kProfNames[i])) {
            t += c->prof_ms[i]; n += c->prof_n[i]; found = true;
        }
    }
    if (!found) throw ApiError(GEORDD_ERR_ARG, "unknown profile class");
    if (ms) *ms = t;
    if (launches) *launches = n;
    API_END
}

int geordd_ctx_set_max_batch(geordd_ctx* c, long mb) {
    if (!c || mb < 0) return GEORDD_ERR_ARG;
    c->max_batch = mb;
    return 0;
}

long long geordd_dbg_guard_violations(void) { return geordd::g_guard_violations_get(); }
int geordd_dbg_set_lu_blocked(int on) { geordd::lu_set_blocked(on != 0); return 0; }

int geordd_ctx_last_min_margin(const geordd_ctx* c, double* margin3) {
    if (!c || !margin3) return GEORDD_ERR_ARG;
    for (int k = 0; k < 3; ++k) margin3[k] = c->min_margin[k];
    return 0;
}

int geordd_ctx_set_mll_forward_only(geordd_ctx* c, int enable) {
    if (!c) return GEORDD_ERR_ARG;
    c->mll_forward_only = enable != 0;
    return 0;
}

int geordd_cov_sym(geordd_ctx* c, const geordd_kernel* k, const double* X, int dim, int n, double diag_add, double* K) {
    API_BEGIN(c)
    check_kernel(k);
    if (!X || !K || n <= 0 || dim <= 0 || dim > 8) throw ApiError(GEORDD_ERR_ARG, "cov_sym: bad arguments");
    double* dX = dalloc((size_t)dim * n, c->stream, false);
    double* dK = nullptr;
    try {
        dK = dalloc((size_t)n * n, c->stream, false);
        GD_CUDA(cudaMemcpyAsync(dX, X, (size_t)dim * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        {
            ProfScope ps(c, PC_COV);
            launch_cov(c->stream, to_dev(*k), dX, n, dX, n, dim, true, true, diag_add, dK, n, n, n, false);
        }
        GD_CUDA(cudaMemcpyAsync(K, dK, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(dX); dfree(dK);
        throw;
    }
    dfree(dX); dfree(dK);
    API_END
}

int geordd_cov_cross(geordd_ctx* c, const geordd_kernel* k, const double* X1, int n1, const double* X2, int n2, int dim,
                     double* K) {
    API_BEGIN(c)
    check_kernel(k);
    if (!X1 || !X2 || !K || n1 <= 0 || n2 <= 0 || dim <= 0 || dim > 8) throw ApiError(GEORDD_ERR_ARG, "cov_cross: bad arguments");
    double *d1 = nullptr, *d2 = nullptr, *dK = nullptr;
    try {
        d1 = dalloc((size_t)dim * n1, c->stream, false);
        d2 = dalloc((size_t)dim * n2, c->stream, false);
        dK = dalloc((size_t)n1 * n2, c->stream, false);
        GD_CUDA(cudaMemcpyAsync(d1, X1, (size_t)dim * n1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GD_CUDA(cudaMemcpyAsync(d2, X2, (size_t)dim * n2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        {
            ProfScope ps(c, PC_COV);
            launch_cov(c->stream, to_dev(*k), d1, n1, d2, n2, dim, false, false, 0.0, dK, n1, n1, n2, false);
        }
        GD_CUDA(cudaMemcpyAsync(K, dK, (size_t)n1 * n2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(d1); dfree(d2); dfree(dK);
        throw;
    }
    dfree(d1); dfree(d2); dfree(dK);
    API_END
}

int geordd_gp_fit(geordd_ctx* c, const geordd_kernel* k, const double* X, int dim, int n, const double* y,
                  double log_noise, geordd_gp** out, double* alpha, double* mll, double* cholU, int* jitter_rounds) {
    if (out) *out = nullptr;
    API_BEGIN(c)
    if (!out) throw ApiError(GEORDD_ERR_ARG, "gp_fit: out is NULL");
    int info = 0;
    geordd_gp* gp = gp_create(c, k, X, dim, n, y, log_noise, &info);
    if (!gp) {
        c->err = "matrix is not positive definite; Cholesky factorization failed";
        return info;
    }
    if (alpha) GD_CUDA(cudaMemcpyAsync(alpha, gp->dAlpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (mll) *mll = gp->mll;
    if (jitter_rounds) *jitter_rounds = gp->jitter_rounds;
    if (cholU) {
        // U = L' : transpose-copy the logical n x n block through scratch
        c->tmp0.ensure((size_t)n * n, c->stream, false);
        launch_factor_unpack(c->stream, gp->dL, gp->npad, c->tmp0.p, n, n, true);
        GD_CUDA(cudaMemcpyAsync(cholU, c->tmp0.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    GD_CUDA(cudaStreamSynchronize(c->stream));
    *out = gp;
    API_END
}

int geordd_gp_fit_batch(geordd_ctx* c, int nfit, const geordd_kernel* k, const double* const* X, int dim, const int* n,
                        const double* const* y, const double* log_noise, geordd_gp** out, double* const* alpha, double* mll,
                        int* jitter_rounds, int* info) {
    if (out) for (int f = 0; f < nfit; ++f) out[f] = nullptr;
    API_BEGIN(c)
    if (!out || nfit <= 0 || !k || !X || !n || !y || !log_noise) throw ApiError(GEORDD_ERR_ARG, "gp_fit_batch: bad arguments");
    std::vector<int> infos(nfit, 0);
    gp_create_batch(c, nfit, k, X, dim, n, y, log_noise, out, infos.data());
    int first_bad = 0;
    for (int f = 0; f < nfit; ++f) {
        if (info) info[f] = infos[f];
        if (!out[f]) { if (!first_bad) first_bad = infos[f] ? infos[f] : GEORDD_ERR_CUDA; continue; }
        if (alpha && alpha[f]) GD_CUDA(cudaMemcpyAsync(alpha[f], out[f]->dAlpha, (size_t)n[f] * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (mll) mll[f] = out[f]->mll;
        if (jitter_rounds) jitter_rounds[f] = out[f]->jitter_rounds;
    }
    GD_CUDA(cudaStreamSynchronize(c->stream));
    if (first_bad != 0) {   // like GPRealisations' constructor, which throws on the first region that is not positive definite
        for (int f = 0; f < nfit; ++f) { if (out[f]) gp_free(out[f]); out[f] = nullptr; }
        c->err = "matrix is not positive definite; Cholesky factorization failed";
        return first_bad;
    }
    API_END
}

int geordd_gp_set_y(geordd_gp* gp, const double* y, double* alpha, double* mll) {
    if (!gp) return GEORDD_ERR_ARG;
    API_BEGIN(gp->ctx)
    if (!y) throw ApiError(GEORDD_ERR_ARG, "set_y: y is NULL");
    geordd::gp_set_resid(gp, y);
    gp_update_alpha(gp);
    if (alpha) GD_CUDA(cudaMemcpyAsync(alpha, gp->dAlpha, (size_t)gp->n * sizeof(double), cudaMemcpyDeviceToHost, c__->stream));
    if (mll) *mll = gp->mll;
    GD_CUDA(cudaStreamSynchronize(c__->stream));
    API_END
}

int geordd_gp_get_K(geordd_gp* gp, double* K) {
    if (!gp) return GEORDD_ERR_ARG;
    API_BEGIN(gp->ctx)
    if (!K) throw ApiError(GEORDD_ERR_ARG, "get_K: K is NULL");
    gp_ensure_full_K(gp);
    download2d(c__, K, gp->n, gp->dK, gp->npad, gp->n, gp->n);
    GD_CUDA(cudaStreamSynchronize(c__->stream));
    API_END
}

int geordd_gp_nobs(const geordd_gp* gp) { return gp ? gp->n : GEORDD_ERR_ARG; }

int geordd_gp_destroy(geordd_gp* gp) {
    if (!gp) return 0;
    try { geordd::ctx_use(gp->ctx); } catch (...) { return GEORDD_ERR_CUDA; }
    cudaStreamSynchronize(gp->ctx->stream);
    gp_free(gp);
    return 0;
}

int geordd_dbg_normals(geordd_ctx* c, unsigned long long seed, unsigned long long draw0, int ndraws, int n, double* Z) {
    API_BEGIN(c)
    if (!Z || n <= 0 || ndraws <= 0) throw ApiError(GEORDD_ERR_ARG, "dbg_normals: bad arguments");
    int npad = round_up(n, 2);
    double* dZ = dalloc((size_t)npad * ndraws, c->stream, false);
    try {
        {
            ProfScope ps(c, PC_RNG);
            launch_philox_normals(c->stream, seed, draw0, ndraws, n, dZ, npad, npad, false);
        }
        download2d(c, Z, n, dZ, npad, n, ndraws);
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(dZ);
        throw;
    }
    dfree(dZ);
    API_END
}

int geordd_dbg_gemm(geordd_ctx* c, int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda,
                    const double* B, int ldb, double beta, double* C, int ldc, int splitk) {
    API_BEGIN(c)
    if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) throw ApiError(GEORDD_ERR_ARG, "dbg_gemm: bad arguments");
    const int Mp = round_up(M, TILE), Np = round_up(N, 64), Kp = round_up(K, BK);
    // op(A) is M x K.  transa == 0: stored M x K col-major -> A(m,k) at k*lda+m (a_mc).  transa == 1: stored K x M.
    const int ar = transa ? K : M, ac = transa ? M : K, arp = transa ? Kp : Mp, acp = transa ? Mp : Kp;
    const int br = transb ? N : K, bc = transb ? K : N, brp = transb ? Np : Kp, bcp = transb ? Kp : Np;
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    try {
        dA = dalloc((size_t)arp * acp, c->stream, true);
        dB = dalloc((size_t)brp * bcp, c->stream, true);
        dC = dalloc((size_t)Mp * Np, c->stream, true);
        upload2d(c, dA, arp, A, lda, ar, ac);
        upload2d(c, dB, brp, B, ldb, br, bc);
        upload2d(c, dC, Mp, C, ldc, M, N);
        // B(k,n): transb == 0 stored K x N col-major -> at n*ldb + k (k contiguous, b_nc = false)
        gemm(c, !transa, transb != 0, Mp, Np, Kp, alpha, dA, arp, dB, brp, beta, dC, Mp, splitk);
        download2d(c, C, ldc, dC, Mp, M, N);
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
    } catch (...) {
        dfree(dA); dfree(dB); dfree(dC);
        throw;
    }
    dfree(dA); dfree(dB); dfree(dC);
    API_END
}

static void upload_spd_padded(geordd_ctx* c, double* dA, int npad, const double* A, int n) {
    GD_CUDA(cudaMemsetAsync(dA, 0, (size_t)npad * npad * sizeof(double), c->stream));
    upload2d(c, dA, npad, A, n, n, n);
    if (npad > n) launch_add_diag(c->stream, dA + (size_t)n * npad + n, npad, npad - n, 1.0);
}

int geordd_dbg_potrf(geordd_ctx* c, const double* A, int n, double* L, double* ms) {
    API_BEGIN(c)
    if (!A || !L || n <= 0) throw ApiError(GEORDD_ERR_ARG, "dbg_potrf: bad arguments");
    const int npad = round_up(n, TILE);
    double *dA = nullptr, *dL = nullptr;
    int info = 0;
    try {
        dA = dalloc((size_t)npad * npad, c->stream, false);
        dL = dalloc(factor_elems(npad), c->stream, true);
        upload_spd_padded(c, dA, npad, A, n);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        GD_CUDA(cudaStreamSynchronize(c->stream));
        cudaEventRecord(e0, c->stream);
        launch_potrf(c->stream, c->sched, dA, dL, npad, npad);
        cudaEventRecord(e1, c->stream);
        GD_CUDA(cudaMemcpyAsync(c->h_info, c->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ms) *ms = t;
        info = c->h_info[0];
        GD_CUDA(cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream));
        if (info < 0) throw ApiError(GEORDD_ERR_WATCHDOG, "potrf watchdog fired");
        launch_factor_unpack(c->stream, dL, npad, dA, n, n, false);   // dense lower factor (reuse dA as staging)
        GD_CUDA(cudaMemcpyAsync(L, dA, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(dA); dfree(dL);
        throw;
    }
    dfree(dA); dfree(dL);
    if (info > 0) return info;
    API_END
}


int geordd_dbg_trsm(geordd_ctx* c, int mode, const double* L, int n, double* B, int nrhs, double* ms) {
    API_BEGIN(c)
    // mode 0 / 1 / 2: forward, backward, multiply; 3 / 4: forward / backward with the right-hand sides in ZB storage
    if (!L || !B || n <= 0 || nrhs <= 0 || mode < 0 || mode > 4) throw ApiError(GEORDD_ERR_ARG, "dbg_trsm: bad arguments");
    const bool xb = mode >= 3;
    if (xb) mode -= 3;
    const int npad = round_up(n, TILE), ncols = round_up(nrhs, 64);
    double *dL = nullptr, *dB = nullptr, *dO = nullptr;
    try {
        dL = dalloc(factor_elems(npad), c->stream, false);
        dB = dalloc((size_t)npad * ncols, c->stream, true);
        c->tmp0.ensure((size_t)n * n, c->stream, false);
        GD_CUDA(cudaMemcpyAsync(c->tmp0.p, L, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        launch_factor_pack(c->stream, c->tmp0.p, n, n, dL, npad);
        launch_mirror_factor(c->stream, dL, npad);
        upload2d(c, dB, npad, B, n, n, nrhs);
        SolveEpilogue epi;
        double* dIn = dB;
        if (mode == 2) {   // triangular multiply reads its input in ZB storage
            dO = dalloc((size_t)npad * ncols, c->stream, true);
            epi.out0 = dO; epi.ld0 = npad; epi.nrows = n; epi.split = n;
            c->tmp1.ensure(zb_elems(npad, ncols), c->stream, false);
            launch_zb_pack(c->stream, dB, npad, n, nrhs, c->tmp1.p, npad, ncols);
            dIn = c->tmp1.p;
        }
        if (xb) {
            epi.xb = true;
            c->tmp1.ensure(zb_elems(npad, ncols), c->stream, false);
            launch_zb_pack(c->stream, dB, npad, n, nrhs, c->tmp1.p, npad, ncols);
            dIn = c->tmp1.p;
        }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        GD_CUDA(cudaStreamSynchronize(c->stream));
        cudaEventRecord(e0, c->stream);
        launch_solve(c->stream, c->sched, (SolveMode)mode, dL, npad, dIn, npad, ncols, epi, xb ? 0 : nrhs);
        cudaEventRecord(e1, c->stream);
        check_abort(c);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ms) *ms = t;
        if (xb) launch_zb_unpack(c->stream, dIn, npad, dB, npad, n, nrhs, 0);
        download2d(c, B, n, mode == 2 ? dO : dB, npad, n, nrhs);
        GD_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        dfree(dL); dfree(dB); dfree(dO);
        throw;
    }
    dfree(dL); dfree(dB); dfree(dO);
    API_END
}

}  // extern "C"
