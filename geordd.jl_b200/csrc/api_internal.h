// Internal host-side structures behind the opaque handles of include/geordd_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>
#include <stdexcept>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "../../include/geordd_b200.h"
#include "kernels.h"
#include "smallops.h"

namespace geordd {

enum ProfClass { PC_COV = 0, PC_POTRF, PC_FWD, PC_BWD, PC_TRMM, PC_GEMM, PC_RNG, PC_FINISH, PC_LU, PC_MISC, PC_SEQ, PC_COUNT };
extern const char* const kProfNames[PC_COUNT];

struct ApiError : std::runtime_error {
    int code;
    ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define GD_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            throw geordd::ApiError(GEORDD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// Device memory comes from a per-context caching pool (api_core.cu): fits and null handles are created and
// destroyed repeatedly by callers (placebo sweeps, boot_* drivers), cudaMalloc/cudaFree would serialise them.
double* dalloc(size_t count, cudaStream_t st, bool zero);
void dfree(void* p);

struct DevBuf {
    double* p = nullptr;
    size_t n = 0;
    void ensure(size_t count, cudaStream_t st, bool zero) {
        if (count > n) {
            release();
            p = dalloc(count, st, zero);
            n = count;
        }
    }
    void release() {
        if (p) dfree(p);
        p = nullptr;
        n = 0;
    }
};

}  // namespace geordd

struct geordd_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    geordd::Sched sched{};
    std::string err;
    long max_batch = 0;
    bool mll_forward_only = false;
    double min_margin[3] = {INFINITY, INFINITY, INFINITY};   // last null call: min |stat - threshold| for chi2, invvar p, mll
    void* comm = nullptr;                     // ncclComm_t of a one-process-per-GPU caller (geordd_comm_init_rank), api_multi.cu
    int comm_size = 0;
    // caching device-memory pool
    std::multimap<size_t, void*> pool_free;
    std::unordered_map<void*, size_t> pool_live;
    size_t pool_cached_bytes = 0;
    // small device / pinned-host scratch
    double* d_scal = nullptr;                 // 64 doubles
    double* h_scal = nullptr;                 // pinned
    int* d_info = nullptr;
    int* h_info = nullptr;                    // pinned, 4 ints
    unsigned long long* d_tally = nullptr;    // 8 counters
    unsigned long long* h_tally = nullptr;    // pinned
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // side streams: independent per-GP tails of a batched fit
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    geordd::DevBuf ws;                        // split-K workspace
    geordd::DevBuf tmp0, tmp1, tmp2, tmp3;    // generic scratch matrices
    geordd::DevBuf ws2, tmp4, tmp5;           // side-stream twins (cliff face: the control side runs beside the treated side)
    // profiling
    bool prof_on = false;
    struct Rec { int cls; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[geordd::PC_COUNT] = {0};
    long long prof_n[geordd::PC_COUNT] = {0};
};

struct geordd_gp {
    geordd_ctx* ctx = nullptr;
    geordd_kernel k{};
    double log_noise = 0.0;
    int n = 0, npad = 0, dim = 0;
    double* dX = nullptr;       // dim x n
    double* dK = nullptr;       // npad x npad  (cK.mat, lower tiles; full once mirrored)
    double* dL = nullptr;       // lower Cholesky factor (= U') in TBP storage, factor_elems(npad) doubles
    double* dDinv = nullptr;    // inverse 32x32 diagonal blocks of dL (built on first use by the null simulations)
    double* dR = nullptr;       // npad         y - m (zero padded)
    double* dAlpha = nullptr;   // npad x 64 slab, column 0 = alpha
    bool k_full = false;
    double logdet = 0.0, mll = 0.0;
    int jitter_rounds = 0;
    std::vector<double> y;      // host copy of gp.y (make_null concatenates them)
    std::vector<double> X;      // host copy of gp.x
};

struct geordd_null {
    geordd_ctx* ctx = nullptr;
    geordd_gp *T = nullptr, *C = nullptr;
    geordd_gp* N = nullptr;     // pooled null GP (owned)
    int nb = 0, nbp = 0;
    double nugget = 0.0;
    int jitter_rounds = 0;
    double* dXb = nullptr;      // 2 x nb
    double* dKbT = nullptr;     // nbp x mpadT : cov(kernel_T, Xb, X_T)
    double* dKbC = nullptr;     // nbp x mpadC
    double* dSig = nullptr;     // nbp x nbp raw cliff covariance (+ nugget, + jitter) = Sigma_raw
    double* dSigRaw = nullptr;  // nbp x nbp cliff covariance + nugget BEFORE make_posdef! jitter (LU solves of the calibration)
    double* dMuObs = nullptr;   // cliff-face mean of the OBSERVED outcomes at the sentinels (geordd_null_chistat_obs / _invvar_obs)
    double* dLSig = nullptr;    // its lower factor
    double* dOnes = nullptr;    // nbp (1 on rows < nb)
    double denom = 0.0;         // sum(Sigma \ 1) with the PD factor
    // hoisted draw-independent part of the analytic calibration (src/hypotest_invvar.jl:121-131)
    bool calib_ready[2] = {false, false};
    double cov_mu_tau[2] = {0.0, 0.0};   // [0] corrected, [1] as shipped
    bool calib_lu_ready = false;         // 3-argument pval_invvar_calib (:74-105): LU of Sigma + nugget, corrected formula
    double cov_mu_tau_lu = 0.0;
    // simulation buffers, sized for cap_cols draws
    long cap_cols = 0;
    geordd::DevBuf Z, RN, AN, AT, AC, M, W, partN, partT, partC, partChi, partNum;
    geordd::DevBuf oChi, oPval, oNum, oNull, oAlt, oNumLU;   // per-call output staging (nsim)
};

struct geordd_multigp {
    geordd_ctx* ctx = nullptr;
    int n = 0, npad = 0, p = 0, ppad = 0, dim = 0;
    std::vector<int> group_off;
    double* dD = nullptr;       // ppad x npad (zero padded), column-major (p contiguous)
    double* dX = nullptr;       // dim x n
    double* dG = nullptr;       // npad x npad  D'D
    double* dK = nullptr;       // npad x npad
    double* dL = nullptr;
    double* dB = nullptr;       // npad x npad  scratch (K^-1 / alpha alpha' - K^-1)
    double* dY = nullptr;       // npad
    double* dR = nullptr;       // npad
    double* dAlpha = nullptr;   // npad x 64
    double logdet = 0.0;
};

namespace geordd {

// ---- helpers implemented in api_core.cu ------------------------------------------------------------
struct ProfScope {
    geordd_ctx* c;
    int cls = 0;
    int idx = -1;
    ProfScope(geordd_ctx* ctx, int cls);
    ~ProfScope();
};
void ctx_use(geordd_ctx* c);                       // cudaSetDevice
int ctx_fail(geordd_ctx* c, const ApiError& e);    // records message, returns code
void check_abort(geordd_ctx* c);                   // sync + throw on watchdog
double read_scalar(geordd_ctx* c, const double* d);   // sync D2H of one double
void upload2d(geordd_ctx* c, double* dst, long ldd, const double* src, long lds, int rows, int cols);
void download2d(geordd_ctx* c, double* dst, long ldd, const double* src, long lds, int rows, int cols);
void ensure_ws(geordd_ctx* c, size_t count);
void ensure_side_streams(geordd_ctx* c);   // c->side[], c->ev_fork, c->ev_join[] (created on first use)
long long g_guard_violations_get();

KernelSpecDev to_dev(const geordd_kernel& k);
// make_posdef!(K) -> L (App. A.5).  K (npad x npad, lower tiles valid, logical size n) is jittered in
// place.  Returns 0 or the LAPACK info after the 11th failed attempt.
int make_posdef(geordd_ctx* c, double* K, double* L, long ld, int n, int npad, int* rounds, const double* Lpre = nullptr,
                int npad_pre = 0, int ptiles = 0);
// B <- L^-T L^-1 B (ncols multiple of 64) with optional fused epilogue on the backward pass.
void pd_solve(geordd_ctx* c, const double* L, int npad, double* B, long ldb, int ncols, const SolveEpilogue& epi_bwd,
              int live_cols = 0);
void gemm(geordd_ctx* c, bool a_mc, bool b_nc, int M, int N, int K, double alpha, const double* A, long lda,
          const double* B, long ldb, double beta, double* C, long ldc, int splitk, bool b_xb = false);
int pick_splitk(geordd_ctx* c, int M, int N, int K);

// `lead` (optional): a fitted GP over the FIRST lead->n units of X with the same kernel and noise and no jitter; the tiles
// of its factor that lie inside the leading block are reused (make_null: pooled = [treated control]).
geordd_gp* gp_create(geordd_ctx* c, const geordd_kernel* k, const double* X, int dim, int n, const double* y,
                     double log_noise, int* info_out, const geordd_gp* lead = nullptr);
void gp_create_batch(geordd_ctx* c, int nfit, const geordd_kernel* ks, const double* const* Xs, int dim, const int* ns,
                     const double* const* ys, const double* log_noise, geordd_gp** gps, int* infos);
void gp_update_alpha(geordd_gp* gp);   // alpha from dR, mll
void gp_free(geordd_gp* gp);
void gp_ensure_full_K(geordd_gp* gp);
const double* gp_diag_inverse(geordd_gp* gp);   // builds gp->dDinv on first use; nullptr when switched off (GEORDD_SEQ_SUBST=1)

// cliff face on device: mu (nbp), Sigma (nbp x nbp, ld nbp), exactly symmetric.
void cliff_face_dev(geordd_gp* T, geordd_gp* C, const double* dXb, int nb, int nbp, double* dMu, double* dSigma);
// mu_b = m(Xb) + cov(k, Xb, X) alpha on device for one side (writes nbp vector into column 0 of dst slab)
void predict_mu_dev(geordd_gp* gp, const double* dKb /*nbp x npad or null*/, const double* dXb, int nb, int nbp,
                    double* dMuSlab /* nbp x 64 */);

}  // namespace geordd
