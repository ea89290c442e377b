// C ABI, part 4: multi-GPU inside the library (SURVEY §5 / §8e, App. D `geordd_ctx_create(ndev, devs)`).
//
// The only parallel structure of the hot path is the embarrassingly parallel draw loop of the null tests
// (`[sim...() for _ in 1:nsim]`, src/hypotest_chi2.jl:49-50, src/hypotest_mll.jl:27-28, src/hypotest_invvar.jl:60-61,
// src/power.jl:67-71).  A GROUP lets ONE host process (one Julia task, one C program) drive all GPUs of a box:
//   * fits are replicated: every device assembles and factors its own copy (deterministic kernels, so the replicas are
//     bit-identical), one host thread per device;
//   * draws are sharded contiguously by index; a draw depends only on (seed, draw index), so results do not depend on the
//     number of devices; per-draw statistics land directly in the caller's host arrays (each device owns a slice);
//   * the exceedance tallies are summed with ONE ncclAllReduce(ncclSum, ncclInt64) over NVLink (ncclCommInitAll).
// For one-process-per-GPU callers (MPI-launched Julia, torchrun) the same reduce is exposed per context:
// geordd_comm_unique_id / geordd_comm_init_rank / geordd_comm_allreduce_tally.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch process this resolves to the NCCL torch already
// loaded, in a plain C / Julia process to the system library.  No NCCL, no group with more than one device: error, not a
// silent host-side sum.
#include "api_internal.h"
#include <dlfcn.h>
#include <algorithm>
#include <functional>
#include <mutex>
#include <thread>

namespace {

// ---- the handful of NCCL entry points, resolved with dlsym (types as in nccl.h 2.x) -----------------------------
struct NcclId { char internal[128]; };
using NcclComm = void*;
constexpr int kNcclInt64 = 4, kNcclSum = 0;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p && api.why.empty()) api.why = std::string("missing NCCL symbol ") + n; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

bool nccl_ok(std::string* why) {
    NcclApi& a = nccl();
    if (a.lib && a.why.empty()) return true;
    if (why) *why = a.why;
    return false;
}

std::string nccl_err(int rc) {
    NcclApi& a = nccl();
    return std::string("NCCL error ") + std::to_string(rc) + (a.GetErrorString ? std::string(": ") + a.GetErrorString(rc) : "");
}

}  // namespace

struct geordd_group {
    int ndev = 0;
    std::vector<geordd_ctx*> ctx;
    std::vector<NcclComm> comm;                 // empty for a single device
    std::vector<unsigned long long*> d_tally;   // per device, 8 counters
    std::string err;
};
struct geordd_ggp {
    geordd_group* g = nullptr;
    std::vector<geordd_gp*> gp;
};
struct geordd_gnull {
    geordd_group* g = nullptr;
    std::vector<geordd_null*> h;
    long nobs = 0;
};

namespace {

// f(d) -> status on one host thread per device; first failure wins and its context's message is recorded.
int group_parallel(geordd_group* g, const std::function<int(int)>& f) {
    std::vector<int> rc(g->ndev, 0);
    if (g->ndev == 1) {
        rc[0] = f(0);
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < g->ndev; ++d) th.emplace_back([&, d] { rc[d] = f(d); });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < g->ndev; ++d)
        if (rc[d] != 0) {
            g->err = "device " + std::to_string(g->ctx[d]->device) + ": " + geordd_last_error(g->ctx[d]);
            return rc[d];
        }
    return 0;
}

void shard(long nsim, int d, int ndev, long* lo, long* cnt) {   // the same contiguous split as parallel.shard_draws
    *lo = nsim * d / ndev;
    *cnt = nsim * (d + 1) / ndev - *lo;
}

// Sum `n` (<= 8) per-device counters over the group: ONE all-reduce of int64 over NVLink.  vals[d*8 + k].
int group_allreduce(geordd_group* g, std::vector<long long>& vals, int n) {
    if (g->ndev == 1) return 0;
    NcclApi& a = nccl();
    for (int d = 0; d < g->ndev; ++d) {
        cudaSetDevice(g->ctx[d]->device);
        if (cudaMemcpyAsync(g->d_tally[d], &vals[(size_t)d * 8], n * sizeof(long long), cudaMemcpyHostToDevice, g->ctx[d]->stream) != cudaSuccess) {
            g->err = "cudaMemcpyAsync(tally) failed";
            return GEORDD_ERR_CUDA;
        }
    }
    int rc = a.GroupStart();
    for (int d = 0; d < g->ndev && rc == 0; ++d)
        rc = a.AllReduce(g->d_tally[d], g->d_tally[d], (size_t)n, kNcclInt64, kNcclSum, g->comm[d], g->ctx[d]->stream);
    int rc2 = a.GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) { g->err = nccl_err(rc); return GEORDD_ERR_NCCL; }
    for (int d = 0; d < g->ndev; ++d) {
        cudaSetDevice(g->ctx[d]->device);
        cudaMemcpyAsync(&vals[(size_t)d * 8], g->d_tally[d], n * sizeof(long long), cudaMemcpyDeviceToHost, g->ctx[d]->stream);
        if (cudaStreamSynchronize(g->ctx[d]->stream) != cudaSuccess) { g->err = "tally all-reduce failed"; return GEORDD_ERR_CUDA; }
    }
    return 0;
}

}  // namespace

extern "C" {

int geordd_group_create(geordd_group** out, int ndev, const int* devices) {
    if (!out || ndev < 1 || ndev > 64) return GEORDD_ERR_ARG;
    *out = nullptr;
    geordd_group* g = new geordd_group();
    g->ndev = ndev;
    std::vector<int> devs(ndev);
    for (int d = 0; d < ndev; ++d) devs[d] = devices ? devices[d] : d;
    for (int d = 0; d < ndev; ++d) {
        geordd_ctx* c = nullptr;
        int rc = geordd_ctx_create(&c, devs[d]);
        if (rc != 0) { geordd_group_destroy(g); return rc; }
        g->ctx.push_back(c);
        unsigned long long* t = nullptr;
        cudaSetDevice(devs[d]);
        if (cudaMalloc(&t, 8 * sizeof(unsigned long long)) != cudaSuccess) { geordd_group_destroy(g); return GEORDD_ERR_NOMEM; }
        g->d_tally.push_back(t);
    }
    if (ndev > 1) {
        std::string why;
        if (!nccl_ok(&why)) { geordd_group_destroy(g); return GEORDD_ERR_NCCL; }
        g->comm.assign(ndev, nullptr);
        int rc = nccl().CommInitAll(g->comm.data(), ndev, devs.data());
        if (rc != 0) { g->comm.clear(); geordd_group_destroy(g); return GEORDD_ERR_NCCL; }
    }
    *out = g;
    return 0;
}

int geordd_group_destroy(geordd_group* g) {
    if (!g) return 0;
    for (auto cm : g->comm) if (cm) nccl().CommDestroy(cm);
    for (size_t d = 0; d < g->d_tally.size(); ++d) { cudaSetDevice(g->ctx[d]->device); cudaFree(g->d_tally[d]); }
    for (auto c : g->ctx) geordd_ctx_destroy(c);
    delete g;
    return 0;
}

int geordd_group_size(const geordd_group* g) { return g ? g->ndev : GEORDD_ERR_ARG; }
const char* geordd_group_last_error(const geordd_group* g) { return g ? g->err.c_str() : "no group"; }
geordd_ctx* geordd_group_ctx(geordd_group* g, int i) { return (g && i >= 0 && i < g->ndev) ? g->ctx[i] : nullptr; }

int geordd_group_gp_fit(geordd_group* g, const geordd_kernel* k, const double* X, int dim, int n, const double* y, double log_noise,
                        geordd_ggp** out, double* alpha, double* mll, int* jitter_rounds) {
    if (!g || !out) return GEORDD_ERR_ARG;
    *out = nullptr;
    geordd_ggp* h = new geordd_ggp();
    h->g = g;
    h->gp.assign(g->ndev, nullptr);
    std::vector<double> mlls(g->ndev, 0.0);
    std::vector<int> jr(g->ndev, 0);
    int rc = group_parallel(g, [&](int d) {
        return geordd_gp_fit(g->ctx[d], k, X, dim, n, y, log_noise, &h->gp[d], d == 0 ? alpha : nullptr, &mlls[d], nullptr, &jr[d]);
    });
    if (rc != 0) { geordd_group_gp_destroy(h); return rc; }
    if (mll) *mll = mlls[0];
    if (jitter_rounds) *jitter_rounds = jr[0];
    *out = h;
    return 0;
}

int geordd_group_gp_destroy(geordd_ggp* h) {
    if (!h) return 0;
    for (auto gp : h->gp) if (gp) geordd_gp_destroy(gp);
    delete h;
    return 0;
}

geordd_gp* geordd_group_gp_replica(geordd_ggp* h, int i) { return (h && i >= 0 && i < (int)h->gp.size()) ? h->gp[i] : nullptr; }

int geordd_group_null_prepare(geordd_ggp* T, geordd_ggp* C, const double* Xb, int nb, double nugget, geordd_gnull** out,
                              int* jitter_rounds, double* mll_null) {
    if (!T || !C || !out || T->g != C->g) return GEORDD_ERR_ARG;
    *out = nullptr;
    geordd_group* g = T->g;
    geordd_gnull* h = new geordd_gnull();
    h->g = g;
    h->h.assign(g->ndev, nullptr);
    h->nobs = geordd_gp_nobs(T->gp[0]) + geordd_gp_nobs(C->gp[0]);
    std::vector<double> m(g->ndev, 0.0);
    std::vector<int> jr(g->ndev, 0);
    int rc = group_parallel(g, [&](int d) { return geordd_null_prepare(T->gp[d], C->gp[d], Xb, nb, nugget, &h->h[d], &jr[d], &m[d]); });
    if (rc != 0) { geordd_group_null_destroy(h); return rc; }
    if (jitter_rounds) *jitter_rounds = jr[0];
    if (mll_null) *mll_null = m[0];
    *out = h;
    return 0;
}

int geordd_group_null_destroy(geordd_gnull* h) {
    if (!h) return 0;
    for (auto n : h->h) if (n) geordd_null_destroy(n);
    delete h;
    return 0;
}

geordd_null* geordd_group_null_replica(geordd_gnull* h, int i) { return (h && i >= 0 && i < (int)h->h.size()) ? h->h[i] : nullptr; }

int geordd_group_null_mll(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                          double dmll_obs, double* mll_null, double* mll_alt, long long* n_exceed) {
    if (!h || nsim <= 0) return GEORDD_ERR_ARG;
    geordd_group* g = h->g;
    std::vector<long long> cnt((size_t)g->ndev * 8, 0);
    int rc = group_parallel(g, [&](int d) {
        long lo, c;
        shard(nsim, d, g->ndev, &lo, &c);
        if (c == 0) return 0;
        return geordd_null_mll(h->h[d], c, seed, draw0 + (unsigned long long)lo, Z ? Z + (size_t)lo * h->nobs : nullptr, dmll_obs,
                               mll_null ? mll_null + lo : nullptr, mll_alt ? mll_alt + lo : nullptr, &cnt[(size_t)d * 8], nullptr, nullptr);
    });
    if (rc != 0) return rc;
    rc = group_allreduce(g, cnt, 1);
    if (rc != 0) return rc;
    if (g->ndev == 1 && n_exceed) *n_exceed = cnt[0];
    else if (n_exceed) *n_exceed = cnt[0];   // every device holds the same sum after the all-reduce
    return 0;
}

int geordd_group_null_chi2(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                           double chi_obs, double* stats, long long* n_exceed) {
    if (!h || nsim <= 0) return GEORDD_ERR_ARG;
    geordd_group* g = h->g;
    std::vector<long long> cnt((size_t)g->ndev * 8, 0);
    int rc = group_parallel(g, [&](int d) {
        long lo, c;
        shard(nsim, d, g->ndev, &lo, &c);
        if (c == 0) return 0;
        return geordd_null_chi2(h->h[d], c, seed, draw0 + (unsigned long long)lo, Z ? Z + (size_t)lo * h->nobs : nullptr, chi_obs,
                                stats ? stats + lo : nullptr, &cnt[(size_t)d * 8]);
    });
    if (rc != 0) return rc;
    rc = group_allreduce(g, cnt, 1);
    if (rc != 0) return rc;
    if (n_exceed) *n_exceed = cnt[0];
    return 0;
}

int geordd_group_null_invvar(geordd_gnull* h, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                             double pval_obs, double* pvals, long long* n_below) {
    if (!h || nsim <= 0) return GEORDD_ERR_ARG;
    geordd_group* g = h->g;
    std::vector<long long> cnt((size_t)g->ndev * 8, 0);
    int rc = group_parallel(g, [&](int d) {
        long lo, c;
        shard(nsim, d, g->ndev, &lo, &c);
        if (c == 0) return 0;
        return geordd_null_invvar(h->h[d], c, seed, draw0 + (unsigned long long)lo, Z ? Z + (size_t)lo * h->nobs : nullptr, pval_obs,
                                  pvals ? pvals + lo : nullptr, &cnt[(size_t)d * 8]);
    });
    if (rc != 0) return rc;
    rc = group_allreduce(g, cnt, 1);
    if (rc != 0) return rc;
    if (n_below) *n_below = cnt[0];
    return 0;
}

int geordd_group_power(geordd_gnull* h, double tau, long nsim, unsigned long long seed, unsigned long long draw0, const double* Z,
                       const double* chi_null, const double* mll_null, const double* pinv_null, long nnull, int compat_calib_bug,
                       double* out5) {
    if (!h || nsim <= 0 || !out5) return GEORDD_ERR_ARG;
    geordd_group* g = h->g;
    return group_parallel(g, [&](int d) {
        long lo, c;
        shard(nsim, d, g->ndev, &lo, &c);
        if (c == 0) return 0;
        return geordd_power(h->h[d], tau, c, seed, draw0 + (unsigned long long)lo, Z ? Z + (size_t)lo * h->nobs : nullptr, chi_null,
                            mll_null, pinv_null, nnull, compat_calib_bug, out5 + 5 * lo);
    });
}

// ---- one process per GPU: the same reduce on a single context ------------------------------------------------------
int geordd_comm_unique_id(char* id128) {
    if (!id128) return GEORDD_ERR_ARG;
    if (!nccl_ok(nullptr)) return GEORDD_ERR_NCCL;
    NcclId id;
    int rc = nccl().GetUniqueId(&id);
    if (rc != 0) return GEORDD_ERR_NCCL;
    std::memcpy(id128, id.internal, 128);
    return 0;
}

int geordd_comm_init_rank(geordd_ctx* c, int nranks, int rank, const char* id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return GEORDD_ERR_ARG;
    std::string why;
    if (!nccl_ok(&why)) { c->err = why; return GEORDD_ERR_NCCL; }
    if (c->comm) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    NcclId id;
    std::memcpy(id.internal, id128, 128);
    cudaSetDevice(c->device);
    int rc = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (rc != 0) { c->err = nccl_err(rc); c->comm = nullptr; return GEORDD_ERR_NCCL; }
    c->comm_size = nranks;
    return 0;
}

int geordd_comm_allreduce_tally(geordd_ctx* c, long long* vals, int n) {
    if (!c || !vals || n < 1 || n > 8) return GEORDD_ERR_ARG;
    if (!c->comm) { c->err = "geordd_comm_init_rank was not called on this context"; return GEORDD_ERR_ARG; }
    cudaSetDevice(c->device);
    if (cudaMemcpyAsync(c->d_tally, vals, n * sizeof(long long), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return GEORDD_ERR_CUDA;
    int rc = nccl().AllReduce(c->d_tally, c->d_tally, (size_t)n, kNcclInt64, kNcclSum, c->comm, c->stream);
    if (rc != 0) { c->err = nccl_err(rc); return GEORDD_ERR_NCCL; }
    cudaMemcpyAsync(vals, c->d_tally, n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "tally all-reduce failed"; return GEORDD_ERR_CUDA; }
    return 0;
}

int geordd_comm_destroy(geordd_ctx* c) {
    if (!c) return 0;
    if (c->comm) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    c->comm_size = 0;
    return 0;
}

}  // extern "C"
