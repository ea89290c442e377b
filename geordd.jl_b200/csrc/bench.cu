// Device-resident micro-benchmarks exposed through the C ABI (used by bench.py / profiles):
//   geordd_bench_dmma_peak : register-resident mma.sync.m8n8k4.f64 loop = the FP64 tensor-pipe
//                            ceiling on this part (MEASURED_PEAKS.json holds no FP64 figure).
//   geordd_bench_potrf     : the dataflow Cholesky on a synthetic SPD covariance built on device.
#include "api_internal.h"
#include "common.cuh"
#include "tileops.cuh"
#include <algorithm>
#include <random>

namespace geordd {

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) sink[0] = s;
}

}  // namespace geordd

using namespace geordd;

namespace geordd {
// Cycle counts of the in-shared-memory tile operations that sit on the critical path of the dataflow Cholesky and
// solves (one CTA, clock64 around each op, best of `reps`).
__global__ void __launch_bounds__(NTHREADS, 1) tileops_cycles_kernel(const double* __restrict__ Aimg, const double* __restrict__ Limg,
                                                                      int reps, long long* out) {
    extern __shared__ __align__(16) double smem[];
    double* Cs = smem;
    double* Lp = Cs + TILE * SC;
    double* invd = Lp + 2 * NB * SC;   // two panel buffers (right solve) >= TILE*LRS (row panel of the backward solve)
    __shared__ int sh_info;
    long long best[4] = {1LL << 62, 1LL << 62, 1LL << 62, 1LL << 62};
    __shared__ long long prof[8];
    if (threadIdx.x < 8) prof[threadIdx.x] = 0;
    for (int op = 0; op < 4; ++op) {
        for (int r = 0; r < reps; ++r) {
            for (int idx = threadIdx.x; idx < TILE * SC; idx += NTHREADS) Cs[idx] = Aimg[idx];
            if (threadIdx.x == 0) sh_info = 0;
            __syncthreads();
            long long t0 = clock64();
            if (op == 0) potrf_tile(Cs, invd, &sh_info, r == reps - 1 ? prof : nullptr);
            if (op == 1) trsm_right_tile(Cs, Lp, Lp + NB * SC, invd, Limg, TLD, r == reps - 1 ? prof + 3 : nullptr);
            if (op == 2) solve_fwd_tile(Cs, 64, Lp, invd, Limg, TLD);
            if (op == 3) solve_bwd_tile(Cs, 64, Lp, invd, Limg, TLD);
            __syncthreads();
            long long t1 = clock64();
            if (t1 - t0 < best[op]) best[op] = t1 - t0;
        }
    }
    if (threadIdx.x == 0)
        for (int op = 0; op < 4; ++op) out[op] = best[op];
    __syncthreads();
    if (threadIdx.x < 6) out[4 + threadIdx.x] = prof[threadIdx.x];
}
}  // namespace geordd
using namespace geordd;

extern "C" {

int geordd_bench_dmma_peak(geordd_ctx* c, double* tflops) {
    if (!c || !tflops) return GEORDD_ERR_ARG;
    try {
        ctx_use(c);
        const int iters = 4096, ctas = c->sched.num_sms * 4;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        dmma_peak_kernel<<<ctas, 256, 0, c->stream>>>(64, c->d_scal + 8); ++g_kernel_launches;
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0, c->stream);
            dmma_peak_kernel<<<ctas, 256, 0, c->stream>>>(iters, c->d_scal + 8); ++g_kernel_launches;
            cudaEventRecord(e1, c->stream);
            GD_CUDA(cudaStreamSynchronize(c->stream));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        GD_CUDA(cudaGetLastError());
        double flops = 2.0 * 256.0 * 16.0 * (double)iters * 8.0 * (double)ctas;   // 8 warps x 16 MMAs x 8x8x4 FMA
        *tflops = flops / (best * 1e-3) / 1e12;
    } catch (const ApiError& e) {
        return ctx_fail(c, e);
    }
    return 0;
}

// DMMA issue-rate probe: `threads` per CTA, `ctas` CTAs in total, 16 independent accumulators per warp.
int geordd_bench_dmma(geordd_ctx* c, int threads, int ctas, int iters, double* tflops) {
    if (!c || !tflops || threads <= 0 || threads > 1024 || threads % 32 || ctas <= 0 || iters <= 0) return GEORDD_ERR_ARG;
    try {
        ctx_use(c);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        dmma_peak_kernel<<<ctas, threads, 0, c->stream>>>(16, c->d_scal + 8); ++g_kernel_launches;
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0, c->stream);
            dmma_peak_kernel<<<ctas, threads, 0, c->stream>>>(iters, c->d_scal + 8); ++g_kernel_launches;
            cudaEventRecord(e1, c->stream);
            GD_CUDA(cudaStreamSynchronize(c->stream));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        GD_CUDA(cudaGetLastError());
        double flops = 2.0 * 256.0 * 16.0 * (double)iters * (threads / 32) * (double)ctas;
        *tflops = flops / (best * 1e-3) / 1e12;
    } catch (const ApiError& e) {
        return ctx_fail(c, e);
    }
    return 0;
}

int geordd_bench_potrf(geordd_ctx* c, int n, int reps, double* ms_best, double* ms_mean) {
    if (!c || n <= 0 || reps <= 0) return GEORDD_ERR_ARG;
    double *dX = nullptr, *dK = nullptr, *dL = nullptr;
    try {
        ctx_use(c);
        const int npad = round_up(n, TILE);
        std::mt19937_64 rng(1);
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::vector<double> X((size_t)2 * n);
        for (auto& v : X) v = U(rng);
        dX = dalloc((size_t)2 * n, c->stream, false);
        dK = dalloc((size_t)npad * npad, c->stream, false);
        dL = dalloc(factor_elems(npad), c->stream, false);
        GD_CUDA(cudaMemcpyAsync(dX, X.data(), X.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        KernelSpecDev spec{2, 0.3, 1.0, 400.0};
        launch_cov(c->stream, spec, dX, n, dX, n, 2, true, false, 0.09, dK, npad, npad, npad, false);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        double best = 1e30, sum = 0.0;
        for (int r = 0; r < reps + 1; ++r) {
            GD_CUDA(cudaStreamSynchronize(c->stream));
            cudaEventRecord(e0, c->stream);
            {
                ProfScope ps(c, PC_POTRF);
                launch_potrf(c->stream, c->sched, dK, dL, npad, npad);
            }
            cudaEventRecord(e1, c->stream);
            GD_CUDA(cudaMemcpyAsync(c->h_info, c->sched.abort_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            GD_CUDA(cudaStreamSynchronize(c->stream));
            GD_CUDA(cudaGetLastError());
            if (c->h_info[0] != 0) {
                int info = c->h_info[0];
                cudaMemsetAsync(c->sched.abort_word, 0, sizeof(int), c->stream);
                throw ApiError(info < 0 ? GEORDD_ERR_WATCHDOG : GEORDD_ERR_ARG, "bench_potrf: factorization failed");
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0) { best = std::min(best, (double)ms); sum += ms; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ms_best) *ms_best = best;
        if (ms_mean) *ms_mean = sum / reps;
    } catch (const ApiError& e) {
        dfree(dX); dfree(dK); dfree(dL);
        return ctx_fail(c, e);
    }
    dfree(dX); dfree(dK); dfree(dL);
    return 0;
}

int geordd_bench_tileops(geordd_ctx* c, double* cycles4) {
    if (!c || !cycles4) return GEORDD_ERR_ARG;
    double *dX = nullptr, *dK = nullptr, *dL = nullptr, *dA = nullptr;
    long long* dOut = nullptr;
    try {
        ctx_use(c);
        const int n = TILE;
        std::mt19937_64 rng(1);
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::vector<double> X((size_t)2 * n);
        for (auto& v : X) v = U(rng);
        dX = dalloc((size_t)2 * n, c->stream, false);
        dK = dalloc((size_t)n * n, c->stream, false);
        dL = dalloc(factor_elems(n), c->stream, false);
        dA = dalloc(factor_elems(n), c->stream, true);
        dOut = reinterpret_cast<long long*>(dalloc(10, c->stream, true));
        GD_CUDA(cudaMemcpyAsync(dX, X.data(), X.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        KernelSpecDev spec{2, 0.3, 1.0, 400.0};
        launch_cov(c->stream, spec, dX, n, dX, n, 2, true, true, 0.09, dK, n, n, n, false);
        launch_potrf(c->stream, c->sched, dK, dL, n, n);
        GD_CUDA(cudaMemcpy2DAsync(dA, TLD * sizeof(double), dK, n * sizeof(double), n * sizeof(double), n, cudaMemcpyDeviceToDevice, c->stream));
        const size_t smem = (size_t)(TILE * SC + 2 * NB * SC + TILE) * sizeof(double);
        cudaFuncSetAttribute(tileops_cycles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tileops_cycles_kernel<<<1, NTHREADS, smem, c->stream>>>(dA, dL, 5, dOut);
        long long h[10];
        GD_CUDA(cudaMemcpyAsync(h, dOut, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        GD_CUDA(cudaStreamSynchronize(c->stream));
        GD_CUDA(cudaGetLastError());
        for (int i = 0; i < 10; ++i) cycles4[i] = (double)h[i];
    } catch (const ApiError& e) {
        dfree(dX); dfree(dK); dfree(dL); dfree(dA); dfree(dOut);
        return ctx_fail(c, e);
    }
    dfree(dX); dfree(dK); dfree(dL); dfree(dA); dfree(dOut);
    return 0;
}

}  // extern "C"
