// Device-resident micro-benchmarks exposed through the C ABI (used by bench.py / profiles):
//   geordd_bench_dmma_peak : register-resident mma.sync.m8n8k4.f64 loop = the FP64 tensor-pipe
//                            ceiling on this part (MEASURED_PEAKS.json holds no FP64 figure).
//   geordd_bench_potrf     : the dataflow Cholesky on a synthetic SPD covariance built on device.
#include "api_internal.h"
#include "common.cuh"
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

}  // extern "C"
