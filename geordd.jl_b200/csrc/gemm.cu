// Generic tiled FP64 GEMM on the DMMA mainloop (K5/K6 building block):
//   mu_b = cK_T * alpha_T - cK_C * alpha_C        (src/gp_utils.jl:1-5, src/hypotest_chi2.jl:13-15)
//   Sigma = K_bb - V'V                            (predict_f full_cov, src/cliff_face.jl:4-5)
//   D'D, X_invA_Xt, the W*K*W' products of pval_invvar_calib (src/hypotest_invvar.jl:88-96)
// Split-K is deterministic: partial tiles go to a workspace and are summed in a fixed order.
#include "kernels.h"
#include "mainloop.cuh"

namespace geordd {

template <bool A_MC, bool B_NC, bool B_XB = false>
__global__ void __launch_bounds__(NTHREADS, 2)
gemm_kernel(int K, double alpha, const double* __restrict__ A, long lda, const double* __restrict__ B, long ldb,
            double beta, double* __restrict__ C, long ldc, int splitk, double* __restrict__ ws, long ws_stride) {
    using ML = Mainloop<64, A_MC, B_NC, 3, ROWS_TMA, B_XB ? BLK_TMA : ROWS_TMA>;
    static_assert(!B_XB || !B_NC, "ZB storage is the k-contiguous B stage");
    extern __shared__ __align__(16) double smem[];
    __shared__ int sh_abort;   // never set: plain GEMM has no dataflow waits, only the ring watchdog
    if (threadIdx.x == 0) sh_abort = 0;
    Ring ring;
    ML::ring_init(ring, reinterpret_cast<uint64_t*>(smem + ML::RING_ELEMS), &sh_abort);
    const int mt = blockIdx.x, nt = blockIdx.y, z = blockIdx.z;
    const int kchunks_total = K / BK;
    const int per = (kchunks_total + splitk - 1) / splitk;
    const int c0 = z * per;
    int nch = kchunks_total - c0;
    if (nch > per) nch = per;
    if (nch < 0) nch = 0;
    const long k0 = (long)c0 * BK;
    PlainSrc src;
    src.lda = lda;
    src.ldb = ldb;
    src.a = A_MC ? A + k0 * lda + (long)mt * TILE : A + (long)mt * TILE * lda + k0;
    src.b = B_XB ? B + ((long)nt * kchunks_total + c0) * ZB_ELEMS
                 : (B_NC ? B + k0 * ldb + (long)nt * 64 : B + (long)nt * 64 * ldb + k0);
    src.a_step = A_MC ? (long)BK * lda : BK;
    src.b_step = B_XB ? (long)ZB_ELEMS : (B_NC ? (long)BK * ldb : BK);
    typename ML::Acc acc;
    ML::zero(acc);
    ML::run(smem, ring, nch, src, acc);
#pragma unroll
    for (int a = 0; a < ML::MI; ++a)
#pragma unroll
        for (int b = 0; b < ML::NJ; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                long m = (long)mt * TILE + ML::acc_row(a), n = (long)nt * 64 + ML::acc_col(b, e);
                if (splitk > 1) {
                    ws[(long)z * ws_stride + n * ldc + m] = acc.v[a][b][e];
                } else {
                    double* p = C + n * ldc + m;
                    double v = alpha * acc.v[a][b][e];
                    if (beta != 0.0) v += beta * *p;
                    *p = v;
                }
            }
}

__global__ void splitk_reduce_kernel(int M, int N, double alpha, double beta, double* C, long ldc, int splitk,
                                     const double* ws, long ws_stride) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * N) return;
    long m = idx % M, n = idx / M;
    double s = 0.0;
    for (int z = 0; z < splitk; ++z) s += ws[(long)z * ws_stride + n * ldc + m];
    double v = alpha * s;
    double* p = C + n * ldc + m;
    if (beta != 0.0) v += beta * *p;
    *p = v;
}

template <bool A_MC, bool B_NC, bool B_XB = false>
static void launch_t(cudaStream_t st, int M, int N, int K, double alpha, const double* A, long lda, const double* B,
                     long ldb, double beta, double* C, long ldc, int splitk, double* ws) {
    using ML = Mainloop<64, A_MC, B_NC, 3, ROWS_TMA, B_XB ? BLK_TMA : ROWS_TMA>;
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(gemm_kernel<A_MC, B_NC, B_XB>, ML::SMEM_BYTES + ML::BAR_WORDS * 8, attr_done);
    dim3 grid(M / TILE, N / 64, splitk);
    long ws_stride = (long)ldc * N;
    gemm_kernel<A_MC, B_NC, B_XB><<<grid, NTHREADS, ML::SMEM_BYTES + ML::BAR_WORDS * 8, st>>>(K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws,
                                                                   ws_stride); ++g_kernel_launches;
    if (splitk > 1) {
        long tot = (long)M * N;
        splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(M, N, alpha, beta, C, ldc, splitk, ws,
                                                                          ws_stride); ++g_kernel_launches;
    }
}

void launch_gemm(cudaStream_t st, bool a_mc, bool b_nc, int M, int N, int K, double alpha, const double* A, long lda,
                 const double* B, long ldb, double beta, double* C, long ldc, int splitk, double* ws, bool b_xb) {
    if (splitk < 1) splitk = 1;
    if (b_xb) launch_t<true, false, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws);
    else if (a_mc && b_nc) launch_t<true, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws);
    else if (a_mc && !b_nc) launch_t<true, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws);
    else if (!a_mc && b_nc) launch_t<false, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws);
    else launch_t<false, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, splitk, ws);
}

}  // namespace geordd
