// CTA-level FP64 GEMM mainloop on the DMMA pipe.
//
//   acc(128 x BN) += sum over K-chunks  A_chunk(128 x 16) * B_chunk(16 x BN)
//
// 256 threads = 8 warps.  Operand chunks are staged global -> shared with 16-byte cp.async in a
// STAGES-deep ring (one __syncthreads per chunk), fragments are read with conflict-free LDS.64
// (row strides == 4 mod 16 doubles) and multiplied with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
//
// Layout flags describe which index is contiguous in GLOBAL memory for a chunk:
//   A_MC: A(m,k) at a[k*lda + m]  (column-major tile)      else  A(m,k) at a[m*lda + k]
//   B_NC: B(k,n) at b[k*ldb + n]                            else  B(k,n) at b[n*ldb + k]
// The chunk source `Src` supplies per-chunk base pointers (so a K loop can walk over many tiles of
// a big matrix) and a `ready(chunk)` hook where dataflow kernels wait for producer tiles.
#pragma once
#include "common.cuh"

namespace geordd {

template <int BN_, bool A_MC_, bool B_NC_, int STAGES_>
struct Mainloop {
    static constexpr int BM = TILE;
    static constexpr int BN = BN_;
    static constexpr bool A_MC = A_MC_;
    static constexpr bool B_NC = B_NC_;
    static constexpr int STAGES = STAGES_;
    static constexpr int WARPS_N = (BN >= 128) ? 4 : 2;
    static constexpr int WARPS_M = 8 / WARPS_N;
    static constexpr int WM = BM / WARPS_M;
    static constexpr int WN = BN / WARPS_N;
    static constexpr int MI = WM / 8;
    static constexpr int NJ = WN / 8;

    static constexpr int A_RUN = A_MC ? BM : BK;    // contiguous run length of one smem row
    static constexpr int A_ROWS = A_MC ? BK : BM;
    static constexpr int SA = A_RUN + PAD;
    static constexpr int A_ELEMS = A_ROWS * SA;
    static constexpr int B_RUN = B_NC ? BN : BK;
    static constexpr int B_ROWS = B_NC ? BK : BN;
    static constexpr int SB = B_RUN + PAD;
    static constexpr int B_ELEMS = B_ROWS * SB;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);

    struct Acc {
        double v[MI][NJ][2];
    };

    __device__ static __forceinline__ void zero(Acc& acc) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
    }

    // Row / column owned by accumulator element (i, j, e) of this thread.
    __device__ static __forceinline__ int acc_row(int i) {
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        return (warp / WARPS_N) * WM + i * 8 + (lane >> 2);
    }
    __device__ static __forceinline__ int acc_col(int j, int e) {
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        return (warp % WARPS_N) * WN + j * 8 + 2 * (lane & 3) + e;
    }

    template <int ROWS, int RUN, int STRIDE>
    __device__ static __forceinline__ void issue_tile(double* s, const double* g, long ld) {
        constexpr int CH_PER_ROW = RUN / 2;
        constexpr int TOTAL = ROWS * CH_PER_ROW;
        static_assert(TOTAL % NTHREADS == 0, "tile copy must divide evenly");
#pragma unroll
        for (int t = 0; t < TOTAL / NTHREADS; ++t) {
            int id = threadIdx.x + t * NTHREADS;
            int r = id / CH_PER_ROW, c = (id % CH_PER_ROW) * 2;
            cp_async16(s + r * STRIDE + c, g + (long)r * ld + c);
        }
    }

    template <class Src>
    __device__ static __forceinline__ void issue(double* smem, int chunk, const Src& src) {
        double* As = smem + (chunk % STAGES) * STAGE_ELEMS;
        double* Bs = As + A_ELEMS;
        issue_tile<A_ROWS, A_RUN, SA>(As, src.a_ptr(chunk), src.lda);
        issue_tile<B_ROWS, B_RUN, SB>(Bs, src.b_ptr(chunk), src.ldb);
    }

    __device__ static __forceinline__ void compute(const double* smem, int chunk, Acc& acc) {
        const double* As = smem + (chunk % STAGES) * STAGE_ELEMS;
        const double* Bs = As + A_ELEMS;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int g = lane >> 2, tq = lane & 3;
        const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int k = kk * 4 + tq;
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                int m = wm0 + i * 8 + g;
                a[i] = A_MC ? As[k * SA + m] : As[m * SA + k];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int n = wn0 + j * 8 + g;
                b[j] = B_NC ? Bs[k * SB + n] : Bs[n * SB + k];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
        }
    }

    // Runs the whole K loop.  Returns false if a dataflow wait was aborted (watchdog / error).
    // On return all cp.async groups are drained and a __syncthreads() has been passed, so `smem`
    // may be reused by the caller.
    template <class Src>
    __device__ static bool run(double* smem, int nchunks, Src& src, Acc& acc) {
        bool ok = true;
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nchunks) {
                ok = ok && src.ready(s);
                if (ok) issue(smem, s, src);
            }
            cp_async_commit();
        }
        for (int kc = 0; kc < nchunks; ++kc) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            int nxt = kc + STAGES - 1;
            if (nxt < nchunks && ok) {
                ok = src.ready(nxt);
                if (ok) issue(smem, nxt, src);
            }
            cp_async_commit();
            if (!ok) break;
            compute(smem, kc, acc);
        }
        cp_async_wait<0>();
        __syncthreads();
        return ok;
    }
};

// Plain strided source: K walks contiguously through A and B (no tiles, no flags).
struct PlainSrc {
    const double* a;
    const double* b;
    long lda, ldb;
    long a_step, b_step;  // pointer advance per chunk
    __device__ __forceinline__ const double* a_ptr(int c) const { return a + c * a_step; }
    __device__ __forceinline__ const double* b_ptr(int c) const { return b + c * b_step; }
    __device__ __forceinline__ bool ready(int) const { return true; }
};

}  // namespace geordd
