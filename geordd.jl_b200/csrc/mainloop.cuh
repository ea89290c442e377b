// CTA-level FP64 GEMM mainloop on the DMMA pipe.
//
//   acc(128 x BN) += sum over K-chunks  A_chunk(128 x 16) * B_chunk(16 x BN)
//
// 256 threads = 8 warps, each both producer and consumer of a STAGES-deep shared-memory ring:
//   * operand rows are moved global -> shared by the TMA unit with bulk async copies (SASS UBLKCP), one
//     contiguous row per copy, landing in padded rows; a stage's arrival is tracked by an mbarrier
//     (expect_tx / complete_tx), its release by a second mbarrier on which each warp arrives after its last
//     fragment read.  There is NO CTA-wide barrier in the K loop: warps drift by up to STAGES-2 chunks.
//   * fragments are read with conflict-free LDS.64 (row strides == 4 mod 16 doubles) and multiplied with
//     mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), 16-32 independent accumulators per warp.
//
// Layout flags describe which index is contiguous in GLOBAL memory for a chunk:
//   A_MC: A(m,k) at a[k*lda + m]  (column-major tile)      else  A(m,k) at a[m*lda + k]
//   B_NC: B(k,n) at b[k*ldb + n]                            else  B(k,n) at b[n*ldb + k]
// The chunk source `Src` supplies per-chunk base pointers (so a K loop can walk over many tiles of a big
// matrix) and a warp-level `ready(chunk)` hook where dataflow kernels wait for producer tiles.
#pragma once
#include "common.cuh"

namespace geordd {

// Ring bookkeeping that persists across the jobs of a persistent kernel.
struct Ring {
    uint64_t* full;    // [STAGES] data landed   (1 arrival + tx bytes)
    uint64_t* empty;   // [STAGES] stage free     (8 arrivals, one per warp)
    unsigned it;       // chunks pushed through the ring so far (same value in every thread)
    int* abort_word;
};

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
    static constexpr unsigned STAGE_TX_BYTES = (unsigned)((A_ROWS * A_RUN + B_ROWS * B_RUN) * sizeof(double));
    static constexpr int RING_ELEMS = STAGES * STAGE_ELEMS;
    static constexpr size_t SMEM_BYTES = (size_t)RING_ELEMS * sizeof(double);
    static constexpr int BAR_WORDS = 2 * STAGES;    // uint64 barriers needed (full + empty)

    struct Acc {
        double v[MI][NJ][2];
    };

    // Call once per kernel (all threads), before the first run().  `bars` points to BAR_WORDS uint64 in smem.
    __device__ static __forceinline__ void ring_init(Ring& ring, uint64_t* bars, int* abort_word) {
        ring.full = bars;
        ring.empty = bars + STAGES;
        ring.it = 0;
        ring.abort_word = abort_word;
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(ring.full + s, 1);
                mbar_init(ring.empty + s, NTHREADS / 32);
            }
            fence_mbar_init();
        }
        __syncthreads();
    }

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

    // This warp's share of one operand tile: rows [warp*ROWS/8, (warp+1)*ROWS/8), one bulk copy per lane.
    template <int ROWS, int RUN, int STRIDE>
    __device__ static __forceinline__ void issue_rows(double* s, const double* g, long ld, uint64_t* bar) {
        constexpr int PER_WARP = ROWS / 8;
        static_assert(ROWS % 8 == 0 && PER_WARP <= 32, "row split");
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane < PER_WARP) {
            const int r = warp * PER_WARP + lane;
            bulk_g2s(s + r * STRIDE, g + (long)r * ld, RUN * (unsigned)sizeof(double), bar);
        }
    }

    template <class Src>
    __device__ static __forceinline__ bool produce(double* smem, Ring& ring, unsigned it_load, int chunk, Src& src) {
        const int s = it_load % STAGES;
        const unsigned ph = (it_load / STAGES) & 1u;
        if (!mbar_wait(ring.empty + s, ph ^ 1u, ring.abort_word)) return false;   // all warps done with this stage
        if (!src.ready(chunk)) return false;                                       // dataflow: operand tiles published
        double* As = smem + s * STAGE_ELEMS;
        double* Bs = As + A_ELEMS;
        if (threadIdx.x == 0) mbar_arrive_expect_tx(ring.full + s, STAGE_TX_BYTES);
        issue_rows<A_ROWS, A_RUN, SA>(As, src.a_ptr(chunk), src.lda, ring.full + s);
        issue_rows<B_ROWS, B_RUN, SB>(Bs, src.b_ptr(chunk), src.ldb, ring.full + s);
        return true;
    }

    __device__ static __forceinline__ void compute(const double* smem, int s, Acc& acc) {
        const double* As = smem + s * STAGE_ELEMS;
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

    // Runs the whole K loop of one job.  Returns false (for this warp) if a wait was aborted; callers combine
    // the result across the CTA before touching shared memory again.  On a true return every chunk issued
    // has been consumed by this warp.
    template <class Src>
    __device__ static bool run(double* smem, Ring& ring, int nchunks, Src& src, Acc& acc) {
        unsigned it_load = ring.it, it_use = ring.it;
        ring.it += (unsigned)nchunks;
        const int lane = threadIdx.x & 31;
        const int npro = nchunks < STAGES - 1 ? nchunks : STAGES - 1;
        for (int c = 0; c < npro; ++c) {
            if (!produce(smem, ring, it_load, c, src)) return false;
            ++it_load;
        }
        for (int kc = 0; kc < nchunks; ++kc) {
            const int nxt = kc + STAGES - 1;
            if (nxt < nchunks) {
                if (!produce(smem, ring, it_load, nxt, src)) return false;
                ++it_load;
            }
            const int s = it_use % STAGES;
            if (!mbar_wait(ring.full + s, (it_use / STAGES) & 1u, ring.abort_word)) return false;
            compute(smem, s, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(ring.empty + s);
            ++it_use;
        }
        return true;
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

// Warp-level wait for up to two dataflow flags (lane 0 polls, result broadcast), then order the async proxy
// behind the acquire so the bulk copies that follow see the published tiles.
__device__ __forceinline__ bool warp_wait_flags(const unsigned* f0, const unsigned* f1, unsigned epoch, int* abort_word) {
    int ok = 1;
    if ((threadIdx.x & 31) == 0) {
        ok = wait_flag(f0, epoch, abort_word);
        if (ok && f1) ok = wait_flag(f1, epoch, abort_word);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    fence_proxy_async();
    return ok != 0;
}

}  // namespace geordd
