// CTA-level FP64 GEMM mainloop on the DMMA pipe.
//
//   acc(128 x BN) += sum over K-chunks  A_chunk(128 x 16) * B_chunk(16 x BN)
//
// 256 threads = 8 warps feed and drain a STAGES-deep shared-memory ring without any CTA-wide barrier in the
// K loop: a stage's arrival is tracked by an mbarrier (`full`), its release by a second one (`empty`) on which
// each warp arrives after its last fragment read, so warps drift by up to STAGES-2 chunks.
//
// How an operand chunk travels global -> shared is a per-operand template choice:
//   ROWS_TMA  one bulk async copy (TMA unit, SASS UBLKCP) per contiguous row, landing in padded rows
//   BLK_TMA   ONE bulk async copy per chunk: the global storage already is the padded shared-memory image
//             (tile-blocked padded factor storage, kernels.h)
//   ROWS_GEN / BLK_GEN   the same two shapes moved with 16-byte cp.async (LDGSTS) by all threads, completion
//             signalled on the same mbarrier (cp.async.mbarrier.arrive.noinc)
// Releasing a stage that the TMA unit refills (any *_TMA mode): the consumer's arrive on `empty` must not issue
// while the warp's last fragment loads are still in flight.  ptxas sinks the DMMAs that consume those loads below
// the arrive and SYNCS.ARRIVE does not wait on LDS scoreboards; the refill then comes through the async proxy,
// which is not ordered behind generic-proxy loads still in the LSU pipe, and can overwrite data not yet read
// (seen as rare wrong accumulator strips in the flagged solves, where long flag waits keep the ring full; 20 000
// right-hand sides against SciPy, scripts/gpu_stress_solve.py).  Fix: the arrive's address carries a run-time-zero
// data dependency on every fragment register of the stage (mbar_arrive_after), so it issues only after all loads
// of the stage have returned.  A fence.proxy.async before the arrive is the textbook alternative and is also
// exact, but measured slower.  Refills through the generic proxy (*_GEN) stay ordered behind the loads in the
// same pipe and need neither.
// The flagged solves and the Cholesky keep the *_GEN modes: with the dependency in place the bulk-copy variants
// are exact there too but not faster (row copies of the in-place X are the limiter); the flag-free triangular
// multiply and the plain GEMM use the TMA modes.
//
// Fragments are read with conflict-free LDS.64 (row strides == 4 mod 16 doubles) and multiplied with
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), 16-32 independent accumulators per warp.
//
// Layout flags say which index is contiguous in a chunk:
//   A_MC: A(m,k) at a[k*lda + m]  (column-major tile)      else  A(m,k) at a[m*lda + k]
//   B_NC: B(k,n) at b[k*ldb + n]                            else  B(k,n) at b[n*ldb + k]
// The chunk source `Src` supplies per-chunk base pointers (so a K loop can walk over many tiles of a big matrix)
// and a warp-level `ready(chunk)` hook where dataflow kernels wait for producer tiles.
#pragma once
#include <type_traits>
#include <utility>
#include "common.cuh"

namespace geordd {

enum CopyMode { ROWS_TMA = 0, BLK_TMA = 1, ROWS_GEN = 2, BLK_GEN = 3 };
__host__ __device__ constexpr bool copy_is_gen(int m) { return m == ROWS_GEN || m == BLK_GEN; }
__host__ __device__ constexpr bool copy_is_blk(int m) { return m == BLK_TMA || m == BLK_GEN; }

// Ring bookkeeping that persists across the jobs of a persistent kernel.
struct Ring {
    uint64_t* full;    // [STAGES] data landed
    uint64_t* empty;   // [STAGES] stage free (8 arrivals, one per warp)
    unsigned it;       // chunks pushed through the ring so far (same value in every thread)
    unsigned zero;     // 0 at run time, opaque to nvcc and ptxas (mbar_arrive_after)
    int* abort_word;
};

// Does the operand source offer the non-blocking try_ready(chunk)?  (see Mainloop::run)
template <class S>
struct src_can_defer {
    template <class U> static auto probe(int) -> decltype(std::declval<U&>().try_ready(0), std::true_type{});
    template <class> static std::false_type probe(...);
    static constexpr bool value = decltype(probe<S>(0))::value;
};

template <int BN_, bool A_MC_, bool B_NC_, int STAGES_, int A_MODE_ = ROWS_TMA, int B_MODE_ = ROWS_TMA>
struct Mainloop {
    static constexpr int A_MODE = A_MODE_, B_MODE = B_MODE_;
    static_assert(!copy_is_blk(A_MODE_) || A_MC_, "block copies of A need the m-contiguous layout");
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
    static constexpr unsigned bytes_of(int mode, int elems, int rows, int run) {
        return copy_is_gen(mode) ? 0u : (unsigned)((copy_is_blk(mode) ? elems : rows * run) * sizeof(double));
    }
    // bytes moved by the TMA unit per stage, announced by thread 0 together with its arrival
    static constexpr unsigned STAGE_TX_BYTES =
        bytes_of(A_MODE_, A_ELEMS, A_ROWS, A_RUN) + bytes_of(B_MODE_, B_ELEMS, B_ROWS, B_RUN);
    static constexpr bool ANY_GEN = copy_is_gen(A_MODE_) || copy_is_gen(B_MODE_);
    static constexpr bool ANY_TMA = !copy_is_gen(A_MODE_) || !copy_is_gen(B_MODE_);
    static constexpr bool ANY_ROWS_TMA = A_MODE_ == ROWS_TMA || B_MODE_ == ROWS_TMA;
    // only thread 0 has anything to do in produce(): warps 1..7 can skip it entirely
    static constexpr bool SOLO = !ANY_GEN && !ANY_ROWS_TMA;
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
        {
            unsigned hi;
            asm volatile("mov.u32 %0, %%clock_hi;" : "=r"(hi));
            ring.zero = hi >> 31;   // stays 0 for centuries of uptime; cannot be constant-folded
        }
        ring.abort_word = abort_word;
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; ++s) {
                // thread 0's arrival (+ TMA bytes), plus one pre-counted cp.async arrival per thread in GEN modes
                mbar_init(ring.full + s, 1 + (ANY_GEN ? NTHREADS : 0));
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

    // ---- copy engines -------------------------------------------------------------------------------------
    // TMA, one copy per row: this warp's share is rows [warp*ROWS/8, (warp+1)*ROWS/8), one copy per lane.
    template <int ROWS, int RUN, int STRIDE>
    __device__ static __forceinline__ void rows_tma(double* s, const double* g, long ld, uint64_t* bar) {
        constexpr int PER_WARP = ROWS / 8;
        static_assert(ROWS % 8 == 0 && PER_WARP <= 32, "row split");
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane < PER_WARP) {
            const int r = warp * PER_WARP + lane;
            bulk_g2s(s + r * STRIDE, g + (long)r * ld, RUN * (unsigned)sizeof(double), bar);
        }
    }
    // generic proxy, padded rows
    template <int ROWS, int RUN, int STRIDE>
    __device__ static __forceinline__ void rows_gen(double* s, const double* g, long ld) {
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
    // generic proxy, contiguous padded image of ELEMS doubles
    template <int ELEMS>
    __device__ static __forceinline__ void blk_gen(double* s, const double* g) {
        constexpr int TOTAL = ELEMS / 2;
#pragma unroll
        for (int t = 0; t < (TOTAL + NTHREADS - 1) / NTHREADS; ++t) {
            int id = threadIdx.x + t * NTHREADS;
            if (TOTAL % NTHREADS == 0 || id < TOTAL) cp_async16(s + 2 * id, g + 2 * id);
        }
    }

    // TRY = false: waits until the chunk's operands are published; false = the wait was aborted.
    // TRY = true (sources with try_ready): one poll; PRODUCE_NOT_READY leaves the ring untouched -- the caller goes on
    // consuming what it has and asks again.
    static constexpr int PRODUCE_ABORT = 0, PRODUCE_OK = 1, PRODUCE_NOT_READY = 2;
    template <bool TRY, class Src>
    __device__ static __forceinline__ int produce(double* smem, Ring& ring, unsigned it_load, int chunk, Src& src) {
        if (SOLO && threadIdx.x >= 32) return PRODUCE_OK;   // warp 0 alone feeds the ring
        const int s = it_load % STAGES;
        const unsigned ph = (it_load / STAGES) & 1u;
        if constexpr (TRY) {
            if (!src.try_ready(chunk)) return PRODUCE_NOT_READY;
        }
        if (!mbar_wait(ring.empty + s, ph ^ 1u, ring.abort_word)) return PRODUCE_ABORT;   // all warps done with this stage
        if constexpr (!TRY) {
            if (!src.ready(chunk)) return PRODUCE_ABORT;                                   // dataflow: operand tiles published
        }
        double* As = smem + s * STAGE_ELEMS;
        double* Bs = As + A_ELEMS;
        uint64_t* bar = ring.full + s;
        if (threadIdx.x == 0) {
            if (STAGE_TX_BYTES) mbar_arrive_expect_tx(bar, STAGE_TX_BYTES);
            else mbar_arrive(bar);
            if (A_MODE == BLK_TMA) bulk_g2s(As, src.a_ptr(chunk), (unsigned)(A_ELEMS * sizeof(double)), bar);
            if (B_MODE == BLK_TMA) bulk_g2s(Bs, src.b_ptr(chunk), (unsigned)(B_ELEMS * sizeof(double)), bar);
        }
        if (A_MODE == ROWS_TMA) rows_tma<A_ROWS, A_RUN, SA>(As, src.a_ptr(chunk), src.lda, bar);
        if (B_MODE == ROWS_TMA) rows_tma<B_ROWS, B_RUN, SB>(Bs, src.b_ptr(chunk), src.ldb, bar);
        if (A_MODE == ROWS_GEN) rows_gen<A_ROWS, A_RUN, SA>(As, src.a_ptr(chunk), src.lda);
        if (B_MODE == ROWS_GEN) rows_gen<B_ROWS, B_RUN, SB>(Bs, src.b_ptr(chunk), src.ldb);
        if (A_MODE == BLK_GEN) blk_gen<A_ELEMS>(As, src.a_ptr(chunk));
        if (B_MODE == BLK_GEN) blk_gen<B_ELEMS>(Bs, src.b_ptr(chunk));
        if (ANY_GEN) cp_async_mbar_arrive_noinc(bar);   // fires when this thread's cp.async of the stage have landed
        return PRODUCE_OK;
    }

    // Returns the OR of the high words of every fragment loaded from the stage (the release dependency; only
    // materialised when the TMA unit refills the ring).
    __device__ static __forceinline__ unsigned compute(const double* smem, int s, Acc& acc) {
        unsigned token = 0;
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
            if (ANY_TMA) {
#pragma unroll
                for (int i = 0; i < MI; ++i) token |= (unsigned)__double2hiint(a[i]);
#pragma unroll
                for (int j = 0; j < NJ; ++j) token |= (unsigned)__double2hiint(b[j]);
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
        }
        return token;
    }

    // Runs the whole K loop of one job.  Returns false (for this warp) if a wait was aborted; callers combine
    // the result across the CTA before touching shared memory again.  On a true return every chunk issued
    // has been consumed by this warp.
    template <class Src>
    __device__ static bool run(double* smem, Ring& ring, int nchunks, Src& src, Acc& acc) {
        // Sources with a non-blocking try_ready (the dataflow Cholesky): a chunk whose operands are not published yet is
        // DEFERRED -- the warp keeps consuming the chunks already in the ring and asks again after each, and blocks only
        // when it has nothing left to compute.  Without this a consumer sat in the wait with STAGES-2 loaded chunks
        // unconsumed, and their compute time landed on the critical path once the flag arrived (round 2 trace: 4.5 of the
        // 46 us per tile column of the factorisation).
        constexpr bool DEFER = src_can_defer<Src>::value;
        const unsigned it0 = ring.it;
        unsigned it_use = ring.it;
        ring.it += (unsigned)nchunks;
        const int lane = threadIdx.x & 31;
        int nl = 0;   // chunks this warp has issued
        const int npro = nchunks < STAGES - 1 ? nchunks : STAGES - 1;
        while (nl < npro) {
            const int r = (DEFER && nl > 0) ? produce<DEFER>(smem, ring, it0 + nl, nl, src) : produce<false>(smem, ring, it0 + nl, nl, src);
            if (r == PRODUCE_ABORT) return false;
            if (r == PRODUCE_NOT_READY) break;
            ++nl;
        }
        for (int kc = 0; kc < nchunks; ++kc) {
            if (DEFER && nl <= kc) {   // nothing left to compute: wait for this chunk's operands
                if (produce<false>(smem, ring, it0 + nl, nl, src) == PRODUCE_ABORT) return false;
                ++nl;
            }
            const int s = it_use % STAGES;
            if (!mbar_wait(ring.full + s, (it_use / STAGES) & 1u, ring.abort_word)) return false;
            const unsigned token = compute(smem, s, acc);
            __syncwarp();
#ifdef GEORDD_RELEASE_FENCE
            // Textbook ordering (documented PTX): a proxy fence between this warp's generic-proxy fragment loads and the
            // async-proxy refill that its arrival permits.  1-4 % slower than the data dependency below (round 1), kept as a
            // build option (-DGEORDD_RELEASE_FENCE) for toolkits whose scheduling the dependency trick has not been checked on.
            if (ANY_TMA) fence_proxy_async();
            if (lane == 0) mbar_arrive(ring.empty + s);
#else
            if (lane == 0) {
                if (ANY_TMA) mbar_arrive_after(ring.empty + s, token & ring.zero);
                else mbar_arrive(ring.empty + s);
            }
#endif
            ++it_use;
            // Refill AFTER computing: the stage to refill held chunk kc-1, which the other warps have had a whole
            // chunk time to release -- refilling before the compute made every producer wait for the slowest warp
            // once per chunk (7 % of all warp samples sat in that wait, profiles/r01d).
            if constexpr (DEFER) {
                while (nl < nchunks && nl <= kc + STAGES - 1) {   // catches up after a deferral
                    const int r = produce<true>(smem, ring, it0 + nl, nl, src);
                    if (r == PRODUCE_ABORT) return false;
                    if (r == PRODUCE_NOT_READY) break;
                    ++nl;
                }
            } else {
                const int nxt = kc + STAGES - 1;
                if (nxt < nchunks) {
                    if (produce<false>(smem, ring, it0 + nxt, nxt, src) == PRODUCE_ABORT) return false;
                }
            }
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

// Warp-level wait for up to two dataflow flags: lane 0 polls with acquire loads, the result is broadcast, and
// __syncwarp orders the other lanes' subsequent (generic-proxy) loads of the published tile behind the acquire.
__device__ __forceinline__ bool warp_wait_flags(const unsigned* f0, const unsigned* f1, unsigned epoch, int* abort_word) {
    int ok = 1;
    if ((threadIdx.x & 31) == 0) {
        ok = wait_flag(f0, epoch, abort_word);
        if (ok && f1) ok = wait_flag(f1, epoch, abort_word);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    if (ok) {   // every lane performs its own acquire of the (already set) flags before it touches the tile
        unsigned v = ld_acquire(f0);
        if (f1) v &= ld_acquire(f1);
        if (v != epoch) ok = 0;
    }
    return __all_sync(0xffffffffu, ok) != 0;
}

}  // namespace geordd
