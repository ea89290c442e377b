// K3n: latency-optimised triangular solves for FEW right-hand sides.
//
// The wide solve kernel (solve.cu) works on 64-column slabs; with 1..128 right-hand sides -- alpha = cK \ (y - m)
// in every GPE fit / update_alpha! (src/gp_utils.jl:15-18, one column) and the whitening solve L^-1 K_xb of
// predict_f at the border sentinels (src/cliff_face.jl:4-5, nb ~ 100 columns) -- it is a chain of T dependent
// row-block jobs and the time per link, not the flops, is what the caller waits for.  This kernel shortens the
// link:
//   * slabs are 8 columns wide (one DMMA n-tile), so nb = 100 sentinels become 13 independent chains on 13 x T
//     CTAs instead of 2 chains;
//   * the factor tiles of a row block are static data: thread 0 streams them through a TMA ring WITHOUT waiting
//     for any flag, so when X(kt) is published only its 8 KB have to be fetched (straight into B fragments);
//   * the diagonal tile is fetched into shared memory by one bulk copy at job start, long before it is needed;
//   * the 128-step substitution runs one warp per column with the column spread over the lanes (shuffle
//     broadcast of the pivot row), no block loop, no CTA barriers inside.
// Same dataflow protocol as solve.cu (atomic job counter in level order, per-(row block, slab) epoch flags).
#include "kernels.h"
#include "mainloop.cuh"
#include "tileops.cuh"

namespace geordd {

constexpr int NCOL = 8;                   // columns per narrow slab
constexpr int NST = 4;                    // ring stages, one 16-column chunk of a factor tile each
constexpr int NCHUNK_ELEMS = BK * TLD;    // 2112 doubles
constexpr int CH_PER_TILE = TILE / BK;    // 8

struct NarrowParams {
    const double* L;
    int T;
    double* B;
    long ldb;
    int nslabs;
    unsigned* flags;
    unsigned epoch;
    unsigned* counter;
    int* abort_word;
};

constexpr int NARROW_RING = NST * NCHUNK_ELEMS;
constexpr int NARROW_SMEM_ELEMS = NARROW_RING + (int)TSZ + NCOL * SC + TILE + (2 * NST + 1);
constexpr size_t NARROW_SMEM = (size_t)NARROW_SMEM_ELEMS * sizeof(double);

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) solve_narrow_kernel(NarrowParams p) {
    extern __shared__ __align__(16) double smem[];
    double* ring = smem;
    double* Ds = ring + NARROW_RING;          // diagonal tile image, D(m,k) at Ds[k*SC + m]
    double* Cs = Ds + TSZ;                    // C(m,n) at Cs[n*SC + m]
    double* invd = Cs + NCOL * SC;
    uint64_t* full = reinterpret_cast<uint64_t*>(invd + TILE);
    uint64_t* empty = full + NST;
    uint64_t* dbar = empty + NST;
    __shared__ int sh_job;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;

    unsigned zero;
    {
        unsigned hi;
        asm volatile("mov.u32 %0, %%clock_hi;" : "=r"(hi));
        zero = hi >> 31;   // run-time 0 that cannot be constant-folded (mainloop.cuh, stage release)
    }
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NTHREADS / 32);
        }
        mbar_init(dbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    unsigned it = 0;        // chunks pushed through the ring so far
    unsigned dphase = 0;    // parity of the next diagonal-tile arrival
    const int njobs = p.T * p.nslabs;

    while (true) {
        if (tid == 0) {
            int jb = (int)atomicAdd(p.counter, 1u);
            if (ld_volatile_int(p.abort_word) != 0) jb = njobs;
            sh_job = jb;
        }
        __syncthreads();
        const int job = sh_job;
        if (job >= njobs) break;
        const int lvl = job / p.nslabs, s = job % p.nslabs;
        const int rb = (MODE == SOLVE_FWD) ? lvl : (p.T - 1 - lvl);
        double* Bcols = p.B + (long)s * NCOL * p.ldb;
        const int ntiles = (MODE == SOLVE_FWD) ? rb : (p.T - 1 - rb);
        const int nchunks = ntiles * CH_PER_TILE;

        if (tid == 0) {   // diagonal tile: needed only after the whole K loop
            mbar_arrive_expect_tx(dbar, (unsigned)(TSZ * sizeof(double)));
            bulk_g2s(Ds, p.L + tile_off(rb, rb, p.T), (unsigned)(TSZ * sizeof(double)), dbar);
        }
        // chunk c of this job = columns [16*(c%8), +16) of factor tile (rb, kt(c/8))
        auto a_src = [&](int c) {
            int t = c / CH_PER_TILE;
            int kt = (MODE == SOLVE_FWD) ? t : (p.T - 1 - t);
            return p.L + tile_off(rb, kt, p.T) + (long)(c % CH_PER_TILE) * NCHUNK_ELEMS;
        };
        bool ok = true;
        unsigned it_load = it, it_use = it;
        it += (unsigned)nchunks;
        auto produce = [&](int c) -> bool {   // thread 0 only
            const int st = it_load % NST;
            if (!mbar_wait(empty + st, ((it_load / NST) & 1u) ^ 1u, p.abort_word)) return false;
            mbar_arrive_expect_tx(full + st, (unsigned)(NCHUNK_ELEMS * sizeof(double)));
            bulk_g2s(ring + st * NCHUNK_ELEMS, a_src(c), (unsigned)(NCHUNK_ELEMS * sizeof(double)), full + st);
            ++it_load;
            return true;
        };
        if (tid == 0) {
            const int npro = nchunks < NST - 1 ? nchunks : NST - 1;
            for (int c = 0; c < npro && ok; ++c) ok = produce(c);
        }

        // The diagonal tile lands while the first flags are still pending: pivot reciprocals and (backward) L' in
        // the strict upper triangle of the image are prepared before the K loop, off the critical path.
        if (!mbar_wait(dbar, dphase, p.abort_word)) ok = false;
        dphase ^= 1u;
        if (__syncthreads_or(!ok)) break;
        if (tid < TILE) invd[tid] = 1.0 / Ds[tid * SC + tid];
        if (MODE == SOLVE_BWD) {   // U(r, r2) = L(r2, r) for r2 > r, stored at image position (row r, column r2)
            for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
                int r = idx / TILE, r2 = idx % TILE;
                if (r2 > r) Ds[r2 * SC + r] = Ds[r * SC + r2];
            }
        }

        // accumulators: rows m0+g (two 8-row tiles per warp), columns 2tq, 2tq+1
        const int m0 = warp * 16;
        double acc[2][2];
        {
            const double* Brb = Bcols + (long)rb * TILE;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[i][e] = -ldcg(Brb + (long)(2 * tq + e) * p.ldb + m0 + i * 8 + g);
        }
        for (int t = 0; t < ntiles && ok; ++t) {
            const int kt = (MODE == SOLVE_FWD) ? t : (p.T - 1 - t);
            if (!warp_wait_flags(p.flags + (long)kt * p.nslabs + s, nullptr, p.epoch, p.abort_word)) { ok = false; break; }
            // B fragments of X(kt): b[ks] = X(kt*128 + 4ks + tq, column g)
            double b[TILE / 4];
            {
                const double* Xk = Bcols + (long)g * p.ldb + (long)kt * TILE + tq;
#pragma unroll
                for (int ks = 0; ks < TILE / 4; ++ks) b[ks] = ldcg(Xk + ks * 4);
            }
#pragma unroll
            for (int j = 0; j < CH_PER_TILE; ++j) {
                const int c = t * CH_PER_TILE + j;
                if (tid == 0 && c + NST - 1 < nchunks) ok = produce(c + NST - 1);
                ok = __shfl_sync(0xffffffffu, ok, 0) && ok;
                const int st = it_use % NST;
                if (ok && !mbar_wait(full + st, (it_use / NST) & 1u, p.abort_word)) ok = false;
                if (!ok) break;
                const double* As = ring + st * NCHUNK_ELEMS;
                unsigned token = 0;
#pragma unroll
                for (int kk = 0; kk < BK / 4; ++kk) {
                    const int k = kk * 4 + tq;
                    double a0 = As[k * TLD + m0 + g], a1 = As[k * TLD + m0 + 8 + g];
                    token |= (unsigned)__double2hiint(a0) | (unsigned)__double2hiint(a1);
                    dmma(acc[0][0], acc[0][1], a0, b[j * 4 + kk]);
                    dmma(acc[1][0], acc[1][1], a1, b[j * 4 + kk]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_after(empty + st, token & zero);
                ++it_use;
            }
        }
        if (__syncthreads_or(!ok)) break;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) Cs[(2 * tq + e) * SC + m0 + i * 8 + g] = -acc[i][e];
        __syncthreads();
        // substitution: warp w solves column w; lane l owns rows l, l+32, l+64, l+96
        {
            double* col = Cs + warp * SC;
            double x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = col[lane + 32 * q];
            if (MODE == SOLVE_FWD) {
#pragma unroll
                for (int q0 = 0; q0 < 4; ++q0) {
#pragma unroll 8
                    for (int rl = 0; rl < 32; ++rl) {
                        const int r = q0 * 32 + rl;
                        double xr = __shfl_sync(0xffffffffu, x[q0], rl) * invd[r];
                        if (lane == rl) x[q0] = xr;
                        const double* Lc = Ds + r * SC;   // column r of L: L(m, r)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q < q0) continue;
                            const int m = lane + 32 * q;
                            if (m > r) x[q] -= xr * Lc[m];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q0 = 3; q0 >= 0; --q0) {
#pragma unroll 8
                    for (int rl = 31; rl >= 0; --rl) {
                        const int r = q0 * 32 + rl;
                        double xr = __shfl_sync(0xffffffffu, x[q0], rl) * invd[r];
                        if (lane == rl) x[q0] = xr;
                        const double* Uc = Ds + r * SC;   // column r of U = L': U(m, r) = L(r, m), m < r
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q > q0) continue;
                            const int m = lane + 32 * q;
                            if (m < r) x[q] -= xr * Uc[m];
                        }
                    }
                }
            }
            double* Brb = Bcols + (long)warp * p.ldb + (long)rb * TILE;
#pragma unroll
            for (int q = 0; q < 4; ++q) Brb[lane + 32 * q] = x[q];
        }
        __threadfence();
        fence_proxy_async();   // this job's generic reads of Ds / ring before the next job's bulk copies
        __syncthreads();
        if (tid == 0) st_release(p.flags + (long)rb * p.nslabs + s, p.epoch);
    }
}

template <int MODE>
static void launch_narrow_mode(cudaStream_t st, Sched& sched, const NarrowParams& p) {
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(solve_narrow_kernel<MODE>, NARROW_SMEM, attr_done);
    int njobs = p.T * p.nslabs;
    int grid = sched.num_sms < njobs ? sched.num_sms : njobs;
    solve_narrow_kernel<MODE><<<grid, NTHREADS, NARROW_SMEM, st>>>(p); ++g_kernel_launches;
}

void launch_solve_narrow(cudaStream_t st, Sched& sched, SolveMode mode, const double* L, int npad, double* B, long ldb,
                         int live_cols, int lane) {
    NarrowParams p;
    p.L = L; p.T = npad / TILE; p.B = B; p.ldb = ldb; p.nslabs = (live_cols + NCOL - 1) / NCOL;
    // lane > 0: own job counter and flag region, so that solves of independent GPs can run side by side on other streams
    p.flags = sched.flags + (size_t)lane * NARROW_LANE_FLAGS; p.counter = sched.counter + lane; p.abort_word = sched.abort_word;
    cudaMemsetAsync(p.counter, 0, sizeof(unsigned), st);
    sched.epoch++;
    p.epoch = sched.epoch;
    if (mode == SOLVE_FWD) launch_narrow_mode<SOLVE_FWD>(st, sched, p);
    else launch_narrow_mode<SOLVE_BWD>(st, sched, p);
}

}  // namespace geordd
