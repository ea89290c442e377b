// In-shared-memory operations on one 128 x N tile (the "finalize" step of every dataflow job):
// Cholesky of a diagonal tile, right/left triangular solves against a diagonal tile.
//
// Tile layout in smem: C(m, n) at Cs[n * SC + m], SC = 132 (column-major, padded).
// Every op is blocked by NB = 32: a thread-per-row (or per-column) substitution on the 32-wide
// diagonal block (true substitution, reciprocal of the pivot like OpenBLAS/LAPACK dpotf2), then a
// DMMA rank-32 update of the remainder straight out of shared memory.
#pragma once
#include "common.cuh"

namespace geordd {

constexpr int SC = TILE + PAD;   // 132: stride of tiles held in smem
constexpr int NB = 32;           // inner block
constexpr int LRS = NB + PAD;    // 36: stride of a row-panel held k-contiguous

// 1/sqrt(x) for normal positive x: MUFU.RSQ64H seed (relative error e0 ~ 2^-22) + ONE third-order step
//   y1 = y0 + y0*e*(1/2 + 3/8 e),  e = 1 - x*y0^2        (error ~ e0^3, below the rounding of the final FMA)
// Four dependent FP64 operations instead of the six of two Newton steps: this sits on the pivot chain of the
// Cholesky.  No slow path (the CUDA rsqrt() carries a call for subnormals that costs registers and a stack frame).
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}

// x[c2] -= xc * row[c2] for c2 in (c, NB): `row` points at 32 consecutive doubles in shared memory, 16-byte aligned;
// pairs starting at an even c2 are fetched with one LDS.128 (the substitution loops are bound by broadcast loads).
template <int C>
__device__ __forceinline__ void axpy_tail(double (&x)[NB], double xc, const double* row) {
    constexpr int first = C + 1;
    if constexpr ((first & 1) != 0 && first < NB) x[first] -= xc * row[first];
    constexpr int p0 = (first + 1) & ~1;   // first even index >= first
#pragma unroll
    for (int c2 = p0; c2 + 1 < NB; c2 += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + c2);
        x[c2] -= xc * v.x;
        x[c2 + 1] -= xc * v.y;
    }
}
template <int C = 0>
__device__ __forceinline__ void forward_subst32(double (&x)[NB], const double* invd, const double* L00, int stride) {
    // x <- x solved against the 32x32 lower block whose column c starts at L00 + c*stride (rows c.. contiguous)
    if constexpr (C < NB) {
        x[C] *= invd[C];
        axpy_tail<C>(x, x[C], L00 + C * stride);
        forward_subst32<C + 1>(x, invd, L00, stride);
    }
}

// C(m0.., n0..) -= A * B over K = NB, for 8x8 output tiles (mt, nt), mt in [0,mtiles), nt in
// [0,ntiles); fa(m,k), fb(k,n) fetch operands (m, n relative to m0, n0).  If `lower`, only tiles
// with m0+8mt >= n0+8nt are touched.  All 8 warps participate; caller syncs before/after.
template <class FA, class FB>
__device__ __forceinline__ void tile_update(double* Cs, int m0, int mtiles, int n0, int ntiles, FA fa, FB fb,
                                            bool lower) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    if (!lower && mtiles == 16) {
        // Full-height rectangular update (the in-tile update of the right solve, on the critical path of the Cholesky): the
        // warp's two m-tiles (warp, warp + 8) advance together, so every B fragment is loaded once for two DMMAs.
        // Per entry the same accumulation in the same order as below.
        const int mt0 = warp, mt1 = warp + 8;
        double a0[NB / 4], a1[NB / 4];
#pragma unroll
        for (int kk = 0; kk < NB / 4; ++kk) {
            a0[kk] = fa(mt0 * 8 + g, kk * 4 + tq);
            a1[kk] = fa(mt1 * 8 + g, kk * 4 + tq);
        }
        for (int nt0 = 0; nt0 < ntiles; nt0 += 4) {
            double c0[4][2], c1[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nt0 + u < ntiles) {
                    const int n = n0 + (nt0 + u) * 8 + 2 * tq, ma = m0 + mt0 * 8 + g, mb = m0 + mt1 * 8 + g;
                    c0[u][0] = -Cs[n * SC + ma];
                    c0[u][1] = -Cs[(n + 1) * SC + ma];
                    c1[u][0] = -Cs[n * SC + mb];
                    c1[u][1] = -Cs[(n + 1) * SC + mb];
                }
            }
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (nt0 + u < ntiles) {
                        const double b = fb(kk * 4 + tq, (nt0 + u) * 8 + g);
                        dmma(c0[u][0], c0[u][1], a0[kk], b);
                        dmma(c1[u][0], c1[u][1], a1[kk], b);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nt0 + u < ntiles) {
                    const int n = n0 + (nt0 + u) * 8 + 2 * tq, ma = m0 + mt0 * 8 + g, mb = m0 + mt1 * 8 + g;
                    Cs[n * SC + ma] = -c0[u][0];
                    Cs[(n + 1) * SC + ma] = -c0[u][1];
                    Cs[n * SC + mb] = -c1[u][0];
                    Cs[(n + 1) * SC + mb] = -c1[u][1];
                }
            }
        }
        return;
    }
    for (int mt = warp; mt < mtiles; mt += 8) {
        double a[NB / 4];
#pragma unroll
        for (int kk = 0; kk < NB / 4; ++kk) a[kk] = fa(mt * 8 + g, kk * 4 + tq);
        int ntl = ntiles;
        if (lower) {
            int lim = (m0 + mt * 8 - n0) / 8 + 1;   // tiles with n-tile start <= m-tile start
            ntl = lim < ntiles ? lim : ntiles;
        }
        for (int nt0 = 0; nt0 < ntl; nt0 += 4) {
            double c[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int nt = nt0 + u;
                if (nt < ntl) {
                    int n = n0 + nt * 8 + 2 * tq, m = m0 + mt * 8 + g;
                    c[u][0] = -Cs[n * SC + m];
                    c[u][1] = -Cs[(n + 1) * SC + m];
                }
            }
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int nt = nt0 + u;
                    if (nt < ntl) {
                        double b = fb(kk * 4 + tq, nt * 8 + g);
                        dmma(c[u][0], c[u][1], a[kk], b);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int nt = nt0 + u;
                if (nt < ntl) {
                    int n = n0 + nt * 8 + 2 * tq, m = m0 + mt * 8 + g;
                    Cs[n * SC + m] = -c[u][0];
                    Cs[(n + 1) * SC + m] = -c[u][1];
                }
            }
        }
    }
}

// Cholesky of the 32x32 diagonal block at offset o of the tile in Cs, by ONE warp (row per lane, registers + shuffles).
// The dependent chain per pivot is kept short: every lane tracks its OWN diagonal entry in `diag` (updated with lane-local
// FMAs) and evaluates rsqrt on it each step, so a step costs fma -> rsqrt -> one shuffle -> mul; the column shuffles for the
// off-diagonal updates stay off the chain.  L(c,c) = a * rsqrt(a), 1/L(c,c) = rsqrt(a).
// A separate non-inlined function: inlined into the unrolled / rolled block loop of potrf_tile the compiler cannot prove the
// warp converged and wraps each of the ~1000 shuffles in WARPSYNC.COLLECTIVE / ENDCOLLECTIVE (round 1: ~3x slower per pivot).
static __device__ __noinline__ void potrf_diag32(double* Cs, double* invd, int* sh_info, int o) {
    const int lane = threadIdx.x & 31;
    {
            double a[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) a[c] = Cs[(o + c) * SC + o + lane];
            double diag = Cs[(o + lane) * SC + o + lane];   // (not a[lane]: that would put a[] in local memory)
            int first_bad = 0;   // 1-based column of the first bad pivot seen by this lane (its own column only)
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                const bool pos = diag >= 2.2250738585072014e-308;  // positive normal (also catches NaN); lane c's counts
                const double inv_own = rsqrt_pos(diag);            // NaN / inf for a bad pivot: flagged below, never used
                const double inv = __shfl_sync(0xffffffffu, inv_own, c);
                if (lane == c && !pos) first_bad = c + 1;          // branch-free in SASS: keeps the block one basic block
                const double l = a[c] * inv;                       // column c of L below the diagonal (lanes > c)
                if (lane == c) invd[o + c] = inv;
                // The finished column goes to its place in shared memory at once: that store is also how the other lanes
                // get L(c2, c) for their updates (broadcast LDS.128, two entries per load) -- the 2 x 496 shuffles of the
                // first version made this function 48 KB of straight-line code, more than an SM's instruction cache.
                double* colc = Cs + (o + c) * SC + o;
                if (lane >= c) colc[lane] = lane == c ? diag * inv : l;
                diag = fma(-l, l, diag);                           // own pivot, no shuffle (only lanes > c still need it)
                if (c + 1 < NB) {                                  // the next pivot's column stays on the shuffle: it is on the chain
                    const double l1 = __shfl_sync(0xffffffffu, l, c + 1);
                    a[c + 1] = fma(-l, l1, a[c + 1]);
                }
                __syncwarp();
                if (((c + 2) & 1) != 0 && c + 2 < NB) a[c + 2] = fma(-l, colc[c + 2], a[c + 2]);
#pragma unroll
                for (int c2 = (c + 3) & ~1; c2 + 1 < NB; c2 += 2) {   // entries on/above the diagonal are never read again
                    const double2 v = *reinterpret_cast<const double2*>(colc + c2);
                    a[c2] = fma(-l, v.x, a[c2]);
                    a[c2 + 1] = fma(-l, v.y, a[c2 + 1]);
                }
            }
            // LAPACK info = first non-positive pivot: smallest failing column over the lanes
            unsigned bad = __ballot_sync(0xffffffffu, first_bad != 0);
            if (bad != 0 && lane == 0) atomicCAS(sh_info, 0, o + __ffs(bad));
        }
}

// Named barriers (barrier 0 is __syncthreads): producer / consumer hand-over between warp 0 and warps 1..7 inside potrf_tile.
// ---- in-situ timeline of the dataflow Cholesky (debug builds with -DGEORDD_POTRF_TRACE only; scripts/gpu_potrf_trace.py) ----
#ifdef GEORDD_POTRF_TRACE
static __device__ unsigned long long g_ptrace_job[8192 * 16];    // per job: 0 picked, 1 updates done, 2 tile op begins, 3 done, 4.. waits
static __device__ unsigned long long g_ptrace_flag[65536];       // release time of every flag word
static __device__ unsigned* g_ptrace_base;
static __device__ int g_ptrace_cur[256];                         // job of the CTA (by block index)
__device__ __forceinline__ unsigned long long ptrace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PTRACE_JOB(slot) do { int jb__ = g_ptrace_cur[blockIdx.x]; if (jb__ < 8192) g_ptrace_job[jb__ * 16 + (slot)] = ptrace_now(); } while (0)
#define PTRACE_IT(o, k) do { if ((o) / NB < 2 && 4 + 6 * ((o) / NB) + (k) < 13) PTRACE_JOB(4 + 6 * ((o) / NB) + (k)); } while (0)
#define PTRACE_FLAG(ptr) do { long o__ = (ptr) - g_ptrace_base; if (o__ >= 0 && o__ < 65536) g_ptrace_flag[o__] = ptrace_now(); } while (0)
#else
#define PTRACE_JOB(slot) do {} while (0)
#define PTRACE_IT(o, k) do {} while (0)
#define PTRACE_FLAG(ptr) do {} while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- Cholesky of the 128x128 lower triangle held in Cs (in place).  sh_info: shared int, must be
// 0 on entry; set to (local index + 1) of the first non-positive pivot.  The strict upper triangle
// is left untouched (caller zeroes it on write-out).  invd[128] receives 1/L(c,c).
//
// Blocked by 32.  Iteration o: (1) 32x32 diagonal block by warp 0 (a chain of 32 dependent pivots, ~200 cycles each: the
// one part of the whole factorisation that nothing can be overlapped with on this CTA); (2) the panel below it, one row per
// thread; (3) the trailing update with DMMA.  LOOK-AHEAD (round 2): warp 0 owns the 32 panel rows that are the NEXT diagonal
// block's rows, so after (2) it updates just that block and goes straight on to factor it, while warps 1..7 store / publish
// block column o and run the rest of the trailing update.  The update leaves the chain: 4 x diag + 3 x (panel + 32x32 update).
//
// Progressive publication (Lg != nullptr; the dataflow Cholesky): block column b = columns [32b, 32b+32) of the factor is
// final as soon as iteration b has solved the panel below its diagonal block.  It is then written to the stored tile image
// Lg (TBP: (m, n) at n*TLD + m, explicit zeros above the diagonal and in the 4 pad rows) and flag sub[b] is released -- the
// right solves of the tiles below this one start on block column 0 while this CTA is still factoring columns 32..127, so
// the three links of the critical path (last update -> potrf -> right solve) overlap instead of running back to back.
// After a non-positive pivot nothing more is published (the caller raises the abort word, which releases every waiter).
__device__ inline void potrf_tile(double* Cs, double* invd, int* sh_info, long long* prof = nullptr, double* Lg = nullptr,
                                  unsigned* sub = nullptr, unsigned epoch = 0) {
    const int tid = threadIdx.x, warp = tid >> 5;
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int k) { if (prof) { long long t = clock64(); if (tid == 0) prof[k] += t - tp; tp = t; } };
    // threads 32..255; rows above the diagonal and pad rows are written as 0
    auto store_block_col = [&](int o, int t0, int nt) {
        for (int idx = t0; idx < NB * TLD; idx += nt) {
            const int n = o + idx / TLD, m = idx % TLD;
            Lg[(long)n * TLD + m] = (m < TILE && m >= n) ? Cs[n * SC + m] : 0.0;
        }
    };
    bool bad = false;   // a non-positive pivot was seen (uniform)
#ifdef GEORDD_POTRF_NO_LOOKAHEAD
    // A/B variant (round 2's first step: progressive publication only): diag -> panel -> full trailing update, in sequence.
    for (int o = 0; o < TILE; o += NB) {
        if (warp == 0) potrf_diag32(Cs, invd, sh_info, o);
        __syncthreads();
        lap(0);
        if (Lg) bad = bad || ld_volatile_int(sh_info) != 0;
        const int rest = TILE - o - NB;
        if (rest == 0) {
            if (Lg && !bad) {
                store_block_col(o, tid, NTHREADS);
                __syncthreads();
                if (tid == 0) st_release(sub + o / NB, epoch);
            }
            break;
        }
        if (tid < rest) {
            const int r = o + NB + tid;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = Cs[(o + c) * SC + r];
            forward_subst32(x, invd + o, Cs + o * SC + o, SC);
#pragma unroll
            for (int c = 0; c < NB; ++c) Cs[(o + c) * SC + r] = x[c];
        }
        __syncthreads();
        lap(1);
        if (Lg && !bad) store_block_col(o, tid, NTHREADS);
        {
            const double* P = Cs + o * SC + (o + NB);
            tile_update(
                Cs, o + NB, rest / 8, o + NB, rest / 8, [&](int m, int k) { return P[k * SC + m]; },
                [&](int k, int n) { return P[k * SC + n]; }, true);
        }
        __syncthreads();
        lap(2);
        if (Lg && !bad && tid == 0) st_release(sub + o / NB, epoch);
    }
    return;
#endif
#ifdef GEORDD_POTRF_TRACE
    const long long c0__ = clock64();
    const unsigned long long g0__ = ptrace_now();
#endif
    if (warp == 0) potrf_diag32(Cs, invd, sh_info, 0);
    if (Lg && tid == 0) PTRACE_JOB(15);   // first diagonal block factored (warp 0, before the barrier)
#ifdef GEORDD_POTRF_TRACE
    if (Lg && tid == 0) {   // cycles and nanoseconds of this call: the SM clock the chain actually runs at
        int jb__ = g_ptrace_cur[blockIdx.x];
        if (jb__ < 8192) { g_ptrace_job[jb__ * 16 + 14] = (unsigned long long)(clock64() - c0__); g_ptrace_job[jb__ * 16 + 13] = ptrace_now() - g0__; }
    }
#endif
    __syncthreads();
    lap(0);
    for (int o = 0; o < TILE; o += NB) {
        if (Lg) bad = bad || ld_volatile_int(sh_info) != 0;   // uniform: written before the last barrier
        const int rest = TILE - o - NB;
        if (rest == 0) {
            if (Lg && !bad) {
                store_block_col(o, tid, NTHREADS);
                __syncthreads();
                if (tid == 0) { st_release(sub + o / NB, epoch); PTRACE_FLAG(sub + o / NB); }
            }
            break;
        }
        // 2. panel below: X = C * L^-T, one row per thread (warp 0: rows o+32..o+63 = the next diagonal block's rows)
        if (tid < rest) {
            const int r = o + NB + tid;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = Cs[(o + c) * SC + r];
            forward_subst32(x, invd + o, Cs + o * SC + o, SC);
#pragma unroll
            for (int c = 0; c < NB; ++c) Cs[(o + c) * SC + r] = x[c];
        }
        const double* P = Cs + o * SC + (o + NB);   // panel: P(m,k) at P[k*SC + m], m relative to row o+32
        if (warp == 0) {
            named_bar_arrive(1, NTHREADS);   // my panel rows are written (the others need them for their update)
            __syncwarp();
            if (Lg && tid == 0) PTRACE_IT(o, 0);
            // 3a. next diagonal block -= P0 P0^T (the 10 lower 8x8 tiles), from this warp's own 32 panel rows.  A and B
            //     fragments coincide (same panel), all ten accumulators are live so the DMMAs issue back to back.
            const int lane = tid & 31, g = lane >> 2, tq = lane & 3;
            double acc[10][2];
            {
                int t = 0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt <= mt; ++nt, ++t) {
                        const int n = o + NB + nt * 8 + 2 * tq, m = o + NB + mt * 8 + g;
                        acc[t][0] = -Cs[n * SC + m];
                        acc[t][1] = -Cs[(n + 1) * SC + m];
                    }
            }
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk) {
                double f[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) f[q] = P[(kk * 4 + tq) * SC + q * 8 + g];
                int t = 0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt <= mt; ++nt, ++t) dmma(acc[t][0], acc[t][1], f[mt], f[nt]);
            }
            {
                int t = 0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt <= mt; ++nt, ++t) {
                        const int n = o + NB + nt * 8 + 2 * tq, m = o + NB + mt * 8 + g;
                        Cs[n * SC + m] = -acc[t][0];
                        Cs[(n + 1) * SC + m] = -acc[t][1];
                    }
            }
            __syncwarp();
            lap(1);
            if (Lg && tid == 0) PTRACE_IT(o, 1);
            // 1'. factor the next diagonal block while warps 1..7 finish the trailing update
            potrf_diag32(Cs, invd, sh_info, o + NB);
            lap(0);
            if (Lg && tid == 0) PTRACE_IT(o, 2);
        } else {
            named_bar_sync(1, NTHREADS);     // all panel rows of this iteration are in shared memory
            if (Lg && tid == 32) PTRACE_IT(o, 3);
            if (Lg && !bad) {                // block column o is final: store it, publish it
                store_block_col(o, tid - 32, NTHREADS - 32);
                named_bar_sync(2, NTHREADS - 32);
                if (tid == 32) { PTRACE_IT(o, 4); st_release(sub + o / NB, epoch); PTRACE_FLAG(sub + o / NB); }
            }
            // 3b. trailing lower triangle -= P P^T for the rows below the next diagonal block (8x8 m-tiles 4.. of the trailing part)
            const int mtiles = rest / 8;
            const int lane = tid & 31, g = lane >> 2, tq = lane & 3;
            // Warp 4 shares warp 0's SM sub-partition, and scalar FP64 instructions of the pivot chain queue behind DMMAs
            // issued there (measured: diag32 26.6k -> 31.6k cycles with warp 4 updating): six warps update, warp 4 sits out.
            const int uw = warp < 4 ? warp - 1 : warp - 2;   // 0..5 for warps 1,2,3,5,6,7
            if (warp != 4)
            for (int mt = NB / 8 + uw; mt < mtiles; mt += 6) {
                double a[NB / 4];
#pragma unroll
                for (int kk = 0; kk < NB / 4; ++kk) a[kk] = P[(kk * 4 + tq) * SC + mt * 8 + g];
                for (int nt0 = 0; nt0 <= mt; nt0 += 4) {
                    double c[4][2];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nt = nt0 + u;
                        if (nt <= mt) {
                            const int n = o + NB + nt * 8 + 2 * tq, m = o + NB + mt * 8 + g;
                            c[u][0] = -Cs[n * SC + m];
                            c[u][1] = -Cs[(n + 1) * SC + m];
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < NB / 4; ++kk) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int nt = nt0 + u;
                            if (nt <= mt) dmma(c[u][0], c[u][1], a[kk], P[(kk * 4 + tq) * SC + nt * 8 + g]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nt = nt0 + u;
                        if (nt <= mt) {
                            const int n = o + NB + nt * 8 + 2 * tq, m = o + NB + mt * 8 + g;
                            Cs[n * SC + m] = -c[u][0];
                            Cs[(n + 1) * SC + m] = -c[u][1];
                        }
                    }
                }
            }
            if (Lg && tid == 32) PTRACE_IT(o, 5);
        }
        __syncthreads();
        lap(2);
    }
}

// Load columns [o, o+NB) of the lower-triangular 128x128 tile Lt (global, leading dim ld), rows
// [o, 128), into Lp(k, m) = Lp[k * SC + m] (m absolute row in the tile) and the pivot reciprocals
// of that block into invd[0..NB).
__device__ __forceinline__ void load_col_panel(double* Lp, double* invd, const double* Lt, long ld, int o) {
    const int rows = TILE - o;
    for (int id = threadIdx.x; id < NB * rows; id += NTHREADS) {
        int k = id / rows, m = o + id % rows;
        Lp[k * SC + m] = ldcg(Lt + (long)(o + k) * ld + m);
    }
    if (threadIdx.x < NB) invd[threadIdx.x] = 1.0 / ldcg(Lt + (long)(o + threadIdx.x) * ld + o + threadIdx.x);
}

// Load rows [o, o+NB) of tile Lt, columns [0, o+NB), into Lr(c, k) = Lr[c * LRS + k], k = row - o.
__device__ __forceinline__ void load_row_panel(double* Lr, double* invd, const double* Lt, long ld, int o) {
    const int cols = o + NB;
    for (int id = threadIdx.x; id < NB * cols; id += NTHREADS) {
        int c = id / NB, k = id % NB;
        Lr[c * LRS + k] = ldcg(Lt + (long)c * ld + o + k);
    }
    if (threadIdx.x < NB) invd[threadIdx.x] = 1.0 / ldcg(Lt + (long)(o + threadIdx.x) * ld + o + threadIdx.x);
}

// cp.async copy of columns [o, o+NB) of the lower-triangular 128x128 tile Lt, rows [o, 128), into Lp(k, m) = Lp[k*SC + m]
// (m absolute row in the tile).  o is a multiple of 32 and ld*8 a multiple of 16, so every 16-byte piece is aligned.
__device__ __forceinline__ void load_col_panel_async(double* Lp, const double* Lt, long ld, int o) {
    const int half = (TILE - o) / 2;   // 16-byte pieces per column
    for (int id = threadIdx.x; id < NB * half; id += NTHREADS) {
        int k = id / half, m = o + 2 * (id % half);
        cp_async16(Lp + k * SC + m, Lt + (long)(o + k) * ld + m);
    }
    cp_async_commit();
}

// ---- X * Lt^T = C  (C: 128 x 128 in Cs, overwritten by X).  Rows are independent. -------------
// The column panels of Lt are double-buffered (Lp0 / Lp1, NB*SC doubles each): the copy of the next panel runs under
// the substitution and update of the current block.
//
// Dataflow form (dtile != nullptr): Lt is the diagonal tile of this tile column, possibly still being factored by another
// CTA.  *dtile == epoch says the whole tile is final; otherwise dsub[b] == epoch says block column b is (potrf_tile's
// progressive publication) and iteration b waits for exactly that.  X's own block columns are final after the substitution
// of their iteration: they are written to the stored tile image Xg and released through xsub[b], so that the diagonal job
// of the next tile column consumes them chunk by chunk.  Returns false (uniformly) if a wait was aborted.
__device__ inline bool trsm_right_tile(double* Cs, double* Lp0, double* Lp1, double* invd, const double* Lt, long ld,
                                       long long* prof = nullptr, const unsigned* dtile = nullptr, const unsigned* dsub = nullptr,
                                       double* Xg = nullptr, unsigned* xsub = nullptr, unsigned epoch = 0, int* abort_word = nullptr) {
    const int tid = threadIdx.x;
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int k) { if (prof) { long long t = clock64(); if (tid == 0) prof[k] += t - tp; tp = t; } };
    bool whole = true;   // the diagonal tile is known to be complete
    if (dtile) {
        int w = 1;
        if (tid == 0) w = ld_acquire(dtile) == epoch;
        whole = __syncthreads_and(w) != 0;
    }
    int loaded = -1;     // highest panel whose copy has been issued
    for (int o = 0, b = 0; o < TILE; o += NB, ++b) {
        double* Lp = (b & 1) ? Lp1 : Lp0;
        if (loaded < b) {
            if (!whole) {
                int ok = 1;
                if (tid == 0) { ok = wait_flag(dsub + b, epoch, abort_word); PTRACE_JOB(4 + b); }
                if (!__syncthreads_and(ok)) return false;
            }
            load_col_panel_async(Lp, Lt, ld, o);
            loaded = b;
        }
        bool pre = o + NB < TILE;
        if (pre && !whole) {   // prefetch the next panel only if it has already been published
            int ok = 1;
            if (tid == 0) ok = ld_acquire(dsub + b + 1) == epoch;
            pre = __syncthreads_and(ok) != 0;
        }
        if (pre) {
            load_col_panel_async((b & 1) ? Lp0 : Lp1, Lt, ld, o + NB);
            loaded = b + 1;
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (tid == 0) PTRACE_JOB(8 + b);   // panel b is in shared memory
        if (tid < NB) invd[tid] = 1.0 / Lp[tid * SC + o + tid];
        __syncthreads();
        lap(0);
        if (tid < TILE) {
            const int r = tid;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = Cs[(o + c) * SC + r];
            forward_subst32(x, invd, Lp + o, SC);
#pragma unroll
            for (int c = 0; c < NB; ++c) Cs[(o + c) * SC + r] = x[c];
        }
        __syncthreads();
        lap(1);
        if (Xg) {   // block column b of X is final
            for (int idx = tid; idx < NB * TLD; idx += NTHREADS) {
                const int n = o + idx / TLD, m = idx % TLD;
                Xg[(long)n * TLD + m] = (m < TILE) ? Cs[n * SC + m] : 0.0;
            }
        }
        const int rest = TILE - o - NB;
        if (rest > 0) {
            // C(:, n') -= X(:, o..o+NB) * L(n', o..o+NB)^T  for n' >= o+NB
            const double* X = Cs + o * SC;
            tile_update(
                Cs, 0, TILE / 8, o + NB, rest / 8, [&](int m, int k) { return X[k * SC + m]; },
                [&](int k, int n) { return Lp[k * SC + o + NB + n]; }, false);
        }
        __syncthreads();
        lap(2);
        if (Xg && tid == 0) { st_release(xsub + b, epoch); PTRACE_FLAG(xsub + b); }   // every thread's stores precede the barrier above
    }
    return true;
}

// ---- Lt * X = C  (forward; C: 128 x ncols in Cs, overwritten).  Columns are independent. -------
__device__ inline void solve_fwd_tile(double* Cs, int ncols, double* Lp, double* invd, const double* Lt, long ld) {
    const int tid = threadIdx.x;
    for (int o = 0; o < TILE; o += NB) {
        load_col_panel(Lp, invd, Lt, ld, o);
        __syncthreads();
        if (tid < ncols) {
            double* col = Cs + tid * SC + o;
            double x[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) x[r] = col[r];
            forward_subst32(x, invd, Lp + o, SC);
#pragma unroll
            for (int r = 0; r < NB; ++r) col[r] = x[r];
        }
        __syncthreads();
        const int rest = TILE - o - NB;
        if (rest > 0) {
            // C(m', :) -= L(m', o..o+NB) * X(o..o+NB, :)   for m' >= o+NB
            tile_update(
                Cs, o + NB, rest / 8, 0, ncols / 8, [&](int m, int k) { return Lp[k * SC + o + NB + m]; },
                [&](int k, int n) { return Cs[n * SC + o + k]; }, false);
        }
        __syncthreads();
    }
}

// ---- Lt^T * X = C  (backward with the transposed lower tile). -----------------------------------
__device__ inline void solve_bwd_tile(double* Cs, int ncols, double* Lr, double* invd, const double* Lt, long ld) {
    const int tid = threadIdx.x;
    for (int o = TILE - NB; o >= 0; o -= NB) {
        load_row_panel(Lr, invd, Lt, ld, o);
        __syncthreads();
        if (tid < ncols) {
            double* col = Cs + tid * SC + o;
            double x[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) x[r] = col[r];
#pragma unroll
            for (int r = NB - 1; r >= 0; --r) {
                x[r] *= invd[r];
#pragma unroll
                for (int r2 = 0; r2 < r; ++r2) x[r2] -= x[r] * Lr[(o + r2) * LRS + r];
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) col[r] = x[r];
        }
        __syncthreads();
        if (o > 0) {
            // C(m', :) -= L(o..o+NB, m')^T * X(o..o+NB, :)   for m' < o
            tile_update(
                Cs, 0, o / 8, 0, ncols / 8, [&](int m, int k) { return Lr[m * LRS + k]; },
                [&](int k, int n) { return Cs[n * SC + o + k]; }, false);
        }
        __syncthreads();
    }
}

// ---- Tile solve with INVERTED 32x32 diagonal blocks (the batched solves of the null simulations, round 2) -----------------
// The substitution above is a chain of scalar FP64 operations, and inside solve_seq_kernel every one of them queues behind
// the DMMAs of the co-resident CTA (in situ: 79 k cycles per job against 27 k alone).  Here the four 32x32 diagonal blocks of
// the diagonal tile are inverted ONCE per factor (launch_diag_inverse), the block solve becomes X_o = inv(L_oo) C_o on the
// tensor pipe -- what cuBLAS / MAGMA trsm do -- and, because every operation acts on columns independently, each warp
// solves its OWN 8 columns of the 128 x 64 block from start to end: no CTA barrier, no panel staging in shared memory.
//   Dv : the tile's four inverse blocks in shared memory, block b at Dv + b*DINV_BLK, inv(m,k) at [k*LRS + m]
//   Ld : the diagonal tile's stored image in global memory (TBP; (m,n) at n*TLD + m, zeros above the diagonal): the
//        off-diagonal blocks of the tile feed the updates straight from L2 into A fragments
// FWD:  X = L^-1 C, blocks ascending;   BWD:  X = L^-T C, blocks descending.
constexpr int DINV_BLK = NB * LRS;           // 1152 doubles
constexpr int DINV_TILE = (TILE / NB) * DINV_BLK;   // 4608 doubles = 36 864 B, one bulk copy

template <bool BWD>
__device__ __forceinline__ void solve_tile_inv(double* Cs, const double* Dv, const double* __restrict__ Ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    double* Cw = Cs + warp * 8 * SC;   // this warp's 8 columns: C(m, n) at Cw[n*SC + m]
#pragma unroll 1
    for (int step = 0; step < TILE / NB; ++step) {
        const int o = BWD ? TILE - NB - NB * step : NB * step;
        const double* D = Dv + (o / NB) * DINV_BLK;
        // (1) X_o = inv(L_oo) C_o  (BWD: inv(L_oo)' C_o): four 8x8 output tiles, K = 32, the triangular half of the k steps
        double b[NB / 4];
#pragma unroll
        for (int kk = 0; kk < NB / 4; ++kk) b[kk] = Cw[g * SC + o + kk * 4 + tq];
        double x[NB / 8][2];
#pragma unroll
        for (int mt = 0; mt < NB / 8; ++mt) {
            x[mt][0] = x[mt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk) {
                if (BWD ? (kk >= 2 * mt) : (kk <= 2 * mt + 1)) {
                    const double a = BWD ? D[(mt * 8 + g) * LRS + kk * 4 + tq] : D[(kk * 4 + tq) * LRS + mt * 8 + g];
                    dmma(x[mt][0], x[mt][1], a, b[kk]);
                }
            }
        }
        __syncwarp();   // every lane has read its C_o fragments
#pragma unroll
        for (int mt = 0; mt < NB / 8; ++mt) {
            Cw[(2 * tq) * SC + o + mt * 8 + g] = x[mt][0];
            Cw[(2 * tq + 1) * SC + o + mt * 8 + g] = x[mt][1];
        }
        __syncwarp();
        // (2) rows outside the block: C(r, :) -= L(r, o..o+31) X_o  (BWD: -= L(o..o+31, r)' X_o), four m-tiles at a time
        const int nm = BWD ? o / 8 : (TILE - NB - o) / 8;   // 12, 8, 4, 0: always a multiple of 4
        const int r0 = BWD ? 0 : o + NB;
        if (nm == 0) continue;
#pragma unroll
        for (int kk = 0; kk < NB / 4; ++kk) b[kk] = Cw[g * SC + o + kk * 4 + tq];   // X_o as the B operand
#pragma unroll 1
        for (int m0 = 0; m0 < nm; m0 += 4) {
            double a[4][NB / 4], c[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = r0 + (m0 + u) * 8 + g;
#pragma unroll
                for (int kk = 0; kk < NB / 4; ++kk)
                    a[u][kk] = BWD ? ldcg(Ld + (long)m * TLD + o + kk * 4 + tq) : ldcg(Ld + (long)(o + kk * 4 + tq) * TLD + m);
                c[u][0] = -Cw[(2 * tq) * SC + m];
                c[u][1] = -Cw[(2 * tq + 1) * SC + m];
            }
#pragma unroll
            for (int kk = 0; kk < NB / 4; ++kk)
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma(c[u][0], c[u][1], a[u][kk], b[kk]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = r0 + (m0 + u) * 8 + g;
                Cw[(2 * tq) * SC + m] = -c[u][0];
                Cw[(2 * tq + 1) * SC + m] = -c[u][1];
            }
        }
        __syncwarp();
    }
}

}  // namespace geordd
