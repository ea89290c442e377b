// K3/K4: batched triangular solves and triangular multiply over many right-hand sides (= null
// draws), as persistent dataflow kernels on the DMMA pipe.
//
// Replaces, batched over draws, the per-draw level-2 BLAS of the reference hot loops:
//   prior_rand      y = mu + U'z            dtrmv   src/hypotest.jl:13-18           -> SOLVE_TRMM
//   update_alpha!   alpha = cK \ (y - m)     dpotrs  src/gp_utils.jl:15-18           -> SOLVE_FWD + SOLVE_BWD
//   PDcliff \ mu                             dpotrs  src/hypotest_chi2.jl:16          -> SOLVE_FWD + SOLVE_BWD
// and the whitening solve of predict_f (src/cliff_face.jl:4-5) with the sentinels as columns.
//
// Right-hand sides are cut into slabs of 64 columns, the factor into 128-row blocks.  Job (rb, s)
// produces the final 128x64 block X(rb, s):
//     acc = B(rb,s) - sum_j L(rb,j) X(j,s)      (DMMA mainloop; L from L2/HBM, X from this slab)
//     X(rb,s) = L(rb,rb)^-1 acc                 (substitution in shared memory)
// Jobs are issued level by level (rb ascending for forward, descending for backward) through an
// atomic counter; a job waits only on lower-index jobs of its own slab (acquire-polled flags).
// With many slabs nobody ever waits; with few slabs (cliff face: nb = 100 sentinels) the row blocks
// of one slab pipeline across CTAs instead of serialising on one SM.
#include <cstdlib>
#include "kernels.h"
#include "mainloop.cuh"
#include "tileops.cuh"

namespace geordd {

constexpr int SLAB = 64;
constexpr int CHUNKS_PER_TILE_S = TILE / BK;

// All three modes read the factor with the row index contiguous: forward / TRMM from the lower tiles of L,
// backward from the mirrored upper tiles (L' stored above the diagonal by launch_mirror_factor).
// Operand feed: the triangular multiply and the XB solves (right-hand sides in ZB storage -- the simulation path)
// move factor chunk and right-hand-side chunk with ONE TMA block copy each, issued by thread 0; only warp 0 ever
// waits on a dataflow flag.  The column-major solves (fits, cliff face, misc callers) use the generic-proxy ring.
// XB: the right-hand sides are in ZB storage, so the flagged solves run the same one-thread TMA feed as the
// triangular multiply (exact since the stage release waits for the fragment loads, mainloop.cuh).
template <int MODE, bool XB>
struct SolveCfg {
    static constexpr bool BULK = XB || MODE == SOLVE_TRMM;
    using ML = Mainloop<SLAB, /*A_MC=*/true, /*B_NC=*/false, 4, BULK ? BLK_TMA : BLK_GEN, BULK ? BLK_TMA : ROWS_GEN>;
};

template <int MODE, bool XB>
struct SolveSrc {
    const double* L;
    const double* X;   // slab base: column 0 of this slab, row 0
    long lda, ldb;
    int rb, T, s, nslabs;
    const unsigned* flags;
    unsigned epoch;
    int* abort_word;
    __device__ __forceinline__ int ktile(int c) const {
        return MODE == SOLVE_BWD ? (T - 1 - c / CHUNKS_PER_TILE_S) : c / CHUNKS_PER_TILE_S;
    }
    __device__ __forceinline__ const double* a_ptr(int c) const {
        int kt = ktile(c), kk = (c % CHUNKS_PER_TILE_S) * BK;
        // A(m,k) = M(rb*128+m, kt*128+kk+k): 16 consecutive columns of TBP tile (rb, kt).  kt < rb: lower tile of L;
        // kt > rb (backward): the mirrored tile holding L(kt-block, rb-block)'.
        return L + tile_off(rb, kt, T) + (long)kk * TLD;
    }
    __device__ __forceinline__ const double* b_ptr(int c) const {
        if (MODE == SOLVE_TRMM) return X + (long)c * ZB_ELEMS;   // ZB storage: chunk c of this slab is one block
        int kt = ktile(c), kk = (c % CHUNKS_PER_TILE_S) * BK;
        if (XB) return X + ((long)kt * CHUNKS_PER_TILE_S + c % CHUNKS_PER_TILE_S) * ZB_ELEMS;
        return X + (long)kt * TILE + kk;   // B(k,n) = X(kt*128+kk+k, n): k contiguous
    }
    __device__ __forceinline__ bool ready(int c) const {   // warp level
        if (MODE == SOLVE_TRMM) return true;
        if (c % CHUNKS_PER_TILE_S != 0) return true;
        bool ok = warp_wait_flags(flags + (long)ktile(c) * nslabs + s, nullptr, epoch, abort_word);
        if (XB) fence_proxy_async();   // the acquire (generic proxy) before the bulk copy of the published block
        return ok;
    }
};

struct SolveParams {
    const double* L;   // TBP factor storage
    int T;
    double* B;         // FWD/BWD: in place.  TRMM: input Z
    long ldb;
    int nslabs;
    unsigned* flags;
    unsigned epoch;
    unsigned* counter;
    int* abort_word;
    SolveEpilogue epi;
};

constexpr int SOLVE_CS_ELEMS = SLAB * SC;                       // 8448
constexpr int SOLVE_PANEL_ELEMS = TILE * LRS;                   // 4608 >= NB*SC = 4224
constexpr int SOLVE_RING_ELEMS = SolveCfg<0, false>::ML::RING_ELEMS;   // identical for all modes
constexpr int SOLVE_WORK_ELEMS = SOLVE_CS_ELEMS + SOLVE_PANEL_ELEMS + NB;
constexpr int SOLVE_SMEM_ELEMS = (SOLVE_RING_ELEMS > SOLVE_WORK_ELEMS ? SOLVE_RING_ELEMS : SOLVE_WORK_ELEMS) + SolveCfg<0, false>::ML::BAR_WORDS + 1;   // + the inverse-block barrier
static_assert(SOLVE_PANEL_ELEMS >= DINV_TILE, "the panel area holds the four inverse diagonal blocks of a tile");
constexpr size_t SOLVE_SMEM = (size_t)SOLVE_SMEM_ELEMS * sizeof(double);

template <int MODE, bool XB>
__global__ void __launch_bounds__(NTHREADS, 2) solve_dataflow_kernel(SolveParams p) {
    using ML = typename SolveCfg<MODE, XB>::ML;
    static_assert(SolveCfg<MODE, XB>::ML::RING_ELEMS == SOLVE_RING_ELEMS, "one shared-memory budget for all variants");
    static_assert(ML::SMEM_BYTES + ML::BAR_WORDS * 8 <= SOLVE_SMEM, "stage ring must fit");
    extern __shared__ __align__(16) double smem[];
    double* Cs = smem;
    double* Lp = smem + SOLVE_CS_ELEMS;
    double* invd = Lp + SOLVE_PANEL_ELEMS;
    __shared__ int sh_job;
    const int tid = threadIdx.x;
    Ring ring;
    ML::ring_init(ring, reinterpret_cast<uint64_t*>(smem + SOLVE_SMEM_ELEMS - ML::BAR_WORDS), p.abort_word);
    const int njobs = p.T * p.nslabs;
    const long ncols = (long)p.nslabs * SLAB;

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
        // FWD ascending, BWD descending (dependencies); TRMM has none and is dealt longest jobs first (row block rb multiplies
        // rb + 1 tiles), so that the kernel's tail consists of its shortest jobs.
        const int rb = (MODE == SOLVE_FWD) ? lvl : (p.T - 1 - lvl);
        double* Bslab = (MODE == SOLVE_TRMM || XB) ? p.B + (long)s * (p.T * CHUNKS_PER_TILE_S) * ZB_ELEMS   // ZB: slab s
                                                   : p.B + (long)s * SLAB * p.ldb;
        // element (row m of block rb, column n of the slab) of an XB right-hand-side slab
        auto xb_at = [&](int m, int n) { return Bslab + ((long)rb * CHUNKS_PER_TILE_S + (m >> 4)) * ZB_ELEMS + n * ZB_LD + (m & 15); };

        typename ML::Acc acc;
        if (MODE == SOLVE_TRMM) {
            ML::zero(acc);
        } else {
            const double* Brb = Bslab + (long)rb * TILE;
#pragma unroll
            for (int a = 0; a < ML::MI; ++a)
#pragma unroll
                for (int b = 0; b < ML::NJ; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        acc.v[a][b][e] = XB ? -ldcg(xb_at(ML::acc_row(a), ML::acc_col(b, e)))
                                            : -ldcg(Brb + (long)ML::acc_col(b, e) * p.ldb + ML::acc_row(a));
        }
        SolveSrc<MODE, XB> src{p.L, Bslab, 0, p.ldb, rb, p.T, s, p.nslabs, p.flags, p.epoch, p.abort_word};
        int nchunks = (MODE == SOLVE_FWD) ? rb * CHUNKS_PER_TILE_S
                    : (MODE == SOLVE_BWD) ? (p.T - 1 - rb) * CHUNKS_PER_TILE_S
                                          : (rb + 1) * CHUNKS_PER_TILE_S;
        bool ok = ML::run(smem, ring, nchunks, src, acc);
        if (__syncthreads_or(!ok)) break;   // also: every warp is done reading the ring before Cs overwrites it
        const double sgn = (MODE == SOLVE_TRMM) ? 1.0 : -1.0;
#pragma unroll
        for (int a = 0; a < ML::MI; ++a)
#pragma unroll
            for (int b = 0; b < ML::NJ; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) Cs[ML::acc_col(b, e) * SC + ML::acc_row(a)] = sgn * acc.v[a][b][e];
        __syncthreads();

        const double* Ldiag = p.L + tile_off(rb, rb, p.T);
        if (MODE == SOLVE_FWD) solve_fwd_tile(Cs, SLAB, Lp, invd, Ldiag, TLD);
        if (MODE == SOLVE_BWD) solve_bwd_tile(Cs, SLAB, Lp, invd, Ldiag, TLD);

        const SolveEpilogue& e = p.epi;
        if (MODE == SOLVE_TRMM) {
            for (int idx = tid; idx < SLAB * TILE; idx += NTHREADS) {
                int n = idx / TILE, m = idx % TILE;
                long col = (long)s * SLAB + n;
                int r = rb * TILE + m;
                double v = Cs[n * SC + m];
                bool live = r < e.nrows;
                if (live) v += e.add_const + (r < e.split ? e.add_T : 0.0);
                if (e.out0) e.out0[e.xb ? zb_off(r, col, (int)e.ld0) : col * e.ld0 + r] = live ? v - e.sub_0 : 0.0;
                if (e.out1) e.out1[e.xb ? zb_off(r, col, (int)e.ld1) : col * e.ld1 + r] = live ? v - e.sub_1 : 0.0;
                if (live) {
                    if (r < e.split) {
                        if (e.outT) e.outT[e.xb ? zb_off(r, col, (int)e.ldT) : col * e.ldT + r] = v - e.sub_T;
                    } else {
                        if (e.outC) e.outC[e.xb ? zb_off(r - e.split, col, (int)e.ldC) : col * e.ldC + (r - e.split)] = v - e.sub_C;
                    }
                }
            }
        } else {
            double* Brb = Bslab + (long)rb * TILE;
            for (int idx = tid; idx < SLAB * TILE; idx += NTHREADS) {
                int n = idx / TILE, m = idx % TILE;
                if (XB) *xb_at(m, n) = Cs[n * SC + m];
                else Brb[(long)n * p.ldb + m] = Cs[n * SC + m];
            }
            // fused statistic partials: 4 threads per column, 32 rows each, quad shuffle-reduce
            if (e.dot0 || e.dot1) {
                const int n = tid >> 2, q = tid & 3;
                const long col = (long)s * SLAB + n;
                double s0 = 0.0, s1 = 0.0;
                // wmat is not written by this kernel: read-only cached loads (a thread's 32 rows share two L1 lines)
#pragma unroll 8
                for (int mm = 0; mm < 32; ++mm) {
                    int m = q * 32 + mm;
                    int r = rb * TILE + m;
                    double x = Cs[n * SC + m];
                    if (e.dot0) {
                        double w = x;
                        if (!e.dot_self) {
                            w = 0.0;
                            if (r < e.wlive) w = __ldg(XB ? e.wmat + zb_off(r + e.wrow0, col, (int)e.ldw) : e.wmat + col * e.ldw + r);
                            if (r < e.wrows) w += e.wshift;
                        }
                        s0 += w * x;
                    }
                    if (e.dot1) s1 += e.wvec[r] * x;
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                if (q == 0) {
                    if (e.dot0) e.dot0[(long)rb * ncols + col] = s0;
                    if (e.dot1) e.dot1[(long)rb * ncols + col] = s1;
                }
            }
            __threadfence();
        }
        fence_proxy_async();   // generic smem accesses of this job before the next job's bulk copies
        __syncthreads();
        if (MODE != SOLVE_TRMM && tid == 0) st_release(p.flags + (long)rb * p.nslabs + s, p.epoch);
    }
}

template <int MODE, bool XB>
static void launch_mode(cudaStream_t st, Sched& sched, const SolveParams& p) {
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(solve_dataflow_kernel<MODE, XB>, SOLVE_SMEM, attr_done);
    int njobs = p.T * p.nslabs;
    int grid = 2 * sched.num_sms;
    if (grid > njobs) grid = njobs;

    solve_dataflow_kernel<MODE, XB><<<grid, NTHREADS, SOLVE_SMEM, st>>>(p); ++g_kernel_launches;
}


// ---- fused solve sequence ---------------------------------------------------------------------------------
// update_alpha! of the three GPs of one batch of draws (src/hypotest_mll.jl:1-19) is six solves: forward and
// backward against L_T, L_C, L_N.  Launched one by one, every kernel ends with a tail in which CTAs idle while the
// longest (last-level) jobs finish.  Here all six run as ONE persistent kernel over a single job list ordered
//   fwd(0), fwd(1), fwd(2), bwd(0), bwd(1), bwd(2)
// (level-major inside each): CTAs that run out of jobs of one solve continue with the next one, whose jobs depend on
// nothing that is still running (a backward job waits for "forward of its slab complete", published long before).
// Right-hand sides are in ZB storage (one-thread TMA feed).
struct SeqSolve {
    const double* L;
    const double* Dinv;   // inverse 32x32 diagonal blocks (nullptr: shared-memory substitution)
    double* B;
    int T;
    int mode;        // SOLVE_FWD / SOLVE_BWD
    int flag_off;    // this solve's flag region
    int after_off;   // >= 0: flag region of the forward solve whose slab must be complete before a job may start
    int after_T;
    SolveEpilogue epi;
};
struct SeqParams {
    int nsolves, nslabs;
    int job_end[SOLVE_SEQ_MAX];
    unsigned* flags;
    unsigned epoch;
    unsigned* counter;
    int* abort_word;
    SeqSolve s[SOLVE_SEQ_MAX];
};

struct SeqSrc {
    const double* L;
    const double* X;
    long lda, ldb;
    int rb, T, s, nslabs, mode;
    const unsigned* flags;
    unsigned epoch;
    int* abort_word;
    __device__ __forceinline__ int ktile(int c) const {
        return mode == SOLVE_BWD ? (T - 1 - c / CHUNKS_PER_TILE_S) : c / CHUNKS_PER_TILE_S;
    }
    __device__ __forceinline__ const double* a_ptr(int c) const {
        return L + tile_off(rb, ktile(c), T) + (long)(c % CHUNKS_PER_TILE_S) * BK * TLD;
    }
    __device__ __forceinline__ const double* b_ptr(int c) const {
        return X + ((long)ktile(c) * CHUNKS_PER_TILE_S + c % CHUNKS_PER_TILE_S) * ZB_ELEMS;
    }
    __device__ __forceinline__ bool ready(int c) const {   // warp 0 only (one-thread feed)
        if (c % CHUNKS_PER_TILE_S != 0) return true;
        bool ok = warp_wait_flags(flags + (long)ktile(c) * nslabs + s, nullptr, epoch, abort_word);
        fence_proxy_async();
        return ok;
    }
};

__global__ void __launch_bounds__(NTHREADS, 2) solve_seq_kernel(const __grid_constant__ SeqParams p) {
    using ML = SolveCfg<SOLVE_FWD, true>::ML;
    extern __shared__ __align__(16) double smem[];
    double* Cs = smem;
    double* Lp = smem + SOLVE_CS_ELEMS;
    double* invd = Lp + SOLVE_PANEL_ELEMS;
    __shared__ int sh_job;
    const int tid = threadIdx.x;
    Ring ring;
    ML::ring_init(ring, reinterpret_cast<uint64_t*>(smem + SOLVE_SMEM_ELEMS - ML::BAR_WORDS - 1), p.abort_word);
    uint64_t* dbar = reinterpret_cast<uint64_t*>(smem + SOLVE_SMEM_ELEMS - 1);   // arrival of the inverse diagonal blocks
    if (tid == 0) {
        mbar_init(dbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    unsigned dphase = 0;
    const int njobs = p.job_end[p.nsolves - 1];
    const long ncols = (long)p.nslabs * SLAB;

    while (true) {
        if (tid == 0) {
            int jb = (int)atomicAdd(p.counter, 1u);
            if (ld_volatile_int(p.abort_word) != 0) jb = njobs;
            sh_job = jb;
        }
        __syncthreads();
        int job = sh_job;
        if (job >= njobs) break;
        int k = 0;
        while (job >= p.job_end[k]) ++k;
        if (k > 0) job -= p.job_end[k - 1];
        const SeqSolve& q = p.s[k];
        const int mode = q.mode, T = q.T;
        const int lvl = job / p.nslabs, s = job % p.nslabs;
        const int rb = (mode == SOLVE_FWD) ? lvl : (T - 1 - lvl);
        const unsigned* flags = p.flags + q.flag_off;
        double* Bslab = q.B + (long)s * (T * CHUNKS_PER_TILE_S) * ZB_ELEMS;
        auto xb_at = [&](int m, int n) { return Bslab + ((long)rb * CHUNKS_PER_TILE_S + (m >> 4)) * ZB_ELEMS + n * ZB_LD + (m & 15); };

        if (q.after_off >= 0) {   // the forward solve of this slab has written all of B (and nobody reads it any more)
            int okf = 1;
            if (tid == 0) okf = wait_flag(p.flags + q.after_off + (long)(q.after_T - 1) * p.nslabs + s, p.epoch, p.abort_word);
            if (!__syncthreads_and(okf)) break;
        }
        typename ML::Acc acc;
#pragma unroll
        for (int a = 0; a < ML::MI; ++a)
#pragma unroll
            for (int b = 0; b < ML::NJ; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc.v[a][b][e] = -ldcg(xb_at(ML::acc_row(a), ML::acc_col(b, e)));
        SeqSrc src{q.L, Bslab, 0, 0, rb, T, s, p.nslabs, mode, flags, p.epoch, p.abort_word};
        const int nchunks = (mode == SOLVE_FWD ? rb : T - 1 - rb) * CHUNKS_PER_TILE_S;
        bool ok = ML::run(smem, ring, nchunks, src, acc);
        if (__syncthreads_or(!ok)) break;   // (every warp is done reading the ring: Cs / Lp alias it)
        const double* Ldiag = q.L + tile_off(rb, rb, T);
        if (q.Dinv && tid == 0) {   // the tile's four inverse diagonal blocks: one bulk copy into the panel area
            fence_proxy_async();
            mbar_arrive_expect_tx(dbar, (unsigned)(DINV_TILE * sizeof(double)));
            bulk_g2s(Lp, q.Dinv + (long)rb * DINV_TILE, (unsigned)(DINV_TILE * sizeof(double)), dbar);
        }
#pragma unroll
        for (int a = 0; a < ML::MI; ++a)
#pragma unroll
            for (int b = 0; b < ML::NJ; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) Cs[ML::acc_col(b, e) * SC + ML::acc_row(a)] = -acc.v[a][b][e];
        __syncthreads();

        if (q.Dinv) {
            // warp-local solve of this warp's 8 columns on the tensor pipe (tileops.cuh: solve_tile_inv), no CTA barrier inside
            const bool okd = mbar_wait(dbar, dphase, p.abort_word);
            dphase ^= 1u;
            if (__syncthreads_or(!okd)) break;
            if (mode == SOLVE_FWD) solve_tile_inv<false>(Cs, Lp, Ldiag);
            else solve_tile_inv<true>(Cs, Lp, Ldiag);
            __syncthreads();   // the write-out below reads other warps' columns
        } else {
            if (mode == SOLVE_FWD) solve_fwd_tile(Cs, SLAB, Lp, invd, Ldiag, TLD);
            else solve_bwd_tile(Cs, SLAB, Lp, invd, Ldiag, TLD);
        }

        const SolveEpilogue& e = q.epi;
        for (int idx = tid; idx < SLAB * TILE; idx += NTHREADS) {
            int n = idx / TILE, m = idx % TILE;
            *xb_at(m, n) = Cs[n * SC + m];
        }
        if (e.dot0 || e.dot1) {   // same partials, same summation order as solve_dataflow_kernel
            const int n = tid >> 2, qq = tid & 3;
            const long col = (long)s * SLAB + n;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
            for (int mm = 0; mm < 32; ++mm) {
                int m = qq * 32 + mm;
                int r = rb * TILE + m;
                double x = Cs[n * SC + m];
                if (e.dot0) {
                    double w = x;
                    if (!e.dot_self) {
                        w = 0.0;
                        if (r < e.wlive) w = __ldg(e.wmat + zb_off(r + e.wrow0, col, (int)e.ldw));
                        if (r < e.wrows) w += e.wshift;
                    }
                    s0 += w * x;
                }
                if (e.dot1) s1 += e.wvec[r] * x;
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            if (qq == 0) {
                if (e.dot0) e.dot0[(long)rb * ncols + col] = s0;
                if (e.dot1) e.dot1[(long)rb * ncols + col] = s1;
            }
        }
        __threadfence();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) st_release(p.flags + q.flag_off + (long)rb * p.nslabs + s, p.epoch);
    }
}

void launch_solve_seq(cudaStream_t st, Sched& sched, const SolveSeqItem* items, int n, int ncols) {
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(solve_seq_kernel, SOLVE_SMEM, attr_done);
    SeqParams p;
    p.nsolves = n; p.nslabs = ncols / SLAB;
    p.flags = sched.flags; p.counter = sched.counter; p.abort_word = sched.abort_word;
    int off = 0, jobs = 0;
    int offs[SOLVE_SEQ_MAX];
    for (int k = 0; k < n; ++k) {
        SeqSolve& q = p.s[k];
        q.L = items[k].L; q.Dinv = items[k].Dinv; q.B = items[k].B; q.T = items[k].npad / TILE; q.mode = (int)items[k].mode;
        q.flag_off = off; offs[k] = off;
        q.after_off = items[k].after >= 0 ? offs[items[k].after] : -1;
        q.after_T = items[k].after >= 0 ? items[items[k].after].npad / TILE : 0;
        q.epi = items[k].epi;
        off += q.T * p.nslabs;
        jobs += q.T * p.nslabs;
        p.job_end[k] = jobs;
    }
    for (int k = n; k < SOLVE_SEQ_MAX; ++k) p.job_end[k] = jobs;
    cudaMemsetAsync(sched.counter, 0, sizeof(unsigned), st);
    sched.epoch++;
    p.epoch = sched.epoch;
    int grid = 2 * sched.num_sms;
    if (grid > jobs) grid = jobs;
    solve_seq_kernel<<<grid, NTHREADS, SOLVE_SMEM, st>>>(p); ++g_kernel_launches;
}

void launch_solve(cudaStream_t st, Sched& sched, SolveMode mode, const double* L, int npad, double* B, long ldb,
                  int ncols, const SolveEpilogue& epi, int live_cols) {
    // Opt-in only (live_cols > 0): the simulation path must use ONE kernel whatever the batch size, so that
    // statistics stay bit-identical across batch and shard boundaries.
    if (mode != SOLVE_TRMM && live_cols > 0 && live_cols <= NARROW_MAX_COLS && live_cols <= ncols && npad > TILE &&
        !epi.dot0 && !epi.dot1 && !epi.xb) {
        launch_solve_narrow(st, sched, mode, L, npad, B, ldb, live_cols);
        return;
    }
    SolveParams p;
    p.L = L; p.T = npad / TILE; p.B = B; p.ldb = ldb; p.nslabs = ncols / SLAB;
    p.flags = sched.flags; p.counter = sched.counter; p.abort_word = sched.abort_word; p.epi = epi;
    cudaMemsetAsync(sched.counter, 0, sizeof(unsigned), st);
    sched.epoch++;
    p.epoch = sched.epoch;
    if (mode == SOLVE_TRMM) launch_mode<SOLVE_TRMM, false>(st, sched, p);
    else if (epi.xb) {
        if (mode == SOLVE_FWD) launch_mode<SOLVE_FWD, true>(st, sched, p);
        else launch_mode<SOLVE_BWD, true>(st, sched, p);
    } else {
        if (mode == SOLVE_FWD) launch_mode<SOLVE_FWD, false>(st, sched, p);
        else launch_mode<SOLVE_BWD, false>(st, sched, p);
    }
}

}  // namespace geordd

