// K2: FP64 Cholesky as ONE persistent dataflow kernel (replaces LAPACK dpotrf reached through
// GaussianProcesses.make_posdef! / PDMats.PDMat -- reference call sites src/gp_utils.jl:45-50,
// src/GPrealisationsCovars.jl:139, src/hypotest.jl:4).
//
// The matrix is cut into 128x128 tiles.  Job (i,j), i >= j, produces the FINAL tile L_ij:
//     acc = A_ij - sum_{k<j} L_ik L_jk^T        (DMMA mainloop, operands streamed from L2/HBM)
//     i == j :  L_jj = chol(acc)                (in shared memory)
//     i >  j :  L_ij = acc * L_jj^-T            (in shared memory)
// so every tile of L is written exactly once (left-looking per tile).  Jobs are handed out in
// column-major order by an atomic counter; a job only ever waits (acquire-polling per-tile flags)
// on jobs with a smaller index, which are already running on resident CTAs, hence no deadlock and no
// cooperative launch.  No host round trips, no per-panel launches, natural look-ahead.
//
// Round 2, two changes aimed at the critical path (one chain link per tile column: last update of the diagonal tile ->
// potrf_tile -> right solve of the sub-diagonal tile; 75 us per column in round 1, which bounds every n <= 8k):
//  * SUB-TILE FLAGS.  Besides its tile flag every tile has four block-column flags.  potrf_tile and trsm_right_tile publish
//    each 32-column block as soon as it is final (tileops.cuh), the right solve waits per block of the diagonal tile, and
//    the K loop waits per block (two 16-column chunks) of its operand tiles whenever the whole tile is not complete yet.
//    The three links of the chain now overlap as a pipeline with four stages per tile.  Arithmetic and its order are
//    unchanged (same jobs, same tile operations), so factors stay bit-identical to round 1's and to each other.
//  * BATCHES.  Several independent matrices (the regions of GPRealisations, src/GPrealisations.jl:10-19; the two halves of
//    a placebo split) are factored by ONE launch over a merged job list: their critical paths interleave on SMs that a
//    single small factorisation leaves idle.
#include <vector>
#include "kernels.h"
#include "mainloop.cuh"
#include "tileops.cuh"

namespace geordd {

// Flagged dataflow kernel: both operands are moved with generic-proxy block copies (see the stage-release note in mainloop.cuh).
using PotrfML = Mainloop<128, /*A_MC=*/true, /*B_NC=*/true, 4, BLK_GEN, BLK_GEN>;
constexpr int CHUNKS_PER_TILE = TILE / BK;
constexpr int SUBS = TILE / NB;   // block-column flags per tile

struct PotrfMat {
    const double* A;   // npad x npad column-major, lower triangle read
    double* L;         // TBP factor storage
    long ld;
    int T;             // tiles per dimension
    int flag_off;      // first flag word of this matrix: T*T tile flags, then T*T*SUBS block-column flags
};
struct PotrfBatch {
    int nm;
    PotrfMat m[POTRF_MAX_BATCH];
};

struct PotrfSrc {
    const double* L;
    long lda, ldb;
    int i, j, T;
    const unsigned* tflags;
    const unsigned* sflags;
    unsigned epoch;
    int* abort_word;
    bool tile_ok;   // (per warp) both operand tiles of the current K tile are complete
    // chunk c = columns [16c, 16c+16) of block row i (A) / j (B): 16 consecutive columns of one TBP tile
    __device__ __forceinline__ const double* a_ptr(int c) const {
        return L + tile_off(i, c / CHUNKS_PER_TILE, T) + (long)(c % CHUNKS_PER_TILE) * BK * TLD;
    }
    __device__ __forceinline__ const double* b_ptr(int c) const {
        return L + tile_off(j, c / CHUNKS_PER_TILE, T) + (long)(c % CHUNKS_PER_TILE) * BK * TLD;
    }
    __device__ __forceinline__ bool ready(int c) {   // warp level; chunks are requested in increasing order
        if (c & 1) return true;
        const int kt = c / CHUNKS_PER_TILE, b = (c % CHUNKS_PER_TILE) >> 1;
        const long ta = (long)i * T + kt, tb = (long)j * T + kt;
        if (b == 0) {   // whole tiles already complete (the common case away from the critical path)?
            bool ok = ld_acquire(tflags + ta) == epoch;
            if (i != j) ok = ok && ld_acquire(tflags + tb) == epoch;
            tile_ok = __all_sync(0xffffffffu, ok) != 0;
        }
        if (tile_ok) return true;
        return warp_wait_flags(sflags + ta * SUBS + b, i != j ? sflags + tb * SUBS + b : nullptr, epoch, abort_word);
    }
    // One poll instead of a wait (Mainloop::run defers the chunk and keeps computing).  May be asked repeatedly for the
    // same chunk; chunks are still requested in increasing order, so an odd chunk's block column has been acquired by
    // the successful call for the even chunk before it.
    __device__ __forceinline__ bool try_ready(int c) {
        if (c & 1) return true;
        const int kt = c / CHUNKS_PER_TILE, b = (c % CHUNKS_PER_TILE) >> 1;
        const long ta = (long)i * T + kt, tb = (long)j * T + kt;
        if (b == 0) {
            bool ok = ld_acquire(tflags + ta) == epoch;
            if (i != j) ok = ok && ld_acquire(tflags + tb) == epoch;
            tile_ok = __all_sync(0xffffffffu, ok) != 0;
        }
        if (tile_ok) return true;
        // every lane acquires the flags itself (as warp_wait_flags does after its wait) before it touches the tile
        bool ok = ld_acquire(sflags + ta * SUBS + b) == epoch;
        if (i != j) ok = ok && ld_acquire(sflags + tb * SUBS + b) == epoch;
        return __all_sync(0xffffffffu, ok) != 0;
    }
};

constexpr int POTRF_CS_ELEMS = TILE * SC;
constexpr int POTRF_LP_ELEMS = NB * SC;
constexpr size_t POTRF_SMEM = (size_t)(POTRF_CS_ELEMS + 2 * POTRF_LP_ELEMS + TILE + PotrfML::BAR_WORDS) * sizeof(double);
static_assert(PotrfML::SMEM_BYTES <= POTRF_CS_ELEMS * sizeof(double), "stage ring must fit under the C tile");
static_assert(SC == TLD, "the shared-memory C tile and the stored tile image share one layout");

__global__ void __launch_bounds__(NTHREADS, 1)
potrf_dataflow_kernel(const __grid_constant__ PotrfBatch pb, const int2* __restrict__ jobs, int njobs, unsigned* flags,
                      unsigned epoch, unsigned* counter, int* abort_word) {
    extern __shared__ __align__(16) double smem[];
    double* Cs = smem;                       // aliases the mainloop stage ring
    double* Lp = smem + POTRF_CS_ELEMS;
    double* Lp1 = Lp + POTRF_LP_ELEMS;   // second panel buffer of the right solve
    double* invd = Lp1 + POTRF_LP_ELEMS;
    __shared__ int sh_job;
    __shared__ int sh_info;
    const int tid = threadIdx.x;
    Ring ring;
    PotrfML::ring_init(ring, reinterpret_cast<uint64_t*>(invd + TILE), abort_word);

    while (true) {
        if (tid == 0) {
            int jb = (int)atomicAdd(counter, 1u);
            if (ld_volatile_int(abort_word) != 0) jb = njobs;
            sh_job = jb;
            sh_info = 0;
        }
        __syncthreads();
        const int job = sh_job;
        if (job >= njobs) break;
#ifdef GEORDD_POTRF_TRACE
        if (tid == 0) { g_ptrace_base = flags; g_ptrace_cur[blockIdx.x] = job; }
        __syncthreads();
        if (tid == 0) PTRACE_JOB(0);
#endif
        const int2 ij = jobs[job];
        const int mat = ij.x >> 16, i = ij.x & 0xffff, j = ij.y;
        const PotrfMat& M = pb.m[mat];
        const double* __restrict__ A = M.A;
        double* __restrict__ L = M.L;
        const long ld = M.ld;
        const int T = M.T;
        unsigned* tflags = flags + M.flag_off;
        unsigned* sflags = tflags + (long)T * T;

        PotrfML::Acc acc;
        {   // acc = -A_ij  (issued first so the loads overlap the pipeline prologue)
            const double* Aij = A + (long)j * TILE * ld + (long)i * TILE;
#pragma unroll
            for (int a = 0; a < PotrfML::MI; ++a)
#pragma unroll
                for (int b = 0; b < PotrfML::NJ; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        acc.v[a][b][e] = -ldcg(Aij + (long)PotrfML::acc_col(b, e) * ld + PotrfML::acc_row(a));
        }
        PotrfSrc src{L, 0, 0, i, j, T, tflags, sflags, epoch, abort_word, false};
        bool ok = PotrfML::run(smem, ring, j * CHUNKS_PER_TILE, src, acc);
        if (__syncthreads_or(!ok)) break;   // also: every warp is done reading the ring before Cs overwrites it
        if (tid == 0) PTRACE_JOB(1);
#pragma unroll
        for (int a = 0; a < PotrfML::MI; ++a)
#pragma unroll
            for (int b = 0; b < PotrfML::NJ; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    Cs[PotrfML::acc_col(b, e) * SC + PotrfML::acc_row(a)] = -acc.v[a][b][e];
        __syncthreads();
        if (tid == 0) PTRACE_JOB(2);

        double* Lij = L + tile_off(i, j, T);   // stored image: (m, n) at n * TLD + m, same as Cs (SC == TLD)
        unsigned* my_sub = sflags + ((long)i * T + j) * SUBS;
        if (i == j) {
            potrf_tile(Cs, invd, &sh_info, nullptr, Lij, my_sub, epoch);   // writes and publishes its block columns
            __syncthreads();
            if (sh_info != 0) {
                if (tid == 0) atomicCAS(abort_word, 0, (mat << 24) | (j * TILE + sh_info));
                break;
            }
        } else {
            const long dj = (long)j * T + j;
            if (!trsm_right_tile(Cs, Lp, Lp1, invd, L + tile_off(j, j, T), TLD, nullptr, tflags + dj, sflags + dj * SUBS, Lij,
                                 my_sub, epoch, abort_word))
                break;
        }
        // all block columns of the tile have been stored (and their stores ordered before a barrier) by the tile operation
        if (tid == 0) { st_release(tflags + (long)i * T + j, epoch); PTRACE_JOB(3); }
    }
}

#ifdef GEORDD_POTRF_TRACE
}  // namespace geordd
// debug builds only: the stamps of the last factorisation (jobs: 8192 x 16, flags: 65536 words), nanoseconds of %globaltimer
extern "C" __attribute__((visibility("default"))) int geordd_dbg_potrf_trace(unsigned long long* jobs_out, unsigned long long* flags_out) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(jobs_out, geordd::g_ptrace_job, sizeof(unsigned long long) * 8192 * 16) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(flags_out, geordd::g_ptrace_flag, sizeof(unsigned long long) * 65536) != cudaSuccess) return -1;
    return 0;
}
namespace geordd {
#endif

// Column-major job list of one matrix; tiles of a prefilled leading block are not jobs.
static void build_potrf_jobs(int mat, int T, int P, std::vector<int2>& out) {
    for (int j = 0; j < T; ++j)
        for (int i = j; i < T; ++i)
            if (!(i < P && j < P)) out.push_back(make_int2((mat << 16) | i, j));
}

// One CTA per tile (i, j), j <= i < P: copy the tile image from the known factor and publish its flag.
__global__ void __launch_bounds__(256) potrf_prefill_kernel(const double* __restrict__ Lpre, int Tpre, double* __restrict__ L, int T,
                                                            int P, unsigned* flags, unsigned epoch) {
    const int i = blockIdx.x % P, j = blockIdx.x / P;
    if (i < j) return;
    const double2* src = reinterpret_cast<const double2*>(Lpre + tile_off(i, j, Tpre));
    double2* dst = reinterpret_cast<double2*>(L + tile_off(i, j, T));
    for (int t = threadIdx.x; t < (int)(TSZ / 2); t += blockDim.x) dst[t] = src[t];
    if (threadIdx.x == 0) flags[(long)i * T + j] = epoch;
    // a consumer whose OTHER operand tile is still in flight waits per block column on both tiles
    if (threadIdx.x < SUBS) flags[(long)T * T + ((long)i * T + j) * SUBS + threadIdx.x] = epoch;
}

int launch_potrf_batch(cudaStream_t st, Sched& sched, const PotrfItem* items, int nm) {
    if (nm < 1 || nm > POTRF_MAX_BATCH) return -1;
    PotrfBatch pb;
    pb.nm = nm;
    std::vector<int> key;
    key.reserve(2 * nm);
    long off = 0;
    int Ps[POTRF_MAX_BATCH];
    for (int m = 0; m < nm; ++m) {
        const PotrfItem& it = items[m];
        const int T = it.npad / TILE;
        if (T > 0xffff) return -1;
        Ps[m] = (it.Lpre && it.ptiles > 0 && it.ptiles <= T && it.ptiles <= it.npad_pre / TILE) ? it.ptiles : 0;
        pb.m[m] = PotrfMat{it.A, it.L, it.ld, T, (int)off};
        off += (long)T * T * (1 + SUBS);
        key.push_back(T);
        key.push_back(Ps[m]);
    }
    if ((size_t)off > sched.flags_cap) return -1;
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(potrf_dataflow_kernel, POTRF_SMEM, attr_done);
    if (sched.jobs_key != key) {   // merged job table, cached per batch shape
        std::vector<std::vector<int2>> lists(nm);
        size_t total = 0;
        for (int m = 0; m < nm; ++m) {
            build_potrf_jobs(m, key[2 * m], key[2 * m + 1], lists[m]);
            total += lists[m].size();
        }
        if (total > sched.jobs_cap) return -1;
        // Proportional merge: every list keeps its own (column-major) order, so a job still waits only on jobs dealt before
        // it, and all factorisations advance at the same relative pace.
        std::vector<int2> merged;
        merged.reserve(total ? total : 1);
        std::vector<size_t> pos(nm, 0);
        for (size_t k = 0; k < total; ++k) {
            int best = -1;
            double bestkey = 0.0;
            for (int m = 0; m < nm; ++m) {
                if (pos[m] >= lists[m].size()) continue;
                const double kk = ((double)pos[m] + 0.5) / (double)lists[m].size();
                if (best < 0 || kk < bestkey) { best = m; bestkey = kk; }
            }
            merged.push_back(lists[best][pos[best]++]);
        }
        if (total) {
            cudaMemcpyAsync(sched.jobs, merged.data(), sizeof(int2) * total, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);   // `merged` is pageable: complete the copy before it goes away
        }
        sched.jobs_key = key;
        sched.jobs_count = (int)total;
    }
    const int njobs = sched.jobs_count;
    cudaMemsetAsync(sched.counter, 0, sizeof(unsigned), st);
    cudaMemsetAsync(sched.abort_word, 0, sizeof(int), st);
    sched.epoch++;
    for (int m = 0; m < nm; ++m) {
        if (Ps[m] > 0) {
            potrf_prefill_kernel<<<Ps[m] * Ps[m], 256, 0, st>>>(items[m].Lpre, items[m].npad_pre / TILE, items[m].L, pb.m[m].T, Ps[m],
                                                               sched.flags + pb.m[m].flag_off, sched.epoch); ++g_kernel_launches;
        }
    }
    if (njobs == 0) return 0;
    int grid = njobs < sched.num_sms ? njobs : sched.num_sms;
    potrf_dataflow_kernel<<<grid, NTHREADS, POTRF_SMEM, st>>>(pb, sched.jobs, njobs, sched.flags, sched.epoch, sched.counter,
                                                              sched.abort_word); ++g_kernel_launches;
    return 0;
}

void launch_potrf(cudaStream_t st, Sched& sched, const double* A, double* L, long ld, int npad, const double* Lpre, int npad_pre,
                  int ptiles) {
    PotrfItem it{A, L, ld, npad, Lpre, npad_pre, ptiles};
    launch_potrf_batch(st, sched, &it, 1);
}

// After the factorisation: store L' in the (otherwise unused) tiles above the diagonal, M(j-block, i-block) =
// L(i-block, j-block)' for i > j, so that backward solves read both L and L' with the row index contiguous.
// Diagonal tiles keep their explicit zeros above the diagonal (the TRMM path relies on them).
__global__ void __launch_bounds__(256) mirror_factor_kernel(double* __restrict__ L, int T) {
    __shared__ double t[64][65];
    // 64x64 quarters of the tiles strictly below the block diagonal
    const int S = 2 * T;
    int bi = blockIdx.x % S, bj = blockIdx.x / S;          // 64-block row / col
    const int I = bi / 2, J = bj / 2;
    if (I <= J) return;
    const double* src = L + tile_off(I, J, T) + (long)(bj % 2) * 64 * TLD + (bi % 2) * 64;
    double* dst = L + tile_off(J, I, T) + (long)(bi % 2) * 64 * TLD + (bj % 2) * 64;
    const int tx = threadIdx.x % 64, ty0 = threadIdx.x / 64;
    for (int c = ty0; c < 64; c += 4) t[c][tx] = src[(long)c * TLD + tx];
    __syncthreads();
    for (int c = ty0; c < 64; c += 4) dst[(long)c * TLD + tx] = t[tx][c];
}

// Inverses of the four 32x32 diagonal blocks of every diagonal tile (tileops.cuh: solve_tile_inv), one CTA per tile, one
// warp per block, one lane per column of the inverse (forward substitution on e_j in registers; 496 FMAs per lane).
__global__ void __launch_bounds__(128) diag_inverse_kernel(const double* __restrict__ L, int T, double* __restrict__ Dinv) {
    __shared__ double Ls[TILE / NB][NB][NB + 1];
    const int t = blockIdx.x, b = threadIdx.x >> 5, j = threadIdx.x & 31, o = NB * b;
    const double* Lt = L + tile_off(t, t, T);
#pragma unroll 4
    for (int c = 0; c < NB; ++c) Ls[b][j][c] = Lt[(long)(o + c) * TLD + o + j];   // row j of the block, coalesced along j
    __syncwarp();
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-Ls[b][i][k], x[k], s);
        x[i] = s / Ls[b][i][i];
    }
    double* out = Dinv + (long)t * DINV_TILE + b * DINV_BLK + j * LRS;   // column j of the inverse, rows contiguous
#pragma unroll
    for (int i = 0; i < NB; ++i) out[i] = (i >= j) ? x[i] : 0.0;
#pragma unroll
    for (int i = NB; i < LRS; ++i) out[i] = 0.0;
}

void launch_diag_inverse(cudaStream_t st, const double* L, int npad, double* Dinv) {
    diag_inverse_kernel<<<npad / TILE, 128, 0, st>>>(L, npad / TILE, Dinv); ++g_kernel_launches;
}

void launch_mirror_factor(cudaStream_t st, double* L, int npad) {
    const int T = npad / TILE;
    if (T < 2) return;
    mirror_factor_kernel<<<4 * T * T, 256, 0, st>>>(L, T); ++g_kernel_launches;
}

// dense lower triangle -> TBP factor ([L 0; 0 I] padding, explicit zeros above the diagonal of diagonal tiles)
__global__ void factor_pack_kernel(const double* __restrict__ dense, long ldd, int n, double* __restrict__ L, int T) {
    const int I = blockIdx.x % T, J = blockIdx.x / T;
    if (I < J) return;
    double* dst = L + tile_off(I, J, T);
    for (int idx = threadIdx.x; idx < TILE * TLD; idx += blockDim.x) {
        int c = idx / TLD, r = idx % TLD;
        int gr = I * TILE + r, gc = J * TILE + c;
        double v = 0.0;
        if (r < TILE && gr >= gc) {
            if (gr < n && gc < n) v = dense[(long)gc * ldd + gr];
            else if (gr == gc) v = 1.0;
        }
        dst[idx] = v;
    }
}
void launch_factor_pack(cudaStream_t st, const double* dense, long ldd, int n, double* L, int npad) {
    const int T = npad / TILE;
    factor_pack_kernel<<<T * T, 256, 0, st>>>(dense, ldd, n, L, T); ++g_kernel_launches;
}

// TBP lower factor -> dense n x n (strict upper zero); transpose = true writes U = L' (upper)
__global__ void factor_unpack_kernel(const double* __restrict__ L, int T, double* __restrict__ dense, long ldd, int n,
                                     int transpose) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    int r = (int)(idx % n), c = (int)(idx / n);          // destination element (r, c), coalesced along r
    int lr = transpose ? c : r, lc = transpose ? r : c;  // source element of L
    double v = 0.0;
    if (lr >= lc) v = L[tile_off(lr / TILE, lc / TILE, T) + (long)(lc % TILE) * TLD + lr % TILE];
    dense[(long)c * ldd + r] = v;
}
void launch_factor_unpack(cudaStream_t st, const double* L, int npad, double* dense, long ldd, int n, bool transpose) {
    long tot = (long)n * n;
    factor_unpack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L, npad / TILE, dense, ldd, n, transpose ? 1 : 0); ++g_kernel_launches;
}

__global__ void logdiag_sum_kernel(const double* __restrict__ L, int T, int n, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        s += log(L[tile_off(i / TILE, i / TILE, T) + (long)(i % TILE) * TLD + i % TILE]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = s;
    }
}

void launch_logdiag_sum(cudaStream_t st, const double* L, int npad, int n, double* out) {
    logdiag_sum_kernel<<<1, 1024, 0, st>>>(L, npad / TILE, n, out); ++g_kernel_launches;
}

}  // namespace geordd
