// K2: FP64 Cholesky as ONE persistent dataflow kernel (replaces LAPACK dpotrf reached through
// GaussianProcesses.make_posdef! / PDMats.PDMat -- reference call sites src/gp_utils.jl:45-50,
// src/GPrealisationsCovars.jl:139, src/hypotest.jl:4).
//
// The matrix is cut into 128x128 tiles.  Job (i,j), i >= j, produces the FINAL tile L_ij:
//     acc = A_ij - sum_{k<j} L_ik L_jk^T        (DMMA mainloop, operands streamed from L2/HBM)
//     i == j :  L_jj = chol(acc)                (in shared memory)
//     i >  j :  L_ij = acc * L_jj^-T            (in shared memory)
// so every tile of L is written exactly once (left-looking per tile).  Jobs are handed out in
// column-major order by an atomic counter; a job only ever waits (acquire-polling a per-tile flag)
// on jobs with a smaller index, which are already running on resident CTAs, hence no deadlock and no
// cooperative launch.  No host round trips, no per-panel launches, natural look-ahead.
#include "kernels.h"
#include "mainloop.cuh"
#include "tileops.cuh"

namespace geordd {

// Flagged dataflow kernel: both operands are moved with generic-proxy block copies (see the stage-release note in mainloop.cuh).
using PotrfML = Mainloop<128, /*A_MC=*/true, /*B_NC=*/true, 4, BLK_GEN, BLK_GEN>;
constexpr int CHUNKS_PER_TILE = TILE / BK;

struct PotrfSrc {
    const double* L;
    long lda, ldb;
    int i, j, T;
    const unsigned* flags;
    unsigned epoch;
    int* abort_word;
    // chunk c = columns [16c, 16c+16) of block row i (A) / j (B): 16 consecutive columns of one TBP tile
    __device__ __forceinline__ const double* a_ptr(int c) const {
        return L + tile_off(i, c / CHUNKS_PER_TILE, T) + (long)(c % CHUNKS_PER_TILE) * BK * TLD;
    }
    __device__ __forceinline__ const double* b_ptr(int c) const {
        return L + tile_off(j, c / CHUNKS_PER_TILE, T) + (long)(c % CHUNKS_PER_TILE) * BK * TLD;
    }
    __device__ __forceinline__ bool ready(int c) const {   // warp level
        if (c % CHUNKS_PER_TILE != 0) return true;
        int kt = c / CHUNKS_PER_TILE;
        return warp_wait_flags(flags + (long)i * T + kt, i != j ? flags + (long)j * T + kt : nullptr, epoch, abort_word);
    }
};

constexpr int POTRF_CS_ELEMS = TILE * SC;
constexpr int POTRF_LP_ELEMS = NB * SC;
constexpr size_t POTRF_SMEM = (size_t)(POTRF_CS_ELEMS + 2 * POTRF_LP_ELEMS + TILE + PotrfML::BAR_WORDS) * sizeof(double);
static_assert(PotrfML::SMEM_BYTES <= POTRF_CS_ELEMS * sizeof(double), "stage ring must fit under the C tile");

__global__ void __launch_bounds__(NTHREADS, 1)
potrf_dataflow_kernel(const double* __restrict__ A, double* __restrict__ L, long ld, int T, const int2* __restrict__ jobs,
                      int njobs, unsigned* flags, unsigned epoch, unsigned* counter, int* abort_word) {
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
        const int2 ij = jobs[job];
        const int i = ij.x, j = ij.y;

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
        PotrfSrc src{L, 0, 0, i, j, T, flags, epoch, abort_word};
        bool ok = PotrfML::run(smem, ring, j * CHUNKS_PER_TILE, src, acc);
        if (__syncthreads_or(!ok)) break;   // also: every warp is done reading the ring before Cs overwrites it
#pragma unroll
        for (int a = 0; a < PotrfML::MI; ++a)
#pragma unroll
            for (int b = 0; b < PotrfML::NJ; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    Cs[PotrfML::acc_col(b, e) * SC + PotrfML::acc_row(a)] = -acc.v[a][b][e];
        __syncthreads();

        double* Lij = L + tile_off(i, j, T);   // stored image: (m, n) at n * TLD + m, same as Cs (SC == TLD)
        if (i == j) {
            potrf_tile(Cs, invd, &sh_info);
            __syncthreads();
            if (sh_info != 0) {
                if (tid == 0) atomicCAS(abort_word, 0, j * TILE + sh_info);
                break;
            }
            for (int idx = tid; idx < TILE * TLD; idx += NTHREADS) {
                int n = idx / TLD, m = idx % TLD;
                Lij[idx] = (m < TILE && m >= n) ? Cs[n * SC + m] : 0.0;
            }
        } else {
            int okf = 1;
            if (tid == 0) okf = wait_flag(flags + (long)j * T + j, epoch, abort_word);
            if (!__syncthreads_and(okf)) break;
            trsm_right_tile(Cs, Lp, Lp1, invd, L + tile_off(j, j, T), TLD);
            for (int idx = tid; idx < TILE * TLD; idx += NTHREADS) {
                int m = idx % TLD;
                Lij[idx] = (m < TILE) ? Cs[idx] : 0.0;
            }
        }
        __threadfence();
        fence_proxy_async();   // generic smem accesses of the finalize step before the next job's bulk copies
        __syncthreads();
        if (tid == 0) st_release(flags + (long)i * T + j, epoch);
    }
}

static int build_potrf_jobs(int T, int P, int2* host) {
    int p = 0;
    for (int j = 0; j < T; ++j)
        for (int i = j; i < T; ++i)
            if (!(i < P && j < P)) host[p++] = make_int2(i, j);   // tiles of the prefilled leading block are not jobs
    return p;
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
}

void launch_potrf(cudaStream_t st, Sched& sched, const double* A, double* L, long ld, int npad, const double* Lpre, int npad_pre,
                  int ptiles) {
    const int T = npad / TILE;
    const int P = (Lpre && ptiles > 0 && ptiles <= T && ptiles <= npad_pre / TILE) ? ptiles : 0;
    const int njobs = T * (T + 1) / 2 - P * (P + 1) / 2;
    static std::atomic<unsigned long long> attr_done{0};
    ensure_dyn_smem(potrf_dataflow_kernel, POTRF_SMEM, attr_done);
    if (sched.jobs_T != T || sched.jobs_P != P) {   // job table cached per (tile count, prefilled tiles)
        int2* hjobs = (int2*)malloc(sizeof(int2) * (njobs > 0 ? njobs : 1));
        build_potrf_jobs(T, P, hjobs);
        cudaMemcpyAsync(sched.jobs, hjobs, sizeof(int2) * njobs, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);   // hjobs is pageable: complete the copy before free
        free(hjobs);
        sched.jobs_T = T;
        sched.jobs_P = P;
    }
    cudaMemsetAsync(sched.counter, 0, sizeof(unsigned), st);
    cudaMemsetAsync(sched.abort_word, 0, sizeof(int), st);
    sched.epoch++;
    if (P > 0) {
        potrf_prefill_kernel<<<P * P, 256, 0, st>>>(Lpre, npad_pre / TILE, L, T, P, sched.flags, sched.epoch); ++g_kernel_launches;
    }
    if (njobs == 0) return;
    int grid = njobs < sched.num_sms ? njobs : sched.num_sms;
    potrf_dataflow_kernel<<<grid, NTHREADS, POTRF_SMEM, st>>>(A, L, ld, T, sched.jobs, njobs, sched.flags, sched.epoch,
                                                              sched.counter, sched.abort_word); ++g_kernel_launches;
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
