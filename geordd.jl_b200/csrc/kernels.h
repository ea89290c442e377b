// Internal launcher interface between api.cu and the kernel translation units.
// Everything here works on DEVICE pointers; matrices are column-major, padded to multiples of
// TILE (128) in both dimensions unless noted.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <vector>

namespace geordd {

constexpr int TILE = 128;        // matrix tile edge: all device matrices are padded to multiples of TILE
constexpr int BK = 16;           // K-chunk of the mainloop

// Cholesky factors are stored TILE-BLOCKED AND PADDED ("TBP"): tile (I,J) of the npad x npad matrix is a
// contiguous 128 x 132 column-major image (column stride TLD = 132, the same padded image the mainloop keeps in
// shared memory), so a 16-column K-chunk of a tile is ONE contiguous 16.5 KB bulk copy.
constexpr int TLD = TILE + 4;
constexpr long TSZ = (long)TILE * TLD;
inline size_t factor_elems(int npad) { size_t T = npad / TILE; return T * T * (size_t)TSZ; }
#ifdef __CUDACC__
__host__ __device__
#endif
inline long tile_off(int I, int J, int T) { return ((long)J * T + I) * TSZ; }

// Standard-normal draws Z are stored SLAB-BLOCKED ("ZB"): for slab s (64 draws) and K-chunk c (16 rows) the block
// [n = 0..63][k = 0..15 (+4 pad)] is a contiguous 64 x 20 image -- exactly the B stage of the mainloop -- so the
// triangular multiply moves a chunk of Z with ONE bulk copy.  Element (row r, draw col):
//   ((col/64) * (npad/16) + r/16) * ZB_ELEMS + (col%64) * ZB_LD + r%16
// The simulation buffers that the solves work on in place (y - m, alpha of the three GPs) use the same storage
// ("XB" in the solve kernels): every K-chunk of a right-hand-side slab is then ONE bulk copy as well.
constexpr int ZB_LD = BK + 4;
constexpr int ZB_ELEMS = 64 * ZB_LD;
inline size_t zb_elems(int npad, long ncols) { return (size_t)(ncols / 64) * (npad / BK) * ZB_ELEMS; }
#ifdef __CUDACC__
__host__ __device__
#endif
inline long zb_off(long r, long col, int npad) {
    return ((col >> 6) * (npad / BK) + (r >> 4)) * ZB_ELEMS + (col & 63) * ZB_LD + (r & 15);
}

struct KernelSpecDev {
    int kind;            // 0 SE, 1 Mat12, 2 Mat32, 3 Mat52
    double ell, sigma2, const_sigma2;
};

// Shared scheduling scratch for the dataflow kernels (one per context / stream).
struct Sched {
    unsigned* counter;   // job counters (device), 16 words
    unsigned* flags;     // done-flags (device)
    size_t flags_cap;    // in words
    int* abort_word;     // device int: 0 ok, >0 potrf info, <0 watchdog
    unsigned epoch;      // host-side launch epoch (monotone)
    int2* jobs;          // device job table scratch
    size_t jobs_cap;
    std::vector<int> jobs_key;   // (tiles, prefilled tiles) per matrix of the batch the cached potrf job table was built for
    int jobs_count = 0;
    int num_sms;
};

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
extern std::atomic<long long> g_kernel_launches;   // (several host threads launch when a multi-device group is in use)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, and one process may hold contexts on
// several GPUs (geordd_ctx_create(out, device), the multi-device group of api_multi.cu): remember it per device, thread-safe.
template <class F>
inline void ensure_dyn_smem(F* fn, size_t bytes, std::atomic<unsigned long long>& done_mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done_mask.load(std::memory_order_acquire) & bit)) {
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        done_mask.fetch_or(bit, std::memory_order_release);
    }
}

// ---- cov.cu ---------------------------------------------------------------------------------------
// K(i,j) = k(|x1_i - x2_j|) (+ diag_add on i==j when sym).  X1: dim x n1, X2: dim x n2 (column =
// unit).  Output leading dimension ldk; rows/cols beyond n1/n2 up to the padded size (n1p, n2p) are
// written as 0 except a unit diagonal when sym (so the padded matrix stays SPD).  sym computes only the
// tiles on/below the diagonal; mirror additionally writes their transposes (full symmetric matrix).
void launch_cov(cudaStream_t st, const KernelSpecDev& spec, const double* X1, int n1, const double* X2, int n2,
                int dim, bool sym, bool mirror, double diag_add, double* K, long ldk, int n1p, int n2p,
                bool accumulate);

// ---- potrf.cu -------------------------------------------------------------------------------------
// L = chol(A) (lower), out of place.  A: npad x npad column-major (leading dimension ld), only its lower
// triangle is read.  L: TBP storage (factor_elems(npad) doubles); the strict upper triangle of the diagonal
// tiles is zeroed, upper tiles are written only by launch_mirror_factor.
// Result code is left in sched.abort_word (0 ok, k>0: pivot k non-positive, <0 watchdog).
// Lpre / npad_pre / ptiles (optional): the leading ptiles x ptiles tile block of the factor is already known -- it is the
// factor Lpre (TBP, npad_pre) of the leading principal submatrix (make_null: the pooled matrix starts with the treated
// block, whose factor the treated GP holds).  Those tiles are copied instead of recomputed; the left-looking tile jobs
// are deterministic, so the result is bit-identical to a full factorisation.
void launch_potrf(cudaStream_t st, Sched& sched, const double* A, double* L, long ld, int npad, const double* Lpre = nullptr,
                  int npad_pre = 0, int ptiles = 0);
// Several independent factorisations as ONE launch over a merged job list (their critical paths interleave).  The abort word
// of a failed batch holds (matrix index << 24) | LAPACK info.  Returns -1 if the batch does not fit the scheduling scratch.
constexpr int POTRF_MAX_BATCH = 64;
struct PotrfItem {
    const double* A; double* L; long ld; int npad;
    const double* Lpre; int npad_pre; int ptiles;
};
int launch_potrf_batch(cudaStream_t st, Sched& sched, const PotrfItem* items, int nm);
// dense (n x n column-major, ldd) lower triangle -> TBP factor with identity padding; and back (optionally transposed).
void launch_factor_pack(cudaStream_t st, const double* dense, long ldd, int n, double* L, int npad);
void launch_factor_unpack(cudaStream_t st, const double* L, int npad, double* dense, long ldd, int n, bool transpose);
// Inverses of the 32x32 diagonal blocks of every diagonal tile: Dinv holds (npad / 128) * 4608 doubles (tileops.cuh).
inline size_t diag_inverse_elems(int npad) { return (size_t)(npad / TILE) * 4 * 32 * 36; }
void launch_diag_inverse(cudaStream_t st, const double* L, int npad, double* Dinv);
// Mirror L' into the tiles above the diagonal (needed before any SOLVE_BWD against this factor).
void launch_mirror_factor(cudaStream_t st, double* L, int npad);
// out[0] = sum_i log L(i,i) for i < n   (L in TBP storage)
void launch_logdiag_sum(cudaStream_t st, const double* L, int npad, int n, double* out);

// ---- solve.cu -------------------------------------------------------------------------------------
enum SolveMode { SOLVE_FWD = 0, SOLVE_BWD = 1, SOLVE_TRMM = 2 };
// In-place triangular solve / multiply against lower factor L (npad x npad) for ncols right-hand
// sides (ncols multiple of 64): B (npad x ncols, ldb) is overwritten.
//   FWD:  B <- L^-1 B      BWD: B <- L^-T B      TRMM: out <- L * Z, Z given in ZB storage (ldb ignored)
// BWD requires launch_mirror_factor to have been applied to L.
struct SolveEpilogue {
    // xb: the right-hand sides B (FWD/BWD, in place), wmat, and the TRMM outputs out0/out1/outT/outC are in ZB
    // storage; their ld fields then hold the padded ROW COUNT of the buffer instead of a leading dimension.
    bool xb = false;
    int wrow0 = 0;   // xb only: wmat row r of this solve is row r + wrow0 of the wmat buffer
    // FWD/BWD: fused statistic partials, written per row block (deterministic; reduced later):
    //   dot0[rb*ncols + c] = sum_{r in block rb} (wmat(r,c) + (r < wrows ? wshift : 0)) * X(r,c)
    //   dot1[rb*ncols + c] = sum_{r in block rb} wvec[r] * X(r,c)
    //   dot_self: dot0 uses X itself as the weight (sum of squares of the solved block; wmat ignored)
    //   wlive: rows r >= wlive of this solve have NO row in the wmat buffer (the treated / control block of the pooled
    //   residuals ends at m_T / m_C while the solve runs over the padded row count): they are not loaded, weight 0 --
    //   the solution is exactly 0 there, so the partials are unchanged.
    const double* wmat = nullptr; long ldw = 0; double wshift = 0.0; int wrows = 0; int wlive = 0x7fffffff;
    bool dot_self = false;
    double* dot0 = nullptr;
    const double* wvec = nullptr;
    double* dot1 = nullptr;
    // TRMM only: Y = L*Z (+ add_const on logical rows, + add_T on rows < split) scattered to
    //   out0 / out1 : full copies of (Y - sub_0), (Y - sub_1)   (padded rows written as 0)
    //   outT        : rows [0, split)      of (Y - sub_T)
    //   outC        : rows [split, nrows)  of (Y - sub_C), re-based to row 0
    double* out0 = nullptr; long ld0 = 0;
    double* out1 = nullptr; long ld1 = 0;
    double* outT = nullptr; long ldT = 0;
    double* outC = nullptr; long ldC = 0;
    int split = 0; int nrows = 0;
    double add_const = 0.0, add_T = 0.0;
    double sub_0 = 0.0, sub_1 = 0.0, sub_T = 0.0, sub_C = 0.0;
};
// live_cols > 0: the caller states that only the first live_cols columns are non-zero / needed; with few of them
// (<= NARROW_MAX_COLS) and no fused epilogue, FWD/BWD go to the latency-optimised narrow kernel (solve_narrow.cu,
// 8-column slabs).  live_cols = 0: always the 64-column-slab kernel.
constexpr int NARROW_MAX_COLS = 128;
void launch_solve(cudaStream_t st, Sched& sched, SolveMode mode, const double* L, int npad, double* B, long ldb,
                  int ncols, const SolveEpilogue& epi, int live_cols = 0);
// Several XB solves over the same ncols right-hand sides as ONE persistent kernel (no per-kernel tails).  `after`
// (index of an earlier item, or -1): a job of this item may start once that item has finished the job's slab
// (backward after forward of the same buffer).
constexpr int SOLVE_SEQ_MAX = 6;
struct SolveSeqItem {
    SolveMode mode;
    const double* L;
    int npad;
    double* B;
    SolveEpilogue epi;
    int after = -1;
    const double* Dinv = nullptr;   // inverse diagonal blocks of L (launch_diag_inverse); nullptr: substitution in shared memory
};
void launch_solve_seq(cudaStream_t st, Sched& sched, const SolveSeqItem* items, int n, int ncols);
// lane (0..NARROW_LANES-1) selects a private job counter / flag region: narrow solves on different lanes may run concurrently
// on different streams (the alpha solves of the fits of a batch).
constexpr int NARROW_LANES = 4;
constexpr size_t NARROW_LANE_FLAGS = (size_t)1 << 16;
void launch_solve_narrow(cudaStream_t st, Sched& sched, SolveMode mode, const double* L, int npad, double* B, long ldb,
                         int live_cols, int lane = 0);

// ---- gemm.cu --------------------------------------------------------------------------------------
// C(M x N) = alpha * op(A) * op(B) + beta * C.  M multiple of 128, N multiple of 64, K multiple of 16.
//   a_mc: A(m,k) at A[k*lda+m] else A[m*lda+k];  b_nc: B(k,n) at B[k*ldb+n] else B[n*ldb+k]
// splitk > 1: partial tiles go to ws (splitk * ldc * N doubles) and are summed in fixed order.
// b_xb (requires a_mc, !b_nc): B is a K x N matrix in ZB storage (K = its padded row count, ldb ignored).
void launch_gemm(cudaStream_t st, bool a_mc, bool b_nc, int M, int N, int K, double alpha, const double* A, long lda,
                 const double* B, long ldb, double beta, double* C, long ldc, int splitk, double* ws, bool b_xb = false);

// ---- stats.cu -------------------------------------------------------------------------------------
// blocked = false: Z is npad x ndraws column-major (ldz);  blocked = true: ZB storage (ndraws multiple of 64).
void launch_philox_normals(cudaStream_t st, unsigned long long seed, unsigned long long draw0, int ndraws, int n,
                           double* Z, long ldz, int npad, bool blocked);
// dense (n x ncols_live, ldz) -> ZB storage for ncols (multiple of 64) draws, zero padded; and back
void launch_zb_pack(cudaStream_t st, const double* dense, long ldz, int n, long ncols_live, double* Zb, int npad, long ncols);
// out[col] = sum over rows of Z(r, col)^2 for ncols draws in ZB storage (padded rows are zero)
void launch_zb_colsumsq(cudaStream_t st, const double* Zb, int npad, long ncols, double* out);
void launch_zb_unpack(cudaStream_t st, const double* Zb, int npad, double* dense, long ldz, int n, long ncols_live, long col0);

// ---- geom.cu --------------------------------------------------------------------------------------
// Nearest point on a (multi-)polyline for n units.  X: 2 x n, V: 2 x nv vertices, brk[v] != 0 where vertex v starts a
// new line (no segment (v-1, v)).  Outputs: Xproj 2 x n, dist n.
void launch_projection(cudaStream_t st, const double* X, int n, const double* V, const unsigned char* brk, int nv,
                       double* Xproj, double* dist);
// inside[i] = 1 if point i lies strictly inside the region given by its rings (R: 2 x nr closed rings, brk marks ring starts)
void launch_in_region(cudaStream_t st, const double* P, int n, const double* R, const unsigned char* brk, int nr,
                      unsigned char* inside);

}  // namespace geordd
