// Small kernels around the dense core: counter-based normal draws, statistic finishing with
// integer exceedance tallies, the pivoted-LU solve used for the OBSERVED statistics (Julia's generic
// `Sigma \ mu` on a Matrix is lu(), src/hypotest_chi2.jl:9, src/average_treatment_effect.jl:6-7),
// and a handful of vector/matrix utilities.
#include "kernels.h"
#include "smallops.h"

namespace geordd {

// ---------------------------------------------------------------------------------------------------
// Philox-4x32-10 -> Box-Muller standard normals.  Element pair (2j, 2j+1) of draw d comes from
// counter (j, 0, d_lo, d_hi), key (seed_lo, seed_hi); the oracle implements the same generator
// (oracle/geordd_oracle.py: philox_normals).  One randn(n) per draw, consumed in order, exactly as
// rand(MvNormal) does in src/hypotest.jl:13-18.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0,
                                             unsigned k1) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__global__ void philox_normals_kernel(unsigned long long seed, unsigned long long draw0, int ndraws, int n,
                                      double* __restrict__ Z, long ldz, int npad, int blocked) {
    const int npairp = npad / 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)npairp * ndraws) return;
    int j = (int)(idx % npairp);
    int s = (int)(idx / npairp);
    double z0 = 0.0, z1 = 0.0;
    if (2 * j < n) {
        unsigned long long d = draw0 + (unsigned long long)s;
        unsigned c0 = (unsigned)j, c1 = 0u, c2 = (unsigned)(d & 0xffffffffull), c3 = (unsigned)(d >> 32);
        unsigned k0 = (unsigned)(seed & 0xffffffffull), k1 = (unsigned)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c0, c1, c2, c3, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        unsigned long long w1 = ((unsigned long long)c0 << 32) | c1;
        unsigned long long w2 = ((unsigned long long)c2 << 32) | c3;
        double u1 = ((double)(w1 >> 11) + 0.5) * 1.1102230246251565e-16;   // 2^-53
        double u2 = ((double)(w2 >> 11) + 0.5) * 1.1102230246251565e-16;
        double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        z0 = rad * cs;
        z1 = (2 * j + 1 < n) ? rad * sn : 0.0;
    }
    const int r = 2 * j;
    long off = blocked ? ((long)(s / 64) * (npad / BK) + r / BK) * ZB_ELEMS + (long)(s % 64) * ZB_LD + r % BK
                       : (long)s * ldz + r;
    *reinterpret_cast<double2*>(Z + off) = make_double2(z0, z1);
}

void launch_philox_normals(cudaStream_t st, unsigned long long seed, unsigned long long draw0, int ndraws, int n,
                           double* Z, long ldz, int npad, bool blocked) {
    long tot = (long)(npad / 2) * ndraws;
    philox_normals_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(seed, draw0, ndraws, n, Z, ldz, npad, blocked ? 1 : 0); ++g_kernel_launches;
}

__global__ void zb_pack_kernel(const double* __restrict__ dense, long ldz, int n, long live, double* __restrict__ Zb,
                               int npad, long ncols) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)npad * ncols) return;
    int r = (int)(idx % npad);
    long col = idx / npad;
    double v = (r < n && col < live) ? dense[col * ldz + r] : 0.0;
    Zb[((col / 64) * (npad / BK) + r / BK) * ZB_ELEMS + (col % 64) * ZB_LD + r % BK] = v;
}
// one warp per draw: lanes stride over the K-chunks of the column, fixed-order shuffle reduction (deterministic)
__global__ void zb_colsumsq_kernel(const double* __restrict__ Zb, int npad, long ncols, double* __restrict__ out) {
    long col = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (col >= ncols) return;
    const int lane = threadIdx.x & 31;
    const int nch = npad / BK;
    const double* base = Zb + ((col >> 6) * nch) * (long)ZB_ELEMS + (col & 63) * ZB_LD;
    double s = 0.0;
    for (int ch = lane; ch < nch; ch += 32) {
        const double2* p = reinterpret_cast<const double2*>(base + (long)ch * ZB_ELEMS);
#pragma unroll
        for (int k = 0; k < BK / 2; ++k) {
            double2 v = p[k];
            s = fma(v.x, v.x, s);
            s = fma(v.y, v.y, s);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[col] = s;
}
void launch_zb_colsumsq(cudaStream_t st, const double* Zb, int npad, long ncols, double* out) {
    zb_colsumsq_kernel<<<(unsigned)((ncols + 7) / 8), 256, 0, st>>>(Zb, npad, ncols, out); ++g_kernel_launches;
}

__global__ void zb_unpack_kernel(const double* __restrict__ Zb, int npad, double* __restrict__ dense, long ldz, int n, long live,
                                 long col0) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * live) return;
    int r = (int)(idx % n);
    long c = idx / n;
    dense[c * ldz + r] = Zb[zb_off(r, col0 + c, npad)];
}
void launch_zb_unpack(cudaStream_t st, const double* Zb, int npad, double* dense, long ldz, int n, long ncols_live, long col0) {
    long tot = (long)n * ncols_live;
    if (tot <= 0) return;
    zb_unpack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Zb, npad, dense, ldz, n, ncols_live, col0); ++g_kernel_launches;
}
void launch_zb_pack(cudaStream_t st, const double* dense, long ldz, int n, long ncols_live, double* Zb, int npad, long ncols) {
    long tot = (long)npad * ncols;
    zb_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dense, ldz, n, ncols_live, Zb, npad, ncols); ++g_kernel_launches;
}

// ---------------------------------------------------------------------------------------------------
// Statistic finishing.  Partials come from the solve epilogue as [nblocks x ncols] arrays and are
// summed in block order (deterministic).  Tallies are integer atomics (order independent).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sum_partials(const double* part, int nblocks, long ncols, long c) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += part[(long)b * ncols + c];
    return s;
}

// Smallest distance between any simulated statistic and the observed threshold, kept as the bit pattern of a non-negative
// double (ordered like the integers): the caller learns how far the exact exceedance count is from flipping
// (SURVEY §5 "minimum margin") without fetching the statistics.
__device__ __forceinline__ void margin_min(bool live, double dist, unsigned long long* slot) {
    unsigned long long b = live && dist == dist ? (unsigned long long)__double_as_longlong(fabs(dist)) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long ob = __shfl_xor_sync(0xffffffffu, b, o);
        b = ob < b ? ob : b;
    }
    if ((threadIdx.x & 31) == 0 && b != ~0ull) atomicMin(slot, b);
}

__device__ __forceinline__ void block_tally(bool hit, unsigned long long* counter) {
    unsigned ballot = __ballot_sync(0xffffffffu, hit);
    __shared__ int sh_cnt;
    if (threadIdx.x == 0) sh_cnt = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&sh_cnt, __popc(ballot));
    __syncthreads();
    if (threadIdx.x == 0 && sh_cnt) atomicAdd(counter, (unsigned long long)sh_cnt);
}

// mll_X = -0.5 * dot_X + c_X with c_X = -0.5 logdet_X - n_X/2 log(2 pi)   (src/gp_utils.jl:19-22)
__global__ void finish_mll_kernel(FinishMll f) {
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    double dist = 0.0;
    if (c < f.ndraws) {
        double dN = sum_partials(f.partN, f.nbN, f.ncols, c);
        double dT = sum_partials(f.partT, f.nbT, f.ncols, c);
        double dC = sum_partials(f.partC, f.nbC, f.ncols, c);
        double mllN = -dN / 2.0 - f.halflogdetN - f.constN;
        double mllT = -dT / 2.0 - f.halflogdetT - f.constT;
        double mllC = -dC / 2.0 - f.halflogdetC - f.constC;
        double alt = mllT + mllC;
        if (f.out_null) f.out_null[f.out_off + c] = mllN;
        if (f.out_alt) f.out_alt[f.out_off + c] = alt;
        hit = (alt - mllN) > f.dmll_obs;     // mean(dmll_sim .> dmll_obs), src/hypotest_mll.jl:42-44
        dist = (alt - mllN) - f.dmll_obs;
    }
    if (f.margin) margin_min(c < f.ndraws, dist, f.margin);
    block_tally(hit, f.tally);
}
void launch_finish_mll(cudaStream_t st, const FinishMll& f) {
    finish_mll_kernel<<<(unsigned)((f.ndraws + 255) / 256), 256, 0, st>>>(f); ++g_kernel_launches;
}

// chi2 = dot(mu, Sigma \ mu) (src/hypotest_chi2.jl:16); invvar: tau = sum(Sigma\mu)/denom,
// V = 1/denom, p = 2 min(cdf(N(tau, sqrt V), 0), ccdf(...)) (src/average_treatment_effect.jl:3-11,
// src/hypotest_invvar.jl:16-18; Normal cdf via erfc as StatsFuns does).
__global__ void finish_cliff_kernel(FinishCliff f) {
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit_chi = false, hit_inv = false;
    double dist_chi = 0.0, dist_inv = 0.0;
    if (c < f.ndraws) {
        if (f.part_chi) {
            double chi = sum_partials(f.part_chi, f.nb, f.ncols, c);
            if (f.out_chi) f.out_chi[f.out_off + c] = chi;
            hit_chi = chi > f.chi_obs;
            dist_chi = chi - f.chi_obs;
        }
        if (f.part_num) {
            double num = sum_partials(f.part_num, f.nb, f.ncols, c);
            double tau = num / f.denom;
            double sd = sqrt(1.0 / f.denom);
            double z = (0.0 - tau) / sd;
            double cdf = erfc(-z * 0.7071067811865476) / 2.0;
            double ccdf = erfc(z * 0.7071067811865476) / 2.0;
            double p = 2.0 * fmin(cdf, ccdf);
            if (f.out_pval) f.out_pval[f.out_off + c] = p;
            if (f.out_num) f.out_num[f.out_off + c] = num;
            hit_inv = p < f.pval_obs;
            dist_inv = p - f.pval_obs;
        }
    }
    if (f.part_chi && f.margin_chi) margin_min(c < f.ndraws, dist_chi, f.margin_chi);
    if (f.part_num && f.margin_inv) margin_min(c < f.ndraws, dist_inv, f.margin_inv);
    if (f.part_chi) block_tally(hit_chi, f.tally_chi);
    if (f.part_num) block_tally(hit_inv, f.tally_inv);
}
void launch_finish_cliff(cudaStream_t st, const FinishCliff& f) {
    finish_cliff_kernel<<<(unsigned)((f.ndraws + 255) / 256), 256, 0, st>>>(f); ++g_kernel_launches;
}

// out[s] = mean(ref .> x[s]) (greater != 0) or mean(ref .< x[s])   (src/power.jl:31,35,40)
__global__ void frac_compare_kernel(const double* ref, long nref, const double* x, long n, int greater, double* out) {
    long s = blockIdx.x;
    if (s >= n) return;
    double xv = x[s];
    int cnt = 0;
    for (long i = threadIdx.x; i < nref; i += blockDim.x) cnt += greater ? (ref[i] > xv) : (ref[i] < xv);
    __shared__ int sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&sh, cnt);
    __syncthreads();
    if (threadIdx.x == 0) out[s] = (double)sh / (double)nref;
}
void launch_frac_compare(cudaStream_t st, const double* ref, long nref, const double* x, long n, bool greater,
                         double* out) {
    if (n > 0) frac_compare_kernel<<<(unsigned)n, 256, 0, st>>>(ref, nref, x, n, greater ? 1 : 0, out); ++g_kernel_launches;
}

// ---------------------------------------------------------------------------------------------------
// Pivoted LU solve, single CTA (n up to a few thousand; nb x nb cliff-face systems).  A is
// overwritten by its LU factors, B (n x nrhs) by the solution.  Partial pivoting picks the first
// row of maximum modulus (idamax), scaling multiplies by the reciprocal pivot (dgetf2).
// ---------------------------------------------------------------------------------------------------
// Multi-CTA form (Asrc != nullptr): every CTA copies the same n x n matrix into its own scratch block, factors it
// redundantly (identical pivots, identical factors: the elimination is deterministic) and solves its own chunk of
// `cols_per_cta` right-hand sides -- lu(Sigma) \ mu_b for thousands of null draws (sim_invvar_calib!, src/hypotest_invvar.jl:138-147).
__global__ void __launch_bounds__(1024) lu_solve_kernel(double* A, long lda, int n, double* B, long ldb, int nrhs,
                                                        int* info, const double* Asrc, long ldsrc, int cols_per_cta) {
    if (Asrc) {
        A += (long)blockIdx.x * lda * n;
        for (long idx = threadIdx.x; idx < (long)n * n; idx += blockDim.x) A[(idx / n) * lda + idx % n] = Asrc[(idx / n) * ldsrc + idx % n];
        B += (long)blockIdx.x * cols_per_cta * ldb;
        nrhs = min(cols_per_cta, nrhs - (int)blockIdx.x * cols_per_cta);
        if (blockIdx.x != 0) info = nullptr;
        __syncthreads();
    }
    __shared__ double sval[32];
    __shared__ int sidx[32];
    __shared__ int spiv;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int bi = n;
        for (int i = k + tid; i < n; i += nt) {
            double v = fabs(A[(long)k * lda + i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, best, o);
            int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { sval[tid >> 5] = best; sidx[tid >> 5] = bi; }
        __syncthreads();
        if (tid < 32) {
            best = tid < (nt >> 5) ? sval[tid] : -1.0;
            bi = tid < (nt >> 5) ? sidx[tid] : n;
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_down_sync(0xffffffffu, best, o);
                int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) {
                spiv = bi;
                if (info && !(best > 0.0) && *info == 0) *info = k + 1;
            }
        }
        __syncthreads();
        const int piv = spiv;
        if (piv != k && piv < n) {
            for (int j = tid; j < n; j += nt) {
                double t = A[(long)j * lda + k];
                A[(long)j * lda + k] = A[(long)j * lda + piv];
                A[(long)j * lda + piv] = t;
            }
            for (int j = tid; j < nrhs; j += nt) {
                double t = B[(long)j * ldb + k];
                B[(long)j * ldb + k] = B[(long)j * ldb + piv];
                B[(long)j * ldb + piv] = t;
            }
        }
        __syncthreads();
        const double rp = 1.0 / A[(long)k * lda + k];
        for (int i = k + 1 + tid; i < n; i += nt) A[(long)k * lda + i] *= rp;
        __syncthreads();
        const int rem = n - k - 1;
        if (rem > 0) {
            const long tot = (long)rem * (rem + nrhs);
            for (long idx = tid; idx < tot; idx += nt) {
                int i = k + 1 + (int)(idx % rem);
                int jj = (int)(idx / rem);
                double lik = A[(long)k * lda + i];
                if (jj < rem) {
                    int j = k + 1 + jj;
                    A[(long)j * lda + i] -= lik * A[(long)j * lda + k];
                } else {
                    int j = jj - rem;
                    B[(long)j * ldb + i] -= lik * B[(long)j * ldb + k];
                }
            }
        }
        __syncthreads();
    }
    // back substitution with U, column oriented
    for (int k = n - 1; k >= 0; --k) {
        const double ukk = A[(long)k * lda + k];
        for (int j = tid; j < nrhs; j += nt) B[(long)j * ldb + k] /= ukk;
        __syncthreads();
        const long tot = (long)k * nrhs;
        for (long idx = tid; idx < tot; idx += nt) {
            int i = (int)(idx % k), j = (int)(idx / k);
            B[(long)j * ldb + i] -= A[(long)k * lda + i] * B[(long)j * ldb + k];
        }
        __syncthreads();
    }
}
// ---------------------------------------------------------------------------------------------------
// The same elimination, blocked for locality (n <= 640: the nb x nb cliff-face systems and the p x p precision of
// postmean_beta).  The unblocked kernel above pays five CTA barriers and several L2 round trips per pivot (8.8 us per
// pivot: 1.8 ms for the 200 x 200 system of one observed chi2 statistic).  Here a panel of NBL columns is held in REGISTERS
// while its pivots are found -- thread t owns row t (and t + 512) of the panel; per pivot three barriers, two row images
// exchanged through shared memory, no global access.  The columns to its right and the right-hand sides are then updated
// warp by warp, JB columns at a time, a column held in registers across the lanes (row r of the panel-relative segment
// in lane r % 32, slot r / 32) against the panel in shared memory -- the panel's row exchanges are applied while the
// column is loaded (`src`: the permutation the NBL exchanges compose to).  Every entry still receives exactly the
// operations of the unblocked kernel in the same order (a = fma(-l_ic, u_cj, a) for c ascending; reciprocal-pivot
// scaling; true division in the back substitution), so factors, pivots and solutions are bit-identical to it -- tested
// (geordd_dbg_set_lu_blocked).  Exchanges are not applied to the L part left of a panel: it is never read again, the
// right-hand sides are eliminated together with the matrix.
// ---------------------------------------------------------------------------------------------------
template <int MQ, int NBL, int JB>
__global__ void __launch_bounds__(512) lu_blocked_kernel(double* A, long lda, int n, double* B, long ldb, int nrhs, int* info,
                                                         const double* Asrc, long ldsrc, int cols_per_cta) {
    constexpr int LDP = 32 * MQ + 1, NT = 512, NW = NT / 32, R = (32 * MQ + NT - 1) / NT;
    extern __shared__ double lu_sm[];
    double* P = lu_sm;                               // NBL panel columns, LDP apart
    int* src = (int*)(P + (size_t)NBL * LDP);        // row r of the permuted segment comes from row src[r]
    __shared__ double sval[NW];
    __shared__ double rowbuf[2 * NBL];               // [0, NBL): old row c, [NBL, 2 NBL): old row piv = the pivot row
    __shared__ int sidx[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (Asrc) {
        A += (long)blockIdx.x * lda * n;
        for (long idx = tid; idx < (long)n * n; idx += NT) A[(idx / n) * lda + idx % n] = Asrc[(idx / n) * ldsrc + idx % n];
        B += (long)blockIdx.x * cols_per_cta * ldb;
        nrhs = min(cols_per_cta, nrhs - (int)blockIdx.x * cols_per_cta);
        if (blockIdx.x != 0) info = nullptr;
        __syncthreads();
    }
    for (int k0 = 0; k0 < n; k0 += NBL) {
        const int w = min(NBL, n - k0), m = n - k0;
        // ---- pivots of the panel: thread t holds rows t (+ 512) of the panel in registers ----
        double r[R][NBL];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int row = tid + NT * rr;
#pragma unroll
            for (int cc = 0; cc < NBL; ++cc) r[rr][cc] = (row < m && cc < w) ? A[(long)(k0 + cc) * lda + k0 + row] : 0.0;
        }
        for (int i = tid; i < m; i += NT) src[i] = i;
#pragma unroll
        for (int c = 0; c < NBL; ++c) {
            if (c < w) {
                double best = -1.0;
                int bi = m;
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int row = tid + NT * rr;
                    const double v = fabs(r[rr][c]);
                    if (row >= c && row < m && v > best) { best = v; bi = row; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, best, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) { sval[warp] = best; sidx[warp] = bi; }
                __syncthreads();
                // every warp finishes the reduction itself (maximum with first-index tie-break is order independent): no
                // second barrier to publish the pivot
                best = lane < NW ? sval[lane] : -1.0;
                bi = lane < NW ? sidx[lane] : m;
#pragma unroll
                for (int o = NW / 2; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                best = __shfl_sync(0xffffffffu, best, 0);
                bi = __shfl_sync(0xffffffffu, bi, 0);
                if (tid == 0 && info && !(best > 0.0) && *info == 0) *info = k0 + c + 1;
                const int piv = bi < m ? bi : c;
                // rows c and piv change owners through shared memory; rowbuf[1] (the old row piv) is the pivot row
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int row = tid + NT * rr;
                    if (row == c || row == piv) {
#pragma unroll
                        for (int cc = 0; cc < NBL; ++cc) {
                            if (row == c) rowbuf[cc] = r[rr][cc];
                            if (row == piv) rowbuf[NBL + cc] = r[rr][cc];
                        }
                    }
                }
                if (tid == NT - 1 && piv != c) {
                    const int t = src[c];
                    src[c] = src[piv];
                    src[piv] = t;
                }
                __syncthreads();
                const double rp = 1.0 / rowbuf[NBL + c];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const int row = tid + NT * rr;
                    if (row == c || row == piv) {
#pragma unroll
                        for (int cc = 0; cc < NBL; ++cc) r[rr][cc] = row == c ? rowbuf[NBL + cc] : rowbuf[cc];
                    }
                    if (row > c && row < m) {
                        const double l = r[rr][c] * rp;
                        r[rr][c] = l;
#pragma unroll
                        for (int jj = c + 1; jj < NBL; ++jj) r[rr][jj] = fma(-l, rowbuf[NBL + jj], r[rr][jj]);
                    }
                }
            }
        }
        // the factored panel: to shared memory for the elimination below, and back to A
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int row = tid + NT * rr;
            if (row < m) {
#pragma unroll
                for (int cc = 0; cc < NBL; ++cc) {
                    if (cc < w) {
                        P[cc * LDP + row] = r[rr][cc];
                        A[(long)(k0 + cc) * lda + k0 + row] = r[rr][cc];
                    }
                }
            }
        }
        __syncthreads();
        // ---- columns right of the panel and the right-hand sides: exchange rows, eliminate with the panel ----
        const int nright = n - k0 - w, total = nright + nrhs;
        for (int g = warp; g * JB < total; g += NW) {
            double* cp[JB];
#pragma unroll
            for (int jb = 0; jb < JB; ++jb) {
                const int jj = g * JB + jb;
                cp[jb] = jj >= total ? nullptr : jj < nright ? A + (long)(k0 + w + jj) * lda + k0 : B + (long)(jj - nright) * ldb + k0;
            }
            double v[JB][MQ];
#pragma unroll
            for (int q = 0; q < MQ; ++q) {
                const int row = q * 32 + lane;
                const int sr = row < m ? src[row] : 0;
#pragma unroll
                for (int jb = 0; jb < JB; ++jb) v[jb][q] = (row < m && cp[jb]) ? cp[jb][sr] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < NBL; ++c) {
                if (c < w) {
                    double u[JB];
#pragma unroll
                    for (int jb = 0; jb < JB; ++jb) u[jb] = __shfl_sync(0xffffffffu, v[jb][0], c);
#pragma unroll
                    for (int q = 0; q < MQ; ++q) {
                        const int row = q * 32 + lane;
                        if (row > c && row < m) {
                            const double l = P[c * LDP + row];
#pragma unroll
                            for (int jb = 0; jb < JB; ++jb) v[jb][q] = fma(-l, u[jb], v[jb][q]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < MQ; ++q) {
                const int row = q * 32 + lane;
#pragma unroll
                for (int jb = 0; jb < JB; ++jb)
                    if (row < m && cp[jb]) cp[jb][row] = v[jb][q];
            }
        }
        __syncthreads();
    }
    // ---- back substitution with U: one warp per JB right-hand sides, the column in registers; the NBL columns of U a
    // stage needs are copied into shared memory by the whole CTA first (the chain of n divisions is what the caller waits
    // for with one or two right-hand sides: nothing on it may wait for L2) ----
    const int ngroups = (nrhs + JB - 1) / JB;
    for (int g0 = 0; g0 < ngroups; g0 += NW) {
        const int g = g0 + warp;
        double* cp[JB];
#pragma unroll
        for (int jb = 0; jb < JB; ++jb) cp[jb] = g * JB + jb < nrhs ? B + (long)(g * JB + jb) * ldb : nullptr;
        double v[JB][MQ];
#pragma unroll
        for (int q = 0; q < MQ; ++q)
#pragma unroll
            for (int jb = 0; jb < JB; ++jb) v[jb][q] = (q * 32 + lane < n && cp[jb]) ? cp[jb][q * 32 + lane] : 0.0;
#pragma unroll
        for (int kb = MQ - 1; kb >= 0; --kb) {
#pragma unroll
            for (int h = 32 / NBL - 1; h >= 0; --h) {
                const int k1 = kb * 32 + h * NBL;
                if (k1 < n) {
                    const int wc = min(NBL, n - k1), nr = min(n, k1 + NBL);
                    __syncthreads();
                    for (int idx = tid; idx < wc * nr; idx += NT) {
                        const int cc = idx / nr, i = idx - cc * nr;
                        P[cc * LDP + i] = A[(long)(k1 + cc) * lda + i];
                    }
                    __syncthreads();
                    if (g < ngroups) {
#pragma unroll 4
                        for (int kk = NBL - 1; kk >= 0; --kk) {
                            const int k = k1 + kk, lk = h * NBL + kk;   // lk: lane that owns row k
                            if (k < n) {
                                const double* Uk = P + kk * LDP;
                                const double d = Uk[k];
                                double x[JB];
#pragma unroll
                                for (int jb = 0; jb < JB; ++jb) x[jb] = __shfl_sync(0xffffffffu, v[jb][kb], lk) / d;
                                const double ud = lane < lk ? Uk[kb * 32 + lane] : 0.0;
#pragma unroll
                                for (int jb = 0; jb < JB; ++jb)
                                    v[jb][kb] = lane == lk ? x[jb] : lane < lk ? fma(-ud, x[jb], v[jb][kb]) : v[jb][kb];
#pragma unroll
                                for (int q = 0; q < MQ; ++q) {
                                    if (q < kb) {
                                        const double uu = Uk[q * 32 + lane];
#pragma unroll
                                        for (int jb = 0; jb < JB; ++jb) v[jb][q] = fma(-uu, x[jb], v[jb][q]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MQ; ++q)
#pragma unroll
            for (int jb = 0; jb < JB; ++jb)
                if (q * 32 + lane < n && cp[jb]) cp[jb][q * 32 + lane] = v[jb][q];
    }
}

static bool g_lu_blocked = true;
void lu_set_blocked(bool on) { g_lu_blocked = on; }

template <int MQ, int NBL, int JB>
static void launch_lu_blocked(cudaStream_t st, int ctas, double* A, long lda, int n, double* B, long ldb, int nrhs, int* info,
                              const double* Asrc, long ldsrc, int cols_per_cta) {
    const size_t smem = (size_t)NBL * (32 * MQ + 1) * sizeof(double) + (size_t)32 * MQ * sizeof(int);
    cudaFuncSetAttribute(lu_blocked_kernel<MQ, NBL, JB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lu_blocked_kernel<MQ, NBL, JB><<<ctas, 512, smem, st>>>(A, lda, n, B, ldb, nrhs, info, Asrc, ldsrc, cols_per_cta); ++g_kernel_launches;
}

static void lu_dispatch(cudaStream_t st, int ctas, double* A, long lda, int n, double* B, long ldb, int nrhs, int* info,
                        const double* Asrc, long ldsrc, int cols_per_cta) {
    if (g_lu_blocked && n <= 256) launch_lu_blocked<8, 32, 4>(st, ctas, A, lda, n, B, ldb, nrhs, info, Asrc, ldsrc, cols_per_cta);
    else if (g_lu_blocked && n <= 640) launch_lu_blocked<20, 16, 2>(st, ctas, A, lda, n, B, ldb, nrhs, info, Asrc, ldsrc, cols_per_cta);
    else { lu_solve_kernel<<<ctas, 1024, 0, st>>>(A, lda, n, B, ldb, nrhs, info, Asrc, ldsrc, cols_per_cta); ++g_kernel_launches; }
}

void launch_lu_solve(cudaStream_t st, double* A, long lda, int n, double* B, long ldb, int nrhs, int* info) {
    cudaMemsetAsync(info, 0, sizeof(int), st);
    lu_dispatch(st, 1, A, lda, n, B, ldb, nrhs, info, nullptr, 0, 0);
}
void launch_lu_solve_multi(cudaStream_t st, const double* A, long lda, int n, double* scratch, double* B, long ldb, long nrhs,
                           int cols_per_cta, int* info) {
    cudaMemsetAsync(info, 0, sizeof(int), st);
    const int ctas = (int)((nrhs + cols_per_cta - 1) / cols_per_cta);
    lu_dispatch(st, ctas, scratch, n, n, B, ldb, (int)nrhs, info, A, lda, cols_per_cta);
}

// out[c] = sum_{r < rows} M(r, c), rows summed in order (one thread per column; rows <= a few hundred)
__global__ void colsum_kernel(const double* __restrict__ M, long ld, int rows, long cols, double* __restrict__ out) {
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += M[c * ld + r];
    out[c] = s;
}
void launch_colsum(cudaStream_t st, const double* M, long ld, int rows, long cols, double* out) {
    if (cols > 0) colsum_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(M, ld, rows, cols, out); ++g_kernel_launches;
}

// ---------------------------------------------------------------------------------------------------
// Utilities
// ---------------------------------------------------------------------------------------------------
__global__ void fill_kernel(double* p, long n, double v) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill(cudaStream_t st, double* p, long n, double v) {
    if (n > 0) fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v); ++g_kernel_launches;
}

// dst(r, c) = src(r, c) for r < rows, c < cols; optional transpose of the source.
__global__ void copy2d_kernel(double* dst, long ldd, const double* src, long lds, int rows, int cols, int trans) {
    __shared__ double t[32][33];
    int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if (!trans) {
        for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
            int r = r0 + threadIdx.x, c = c0 + cc;
            if (r < rows && c < cols) dst[(long)c * ldd + r] = src[(long)c * lds + r];
        }
    } else {
        // dst(r, c) = src(c, r)
        for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
            int sr = c0 + threadIdx.x, sc = r0 + cc;   // src row (contiguous) = dst col
            if (sr < cols && sc < rows) t[cc][threadIdx.x] = src[(long)sc * lds + sr];
        }
        __syncthreads();
        for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
            int r = r0 + threadIdx.x, c = c0 + cc;
            if (r < rows && c < cols) dst[(long)c * ldd + r] = t[threadIdx.x][cc];
        }
    }
}
void launch_copy2d(cudaStream_t st, double* dst, long ldd, const double* src, long lds, int rows, int cols,
                   bool trans) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((rows + 31) / 32, (cols + 31) / 32), block(32, 8);
    copy2d_kernel<<<grid, block, 0, st>>>(dst, ldd, src, lds, rows, cols, trans ? 1 : 0); ++g_kernel_launches;
}

// A(i,i) += v for i < n
__global__ void add_diag_kernel(double* A, long ld, int n, double v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(long)i * ld + i] += v;
}
void launch_add_diag(cudaStream_t st, double* A, long ld, int n, double v) {
    if (n > 0) add_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(A, ld, n, v); ++g_kernel_launches;
}

// Single-block reductions: mode 0 trace(A), 1 sum(x), 2 dot(x,y), 3 sum of all entries of A (n x n2)
__global__ void reduce_kernel(int mode, const double* x, const double* y, long ld, long n, long n2, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    if (mode == 0) for (long i = threadIdx.x; i < n; i += blockDim.x) s += x[i * ld + i];
    else if (mode == 1) for (long i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    else if (mode == 2) for (long i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * y[i];
    else for (long idx = threadIdx.x; idx < n * n2; idx += blockDim.x) s += x[(idx / n) * ld + idx % n];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = s;
    }
}
void launch_reduce(cudaStream_t st, int mode, const double* x, const double* y, long ld, long n, long n2, double* out) {
    reduce_kernel<<<1, 1024, 0, st>>>(mode, x, y, ld, n, n2, out); ++g_kernel_launches;
}

// M(r, c) += v for r < rows, c < cols
__global__ void add_const_kernel(double* M, long ld, int rows, long cols, double v) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)rows * cols) return;
    M[(idx / rows) * ld + idx % rows] += v;
}
void launch_add_const(cudaStream_t st, double* M, long ld, int rows, long cols, double v) {
    long tot = (long)rows * cols;
    if (tot > 0 && v != 0.0) add_const_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(M, ld, rows, cols, v); ++g_kernel_launches;
}

// y = a*x + b*y elementwise over n
__global__ void axpby_kernel(long n, double a, const double* x, double b, double* y) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a * x[i] + (b != 0.0 ? b * y[i] : 0.0);
}
void launch_axpby(cudaStream_t st, long n, double a, const double* x, double b, double* y) {
    if (n > 0) axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, x, b, y); ++g_kernel_launches;
}

// Mirror the lower triangle of an n x n matrix into the upper (copytri!)
__global__ void mirror_lower_kernel(double* A, long ld, int n) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    int r = (int)(idx % n), c = (int)(idx / n);
    if (r < c) A[(long)c * ld + r] = A[(long)r * ld + c];
}
void launch_mirror_lower(cudaStream_t st, double* A, long ld, int n) {
    long tot = (long)n * n;
    if (tot > 0) mirror_lower_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(A, ld, n); ++g_kernel_launches;
}

}  // namespace geordd
