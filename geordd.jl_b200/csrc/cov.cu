// K1: covariance-matrix assembly (SE / Matern-1/2,3/2,5/2 isotropic kernels + Const + noise on the
// diagonal).  Replaces GaussianProcesses.cov / cov! / cov_ij over KernelData distance matrices
// (call sites: GPE ctor at src/region_gp_fit.jl:12,28 and src/hypotest.jl:4; addcov!/add_diag!
// src/gp_utils.jl:24-44; cov(kernel, Xb, gp.x) src/hypotest_chi2.jl:38-39, src/power.jl:60-62).
//
// One 64x64 output tile per CTA, the 2x64 coordinate columns of the tile are staged in shared memory once, distances
// are direct differences (no Gram trick), every store is a coalesced 512-byte row segment.  Symmetric
// assembly computes only the tiles on or below the diagonal and mirrors them through a padded smem transpose, halving
// the exp() work.
//
// What bounds it (round 2, measured): NOT the HBM write.  An entry costs ~50 FP64-pipe instructions with the CUDA math
// library (sqrt ~10, the division by ell ~12, exp ~22, the rest ~6) against 8 bytes stored; B200 issues 64 FP64 FMA per
// clock per SM, so 32.8 M lower-triangle entries of the n = 8064 pooled matrix need >= 88 us of FP64 issue, while the
// 262 MB they occupy need 40 us of HBM time.  The fast path below (cov2_kernel: 2-d coordinates, the case of every fit)
// therefore cuts FP64 instructions first -- reciprocals of ell hoisted to the host, r = r2 * rsqrt(r2) from the MUFU seed --
// its own branch-free exp for non-positive arguments, the kernel family a template parameter, interior tiles without bounds
// tests -- and only then widens the stores to 16 bytes (two rows per thread).
#include <cstdint>
#include "kernels.h"

namespace geordd {

constexpr int CT = 64;
constexpr int MAXDIM = 8;

__device__ __forceinline__ double kern_eval(int kind, double r2, double ell, double sigma2, double csig2) {
    double k;
    if (kind == 0) {
        k = sigma2 * exp(-0.5 * r2 / (ell * ell));
    } else {
        double r = sqrt(r2);
        if (kind == 1) {
            k = sigma2 * exp(-r / ell);
        } else if (kind == 2) {
            double s = 1.7320508075688772 * r / ell;
            k = sigma2 * (1.0 + s) * exp(-s);
        } else {
            double s = 2.23606797749979 * r / ell;
            k = sigma2 * (1.0 + s + s * s / 3.0) * exp(-s);
        }
    }
    return k + csig2;
}

struct CovParams {
    KernelSpecDev spec;
    const double* X1; int n1;
    const double* X2; int n2;
    int dim;
    int sym;          // 1: X1 == X2, compute tiles bi >= bj only
    int mirror;       // sym only: also write the transposed tile
    double diag_add;
    double* K; long ldk;
    int n1p, n2p;     // padded extents to fill (>= n1, n2)
    int accumulate;   // K += instead of K =
};

__global__ void __launch_bounds__(256) cov_kernel(CovParams p) {
    __shared__ double xs1[MAXDIM][CT], xs2[MAXDIM][CT];
    __shared__ double tile[CT][CT + 1];
    int bi, bj;
    if (p.sym) {
        // linear index over the lower triangle of tile pairs
        int t = blockIdx.x;
        int bjj = (int)floor((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((long)(bjj + 1) * (bjj + 2) / 2 <= t) ++bjj;
        while ((long)bjj * (bjj + 1) / 2 > t) --bjj;
        bi = bjj;                       // row tile
        bj = t - bjj * (bjj + 1) / 2;   // col tile <= row tile
    } else {
        bi = blockIdx.x;
        bj = blockIdx.y;
    }
    const int tid = threadIdx.x;
    const int i0 = bi * CT, j0 = bj * CT;
    for (int idx = tid; idx < p.dim * CT; idx += 256) {
        int d = idx % p.dim, c = idx / p.dim;
        xs1[d][c] = (i0 + c < p.n1) ? p.X1[(long)(i0 + c) * p.dim + d] : 0.0;
        xs2[d][c] = (j0 + c < p.n2) ? p.X2[(long)(j0 + c) * p.dim + d] : 0.0;
    }
    __syncthreads();
    const int li = tid % CT, lj0 = tid / CT;
    const int i = i0 + li;
    double xi[MAXDIM];
    for (int d = 0; d < p.dim; ++d) xi[d] = xs1[d][li];
#pragma unroll 4
    for (int jj = 0; jj < CT / 4; ++jj) {
        const int lj = lj0 + 4 * jj;
        const int j = j0 + lj;
        double v;
        if (i < p.n1 && j < p.n2) {
            double r2 = 0.0;
            for (int d = 0; d < p.dim; ++d) {
                double df = xi[d] - xs2[d][lj];
                r2 += df * df;
            }
            v = kern_eval(p.spec.kind, r2, p.spec.ell, p.spec.sigma2, p.spec.const_sigma2);
            if (p.sym && i == j) v += p.diag_add;
        } else {
            v = (p.sym && i == j) ? 1.0 : 0.0;
        }
        if (i < p.n1p && j < p.n2p) {
            double* dst = p.K + (long)j * p.ldk + i;
            if (p.accumulate) *dst += v; else *dst = v;
        }
        if (p.mirror) tile[li][lj] = v;
    }
    if (p.mirror && bi != bj) {
        __syncthreads();
        // transposed tile: element (row = j0 + a, col = i0 + b) = tile[b][a]; a contiguous
        const int a = tid % CT, b0 = tid / CT;
#pragma unroll 4
        for (int bb = 0; bb < CT / 4; ++bb) {
            const int b = b0 + 4 * bb;
            const int row = j0 + a, col = i0 + b;
            if (row < p.n2p && col < p.n1p) {
                double* dst = p.K + (long)col * p.ldk + row;
                if (p.accumulate) *dst += tile[b][a]; else *dst = tile[b][a];
            }
        }
    }
}

// ---- fast path: dim == 2, plain assignment, even leading dimension and 16-byte aligned K (all device-resident fits) ------
struct Cov2Consts {
    int kind;
    double a;        // SE: -0.5 / ell^2 ; Matern: sqrt(1|3|5) / ell
    double sigma2, csig2;
};

__device__ __forceinline__ double rsqrt_seeded(double x) {   // x > 0 normal; MUFU.RSQ64H + one third-order step
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// exp(x) for x <= 0, branch-free, 17 FP64 instructions: x = k ln2 + r with k = rint(x log2 e) taken from the low mantissa
// bits of (x log2 e + 1.5 * 2^52), |r| <= ln2 / 2, exp(r) by a degree-13 Taylor polynomial (remainder 4e-18), 2^k added
// into the exponent field.  Relative error <= 2.3e-16 over [-700, 0] (checked against libm); below -700 the result is 0
// (a covariance term of 1e-304 sigma^2).  The CUDA library exp() costs ~22 FP64 instructions plus a range branch.
// (coefficients live in constant memory: a DFMA takes them as c[bank][offset] operands, whereas 64-bit immediates are
//  re-materialised with a UMOV + IMAD.MOV pair per use -- 22 extra issue slots per entry in the first version)
__constant__ double kExpC[16] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
                                 2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
                                 8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
                                 1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10};
__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, kExpC[12], kExpC[13]);
    const int k = __double2loint(t);
    const double kf = t - kExpC[13];
    double r = fma(kf, kExpC[14], x);
    r = fma(kf, kExpC[15], r);
    double p = kExpC[0];                      // 1/13!
#pragma unroll
    for (int n = 1; n < 12; ++n) p = fma(p, r, kExpC[n]);   // 1/12! ... 1/2!
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return x < -700.0 ? 0.0 : p;
}

template <int KIND>
__device__ __forceinline__ double kern_eval2(const Cov2Consts& k, double r2) {
    double v;
    if (KIND == 0) {
        v = k.sigma2 * exp_nonpos(r2 * k.a);
    } else {
        const double r = r2 * rsqrt_seeded(fmax(r2, 1e-280));   // branch-free: r2 == 0 gives 0 * 1e140 = 0
        const double s = k.a * r;
        const double e = exp_nonpos(-s);
        if (KIND == 1) v = k.sigma2 * e;
        else if (KIND == 2) v = k.sigma2 * (1.0 + s) * e;
        else v = k.sigma2 * (1.0 + s + s * s * (1.0 / 3.0)) * e;
    }
    return v + k.csig2;
}

// KIND: kernel family (compile time); FULL: the 64 x 64 tile lies entirely inside the logical matrix and off the diagonal
// of a symmetric assembly -- no per-entry bounds tests, no diagonal term.
template <int KIND, bool FULL>
__device__ __forceinline__ void cov2_tile(const CovParams& p, const Cov2Consts& kc, const double2* xs1, const double2* xs2,
                                          double (*tile)[CT + 2], int i0, int j0) {
    const int tid = threadIdx.x;
    // thread = one PAIR of rows (2 li, 2 li + 1) x 8 columns; a warp stores 32 x 16 B = one 512-byte segment per column
    const int li = tid % 32, lj0 = tid / 32;
    const int r0 = i0 + 2 * li;
    const double2 xa = xs1[2 * li], xb = xs1[2 * li + 1];
#pragma unroll 4
    for (int jj = 0; jj < CT / 8; ++jj) {
        const int lj = lj0 + 8 * jj;
        const int j = j0 + lj;
        const double2 xc = xs2[lj];
        double v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = r0 + h;
            const double2 xr = h ? xb : xa;
            const double dx = xr.x - xc.x, dy = xr.y - xc.y;
            if (FULL) {
                v[h] = kern_eval2<KIND>(kc, fma(dy, dy, dx * dx));
            } else if (i < p.n1 && j < p.n2) {
                v[h] = kern_eval2<KIND>(kc, fma(dy, dy, dx * dx));
                if (p.sym && i == j) v[h] += p.diag_add;
            } else {
                v[h] = (p.sym && i == j) ? 1.0 : 0.0;
            }
        }
        if (FULL || (j < p.n2p && r0 < p.n1p))   // n1p is even on this path
            *reinterpret_cast<double2*>(p.K + (long)j * p.ldk + r0) = make_double2(v[0], v[1]);
        if (p.mirror) { tile[2 * li][lj] = v[0]; tile[2 * li + 1][lj] = v[1]; }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) cov2_kernel(CovParams p, Cov2Consts kc, int ntiles, int tb1) {
    __shared__ double2 xs1[CT], xs2[CT];
    extern __shared__ __align__(16) double tile_raw[];   // [CT][CT + 2], only allocated for mirrored assemblies
    double (*tile)[CT + 2] = reinterpret_cast<double (*)[CT + 2]>(tile_raw);
    const int tid = threadIdx.x;
    // (written as a loop over tiles; launched with one tile per CTA, see launch_cov2)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int bi, bj;
        if (p.sym) {   // linear index over the lower triangle of tile pairs
            int bjj = (int)floor((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while ((long)(bjj + 1) * (bjj + 2) / 2 <= t) ++bjj;
            while ((long)bjj * (bjj + 1) / 2 > t) --bjj;
            bi = bjj;
            bj = t - bjj * (bjj + 1) / 2;
        } else {
            bi = t % tb1;
            bj = t / tb1;
        }
        const int i0 = bi * CT, j0 = bj * CT;
        if (tid < CT) {
            const int i = i0 + tid;
            xs1[tid] = i < p.n1 ? *reinterpret_cast<const double2*>(p.X1 + 2L * i) : make_double2(0.0, 0.0);
        } else if (tid < 2 * CT) {
            const int j = j0 + tid - CT;
            xs2[tid - CT] = j < p.n2 ? *reinterpret_cast<const double2*>(p.X2 + 2L * j) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const bool full = i0 + CT <= p.n1 && j0 + CT <= p.n2 && !(p.sym && bi == bj);   // uniform
        if (full) cov2_tile<KIND, true>(p, kc, xs1, xs2, tile, i0, j0);
        else cov2_tile<KIND, false>(p, kc, xs1, xs2, tile, i0, j0);
        if (p.mirror && bi != bj) {
            __syncthreads();
            // transposed tile: element (row = j0 + a, col = i0 + b) = tile[b][a]; a contiguous, two rows per thread
            const int a2 = 2 * (tid % 32), b0 = tid / 32;
#pragma unroll 4
            for (int bb = 0; bb < CT / 8; ++bb) {
                const int b = b0 + 8 * bb;
                const int row = j0 + a2, col = i0 + b;
                if (row < p.n2p && col < p.n1p)
                    *reinterpret_cast<double2*>(p.K + (long)col * p.ldk + row) = make_double2(tile[b][a2], tile[b][a2 + 1]);
            }
        }
        __syncthreads();   // xs1 / xs2 / tile are rewritten by the next tile
    }
}

template <int KIND>
static void launch_cov2(cudaStream_t st, const CovParams& p, const Cov2Consts& kc, int tb1, int tb2) {
    const int ntiles = p.sym ? tb1 * (tb1 + 1) / 2 : tb1 * tb2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = p.mirror ? sizeof(double) * CT * (CT + 2) : 0;
    (void)sms;
    const int grid = ntiles;   // one tile per CTA: measured faster than a persistent loop over tiles (97 vs 104 us at n = 8064)
    cov2_kernel<KIND><<<grid, 256, smem, st>>>(p, kc, ntiles, tb1); ++g_kernel_launches;
}

void launch_cov(cudaStream_t st, const KernelSpecDev& spec, const double* X1, int n1, const double* X2, int n2,
                int dim, bool sym, bool mirror, double diag_add, double* K, long ldk, int n1p, int n2p,
                bool accumulate) {
    CovParams p;
    p.spec = spec; p.X1 = X1; p.n1 = n1; p.X2 = X2; p.n2 = n2; p.dim = dim; p.sym = sym ? 1 : 0;
    p.mirror = (sym && mirror) ? 1 : 0; p.diag_add = diag_add; p.K = K; p.ldk = ldk; p.n1p = n1p; p.n2p = n2p;
    p.accumulate = accumulate ? 1 : 0;
    int tb1 = (n1p + CT - 1) / CT, tb2 = (n2p + CT - 1) / CT;
    const bool fast = dim == 2 && !accumulate && ldk % 2 == 0 && n1p % 2 == 0 && (!p.mirror || n2p % 2 == 0) &&
                      reinterpret_cast<uintptr_t>(K) % 16 == 0 && reinterpret_cast<uintptr_t>(X1) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(X2) % 16 == 0;
    if (fast) {
        Cov2Consts kc;
        kc.kind = spec.kind; kc.sigma2 = spec.sigma2; kc.csig2 = spec.const_sigma2;
        kc.a = spec.kind == 0 ? -0.5 / (spec.ell * spec.ell)
             : (spec.kind == 1 ? 1.0 : spec.kind == 2 ? 1.7320508075688772 : 2.23606797749979) / spec.ell;
        switch (spec.kind) {
            case 0: launch_cov2<0>(st, p, kc, tb1, tb2); break;
            case 1: launch_cov2<1>(st, p, kc, tb1, tb2); break;
            case 2: launch_cov2<2>(st, p, kc, tb1, tb2); break;
            default: launch_cov2<3>(st, p, kc, tb1, tb2); break;
        }
        return;
    }
    if (sym) {
        int nt = tb1 * (tb1 + 1) / 2;
        cov_kernel<<<nt, 256, 0, st>>>(p); ++g_kernel_launches;
    } else {
        dim3 grid(tb1, tb2);
        cov_kernel<<<grid, 256, 0, st>>>(p); ++g_kernel_launches;
    }
}

}  // namespace geordd
