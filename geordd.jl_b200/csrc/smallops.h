// Launchers of the small kernels in stats.cu (device pointers throughout).
#pragma once
#include <cuda_runtime.h>

namespace geordd {

struct FinishMll {
    const double *partN, *partT, *partC;   // [nb? x ncols] dot partials from the backward solves
    int nbN, nbT, nbC;
    long ncols;                             // padded draw count of this batch (row stride of partials)
    long ndraws;                            // live draws in this batch
    double halflogdetN, halflogdetT, halflogdetC;   // logdet/2
    double constN, constT, constC;                  // nobs*log(2pi)/2
    double dmll_obs;
    double *out_null, *out_alt;             // nullable, indexed out_off + c
    long out_off;
    unsigned long long* tally;
    unsigned long long* margin;             // atomicMin of the bits of |statistic - threshold| (certificate for the exact count)
};
void launch_finish_mll(cudaStream_t st, const FinishMll& f);

struct FinishCliff {
    const double* part_chi;   // nullable
    const double* part_num;   // nullable
    int nb;                   // row blocks of the sentinel system
    long ncols, ndraws;
    double chi_obs, denom, pval_obs;
    double *out_chi, *out_pval, *out_num;   // nullable
    long out_off;
    unsigned long long *tally_chi, *tally_inv;
    unsigned long long *margin_chi, *margin_inv;   // as FinishMll::margin
};
void launch_finish_cliff(cudaStream_t st, const FinishCliff& f);

void launch_frac_compare(cudaStream_t st, const double* ref, long nref, const double* x, long n, bool greater,
                         double* out);
void launch_lu_solve(cudaStream_t st, double* A, long lda, int n, double* B, long ldb, int nrhs, int* info);
void lu_set_blocked(bool on);   // test hook: false = always the unblocked kernel
// The same n x n system against nrhs right-hand sides split over CTAs; scratch: ceil(nrhs / cols_per_cta) * n * n doubles.
void launch_lu_solve_multi(cudaStream_t st, const double* A, long lda, int n, double* scratch, double* B, long ldb, long nrhs,
                           int cols_per_cta, int* info);
void launch_colsum(cudaStream_t st, const double* M, long ld, int rows, long cols, double* out);
void launch_fill(cudaStream_t st, double* p, long n, double v);
void launch_copy2d(cudaStream_t st, double* dst, long ldd, const double* src, long lds, int rows, int cols, bool trans);
void launch_add_diag(cudaStream_t st, double* A, long ld, int n, double v);
// mode 0 trace(A[n x n], ld), 1 sum(x[n]), 2 dot(x,y)[n], 3 sum of entries of x[n x n2] (ld)
void launch_reduce(cudaStream_t st, int mode, const double* x, const double* y, long ld, long n, long n2, double* out);
void launch_add_const(cudaStream_t st, double* M, long ld, int rows, long cols, double v);
void launch_axpby(cudaStream_t st, long n, double a, const double* x, double b, double* y);
void launch_mirror_lower(cudaStream_t st, double* A, long ld, int n);

}  // namespace geordd
