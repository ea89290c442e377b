/* boot_mlltest and boot_chi2test on SEVERAL GPUs from plain C -- no Python, no MPI: the multi-device group of the C ABI
 * (geordd_group_*) replicates the fits, shards the null draws by index and sums the exceedance tallies with one NCCL
 * all-reduce inside the library.  The program runs the same tests on 1 GPU and on `ndev` GPUs and checks that p-values,
 * tallies and every per-draw statistic agree BIT FOR BIT (a draw depends only on (seed, draw index)).
 *     gcc -std=c99 -Iinclude examples/c_driver_multi.c -Lgeordd.jl_b200/lib -lgeordd_b200 -Wl,-rpath,$PWD/geordd.jl_b200/lib -lm -o c_driver_multi
 *     ./c_driver_multi [ndev=2] [m=600] [nsim=4000]
 * Exit code 0 = identical; 2 = no usable GPU; 3 = fewer than ndev GPUs / NCCL missing; 1 = mismatch or error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "geordd_b200.h"

static unsigned lcg_state = 2463534242u;
static double lcg(void) { lcg_state = lcg_state * 1664525u + 1013904223u; return (lcg_state >> 8) / 16777216.0; }

typedef struct { double p_mll, p_chi; long long n_mll, n_chi; double *mn, *ma, *chi; } result;

static int run(int ndev, int m, int nb, long nsim, const double* XT, const double* XC, const double* YT, const double* YC,
               const double* Xb, result* r) {
    geordd_group* g = NULL;
    int rc = geordd_group_create(&g, ndev, NULL);
    if (rc != 0) { fprintf(stderr, "geordd_group_create(%d) failed: rc=%d\n", ndev, rc); return rc == GEORDD_ERR_CUDA && ndev == 1 ? 2 : 3; }
    geordd_kernel k;
    k.kind = 2; k.ell = 0.3; k.sigma2 = 1.0; k.const_sigma2 = 25.0; k.mean_const = 0.0;   /* Mat32Iso + Const(5), MeanZero */
    geordd_ggp *gT = NULL, *gC = NULL;
    geordd_gnull* hn = NULL;
    double mllT = 0, mllC = 0, mllN = 0, chi_obs = 0;
#define GCHECK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "%s failed: rc=%d (%s)\n", #call, rc__, geordd_group_last_error(g)); return 1; } } while (0)
    GCHECK(geordd_group_gp_fit(g, &k, XT, 2, m, YT, log(0.3), &gT, NULL, &mllT, NULL));
    GCHECK(geordd_group_gp_fit(g, &k, XC, 2, m, YC, log(0.3), &gC, NULL, &mllC, NULL));
    GCHECK(geordd_group_null_prepare(gT, gC, Xb, nb, 0.0, &hn, NULL, &mllN));
    /* boot_mlltest (src/hypotest_mll.jl:32-45) */
    GCHECK(geordd_group_null_mll(hn, nsim, 7, 0, NULL, mllT + mllC - mllN, r->mn, r->ma, &r->n_mll));
    r->p_mll = (double)r->n_mll / (double)nsim;
    /* boot_chi2test (src/hypotest_chi2.jl:54-59): observed statistic on device 0's replicas */
    if (geordd_chistat_obs(geordd_group_gp_replica(gT, 0), geordd_group_gp_replica(gC, 0), Xb, nb, &chi_obs) != 0) return 1;
    GCHECK(geordd_group_null_chi2(hn, nsim, 7, 0, NULL, chi_obs, r->chi, &r->n_chi));
    r->p_chi = (double)r->n_chi / (double)nsim;
    geordd_group_null_destroy(hn);
    geordd_group_gp_destroy(gT);
    geordd_group_gp_destroy(gC);
    geordd_group_destroy(g);
    return 0;
}

int main(int argc, char** argv) {
    const int ndev = argc > 1 ? atoi(argv[1]) : 2, m = argc > 2 ? atoi(argv[2]) : 600, nb = 40;
    const long nsim = argc > 3 ? atol(argv[3]) : 4000;
    double *XT = malloc(sizeof(double) * 2 * m), *XC = malloc(sizeof(double) * 2 * m), *YT = malloc(sizeof(double) * m),
           *YC = malloc(sizeof(double) * m), *Xb = malloc(sizeof(double) * 2 * nb);
    int i, rc;
    for (i = 0; i < m; ++i) {   /* treated: x in [0, 0.5), control: x in [0.5, 1) */
        XT[2 * i] = 0.5 * lcg(); XT[2 * i + 1] = lcg();
        XC[2 * i] = 0.5 + 0.5 * lcg(); XC[2 * i + 1] = lcg();
        YT[i] = sin(3.0 * XT[2 * i]) + XT[2 * i + 1] + 0.3 * (lcg() - 0.5) + 0.1;
        YC[i] = sin(3.0 * XC[2 * i]) + XC[2 * i + 1] + 0.3 * (lcg() - 0.5);
    }
    for (i = 0; i < nb; ++i) { Xb[2 * i] = 0.5; Xb[2 * i + 1] = (i + 0.5) / nb; }
    result a, b;
    a.mn = malloc(sizeof(double) * nsim); a.ma = malloc(sizeof(double) * nsim); a.chi = malloc(sizeof(double) * nsim);
    b.mn = malloc(sizeof(double) * nsim); b.ma = malloc(sizeof(double) * nsim); b.chi = malloc(sizeof(double) * nsim);
    rc = run(1, m, nb, nsim, XT, XC, YT, YC, Xb, &a);
    if (rc != 0) return rc;
    rc = run(ndev, m, nb, nsim, XT, XC, YT, YC, Xb, &b);
    if (rc != 0) return rc;
    printf("1 GPU : boot_mlltest p = %.6f (%lld / %ld)   boot_chi2test p = %.6f (%lld / %ld)\n", a.p_mll, a.n_mll, nsim, a.p_chi, a.n_chi, nsim);
    printf("%d GPUs: boot_mlltest p = %.6f (%lld / %ld)   boot_chi2test p = %.6f (%lld / %ld)\n", ndev, b.p_mll, b.n_mll, nsim, b.p_chi, b.n_chi, nsim);
    if (a.n_mll != b.n_mll || a.n_chi != b.n_chi || memcmp(a.mn, b.mn, sizeof(double) * nsim) != 0 ||
        memcmp(a.ma, b.ma, sizeof(double) * nsim) != 0 || memcmp(a.chi, b.chi, sizeof(double) * nsim) != 0) {
        fprintf(stderr, "MISMATCH between 1 and %d GPUs\n", ndev);
        return 1;
    }
    printf("identical bit for bit: tallies, p-values and all %ld x 3 per-draw statistics\n", nsim);
    return 0;
}
