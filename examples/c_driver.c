/* Minimal C consumer of the geordd_b200 C ABI: fits two regions, computes the cliff face at three sentinels and runs
 * the chi-square sharp-null test.  Shows that the boundary is plain C (no C++ / torch / Python types):
 *     gcc -std=c99 -Iinclude examples/c_driver.c -Lgeordd.jl_b200/lib -lgeordd_b200 -Wl,-rpath,$PWD/geordd.jl_b200/lib -lm -o c_driver
 * Prints: mu[0..2], Sigma[0][0], chi2 statistic, p-value from 2000 null draws (seed 1). */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "geordd_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc__ = (call);                                                                   \
        if (rc__ != 0) {                                                                     \
            fprintf(stderr, "%s failed: rc=%d (%s)\n", #call, rc__, geordd_last_error(ctx)); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(void) {
    enum { M = 60, NB = 3 };
    double XT[2 * M], XC[2 * M], YT[M], YC[M], Xb[2 * NB], mu[NB], Sigma[NB * NB];
    unsigned s = 12345u;
    int i;
    for (i = 0; i < M; ++i) {   /* treated: x in [0, 0.5), control: x in [0.5, 1); small LCG, fixed data */
        s = s * 1664525u + 1013904223u; XT[2 * i] = 0.5 * (s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; XT[2 * i + 1] = (s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; XC[2 * i] = 0.5 + 0.5 * (s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; XC[2 * i + 1] = (s >> 8) / 16777216.0;
        YT[i] = sin(3.0 * XT[2 * i]) + XT[2 * i + 1] + 0.4;
        YC[i] = sin(3.0 * XC[2 * i]) + XC[2 * i + 1];
    }
    for (i = 0; i < NB; ++i) { Xb[2 * i] = 0.5; Xb[2 * i + 1] = 0.25 * (i + 1); }

    geordd_ctx* ctx = NULL;
    if (geordd_ctx_create(&ctx, 0) != 0 || !ctx) { fprintf(stderr, "no usable GPU / context\n"); return 2; }
    geordd_kernel k;
    k.kind = 0; k.ell = 0.3; k.sigma2 = 1.0; k.const_sigma2 = 0.0; k.mean_const = 0.0;   /* SEIso(log 0.3, 0), MeanZero */
    geordd_gp *gT = NULL, *gC = NULL;
    double mllT = 0.0, mllC = 0.0, chi_obs = 0.0;
    CHECK(geordd_gp_fit(ctx, &k, XT, 2, M, YT, log(0.1), &gT, NULL, &mllT, NULL, NULL));
    CHECK(geordd_gp_fit(ctx, &k, XC, 2, M, YC, log(0.1), &gC, NULL, &mllC, NULL, NULL));
    CHECK(geordd_cliff_face(gT, gC, Xb, NB, mu, Sigma));
    CHECK(geordd_chistat_obs(gT, gC, Xb, NB, &chi_obs));
    geordd_null* hn = NULL;
    CHECK(geordd_null_prepare(gT, gC, Xb, NB, 0.0, &hn, NULL, NULL));
    long long count = 0;
    CHECK(geordd_null_chi2(hn, 2000, 1, 0, NULL, chi_obs, NULL, &count));
    printf("%.15g %.15g %.15g %.15g %.15g %.15g\n", mu[0], mu[1], mu[2], Sigma[0], chi_obs, (double)count / 2000.0);
    geordd_null_destroy(hn);
    geordd_gp_destroy(gT);
    geordd_gp_destroy(gC);
    geordd_ctx_destroy(ctx);
    return 0;
}
