/* A plain-C99 consumer of include/gslnls_b200.h: what the reference's .Call shim does
 * minus the SEXP conversions.  Reads "n" then n pairs "x y" from the file named on the command line, fits
 * y ~ A * exp(-lam * x) + b from (A, lam, b) = (0, 0, 0) with gsl_nls_control() defaults and prints the
 * result list.  Used by tests/test_c_consumer.py (compiled with gcc -std=c99 -pedantic). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "gslnls_b200.h"

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s data.txt [ngpu]\n", argv[0]);
        return 2;
    }
    FILE *fh = fopen(argv[1], "r");
    long n = 0;
    if (!fh || fscanf(fh, "%ld", &n) != 1 || n < 1)
        return 2;
    double *x = malloc(sizeof(double) * (size_t)n), *y = malloc(sizeof(double) * (size_t)n);
    for (long i = 0; i < n; ++i)
        if (fscanf(fh, "%lf %lf", &x[i], &y[i]) != 2)
            return 2;
    fclose(fh);
    const int ngpu = argc > 2 ? atoi(argv[2]) : 1;

    const char *pn[] = {"A", "lam", "b"}, *vn[] = {"x"};
    char err[1024] = "";
    gslnls_model *m = NULL;
    int rc = gslnls_model_compile("A * exp(-lam * x) + b", pn, 3, vn, 1, GSLNLS_JAC_SYMBOLIC, GSLNLS_FVV_NONE, &m,
                                  err, sizeof err);
    if (rc) {
        printf("compile_rc %d %s\n", rc, err);
        return 1;
    }
    /* .ctrl_int / .ctrl_dbl as R/nls_large.R:383-407 packs them: maxiter, trace, algorithm (0 = lm),
       scale (0 = more), fdtype, jacclass (-2 = dense), jacnz; factor_up, factor_down, avmax, h_df, h_fvv,
       xtol, ftol, gtol */
    const double eps = sqrt(2.220446049250313e-16);
    const int ci[7] = {100, 0, 0, 0, 0, -2, 0};
    const double cd[8] = {2.0, 3.0, 0.75, eps, 0.02, eps, eps, eps};
    const double start[3] = {0.0, 0.0, 0.0};
    const double *cols[1];
    cols[0] = x;
    gslnls_result res;
    rc = gslnls_fit_large_multi(m, cols, y, NULL, (int64_t)n, start, ci, cd, ngpu, NULL, 1, &res);
    printf("rc %d\n", rc);
    if (rc >= 1000) {
        printf("error %s: %s\n", gslnls_strerror(rc), gslnls_last_error());
    } else {
        printf("status %s\nalgorithm %s\nniter %d\nconv %d\nssr %.17g\n", res.status, res.algorithm, res.niter,
               res.conv, res.ssr);
        printf("par %.17g %.17g %.17g\n", res.par[0], res.par[1], res.par[2]);
        printf("covar00 %.17g\n", res.covar[0]);
        double s = 0.0;
        for (long i = 0; i < n; ++i)
            s += res.resid[i] * res.resid[i];
        printf("resid_ss %.17g\n", s);
        gslnls_result_free(&res);
    }
    gslnls_model_free(m);
    gslnls_cache_clear();
    free(x);
    free(y);
    return 0;
}
