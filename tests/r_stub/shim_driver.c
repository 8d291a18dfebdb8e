/* tests/r_stub/shim_driver.c -- TEST INFRASTRUCTURE: plays the part of R/nls_large_cuda.R for the compiled
 * shim.  Reads a problem from a text file (n, then n lines "x y [w]"), looks the routines up in the
 * registration table of r-package/src/init.c as .Call would, fits "y ~ A * exp(-lam * x) + b", then asks
 * for the lazy residuals / gradient, and prints everything as JSON. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

void R_init_gslnlscuda(DllInfo *dll);

static DL_FUNC lookup(const DllInfo *dll, const char *name, int nargs)
{
    for (const R_CallMethodDef *d = dll->call; d && d->name; ++d)
        if (strcmp(d->name, name) == 0 && d->numArgs == nargs)
            return d->fun;
    fprintf(stderr, "routine %s/%d not registered\n", name, nargs);
    exit(4);
}

static SEXP strvec(int n, const char **s)
{
    SEXP v = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; ++i)
        SET_STRING_ELT(v, i, Rf_mkChar(s[i]));
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: shim_driver data.txt algorithm(0..5) weights_mode(0|1) [ngpu]\n");
        return 2;
    }
    FILE *fh = fopen(argv[1], "r");
    int n = 0, has_w = 0;
    if (!fh || fscanf(fh, "%d %d", &n, &has_w) != 2)
        return 2;
    SEXP x = Rf_allocVector(REALSXP, n), y = Rf_allocVector(REALSXP, n), w = has_w ? Rf_allocVector(REALSXP, n) : R_NilValue;
    for (int i = 0; i < n; ++i) {
        if (fscanf(fh, "%lf %lf", &REAL(x)[i], &REAL(y)[i]) != 2)
            return 2;
        if (has_w && fscanf(fh, "%lf", &REAL(w)[i]) != 1)
            return 2;
    }
    fclose(fh);
    const int alg = atoi(argv[2]), wmode = atoi(argv[3]), ngpu = argc > 4 ? atoi(argv[4]) : 1;
    DllInfo dll = {NULL};
    R_init_gslnlscuda(&dll);
    typedef SEXP (*fit_fn)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    typedef SEXP (*eval_fn)(SEXP, SEXP, SEXP);
    typedef SEXP (*free_fn)(SEXP);
    fit_fn fit = (fit_fn)lookup(&dll, "C_nls_large_cuda", 11);
    eval_fn ev = (eval_fn)lookup(&dll, "C_nls_large_cuda_eval", 3);
    free_fn fr = (free_fn)lookup(&dll, "C_nls_large_cuda_free", 1);

    const char *pn[] = {"A", "lam", "b"}, *vn[] = {"x"};
    SEXP cols = Rf_allocVector(VECSXP, 1);
    SET_VECTOR_ELT(cols, 0, x);
    SEXP start = Rf_allocVector(REALSXP, 3);
    REAL(start)[0] = 1.0; REAL(start)[1] = 1.0; REAL(start)[2] = 0.0;
    SEXP ci = Rf_allocVector(INTSXP, 7), cd = Rf_allocVector(REALSXP, 8);
    const int civ[7] = {100, 1, alg, 0, 0, -2, 0};                       /* R/nls_large.R:383-391 */
    const double cdv[8] = {2.0, 3.0, 0.75, 1.4901161193847656e-08, 0.02, /* R/nls.R:1187-1188 */
                           1.4901161193847656e-08, 1.4901161193847656e-08, 1.4901161193847656e-08};
    memcpy(INTEGER(ci), civ, sizeof civ);
    memcpy(REAL(cd), cdv, sizeof cdv);
    SEXP modes = Rf_allocVector(INTSXP, 3), devices = Rf_allocVector(INTSXP, ngpu);
    INTEGER(modes)[0] = 0; INTEGER(modes)[1] = alg == 1 ? 1 : 0; INTEGER(modes)[2] = wmode;
    for (int d = 0; d < ngpu; ++d)
        INTEGER(devices)[d] = d;

    SEXP r = fit(Rf_mkString("A * exp(-lam * x) + b"), strvec(3, pn), strvec(1, vn), cols, y, start, w, ci, cd, modes, devices);
    if (r_stub_protect_depth() != 0) {
        fprintf(stderr, "unbalanced PROTECT: %d\n", r_stub_protect_depth());
        return 5;
    }
    SEXP names = Rf_getAttrib(r, R_NamesSymbol);
    printf("{\"names\": [");
    for (int i = 0; i < LENGTH(names); ++i)
        printf("%s\"%s\"", i ? ", " : "", CHAR(STRING_ELT(names, i)));
    SEXP par = VECTOR_ELT(r, 0), covar = VECTOR_ELT(r, 1);
    printf("], \"par\": [%.17g, %.17g, %.17g], \"parnames\": \"%s\"", REAL(par)[0], REAL(par)[1], REAL(par)[2],
           CHAR(STRING_ELT(Rf_getAttrib(par, R_NamesSymbol), 1)));
    printf(", \"covar\": [");
    for (int i = 0; i < 9; ++i)
        printf("%s%.17g", i ? ", " : "", REAL(covar)[i]);
    printf("], \"resid_is_null\": %d, \"grad_is_null\": %d", VECTOR_ELT(r, 2) == R_NilValue, VECTOR_ELT(r, 3) == R_NilValue);
    printf(", \"niter\": %d, \"status\": \"%s\", \"conv\": %d, \"ssr\": %.17g, \"ssrtol\": %.17g, \"algorithm\": \"%s\"",
           INTEGER(VECTOR_ELT(r, 4))[0], CHAR(STRING_ELT(VECTOR_ELT(r, 5), 0)), INTEGER(VECTOR_ELT(r, 6))[0],
           REAL(VECTOR_ELT(r, 7))[0], REAL(VECTOR_ELT(r, 8))[0], CHAR(STRING_ELT(VECTOR_ELT(r, 9), 0)));
    SEXP neval = VECTOR_ELT(r, 10), st = VECTOR_ELT(r, 12), pt = VECTOR_ELT(r, 11);
    printf(", \"neval\": [%d, %d, %d, %d]", INTEGER(neval)[0], INTEGER(neval)[1], INTEGER(neval)[2], INTEGER(neval)[3]);
    printf(", \"ntrace\": %d, \"partrace_dim\": [%d, %d], \"ssrtrace0\": %.17g", LENGTH(st),
           INTEGER(Rf_getAttrib(pt, R_DimSymbol))[0], INTEGER(Rf_getAttrib(pt, R_DimSymbol))[1], REAL(st)[0]);
    /* lazy accessors */
    SEXP e = ev(VECTOR_ELT(r, 14), par, Rf_ScalarLogical(1));
    SEXP resid = VECTOR_ELT(e, 0), grad = VECTOR_ELT(e, 1);
    double ss = 0.0;
    for (int i = 0; i < n; ++i)
        ss += REAL(resid)[i] * REAL(resid)[i];
    printf(", \"resid_ss\": %.17g, \"grad_dim\": [%d, %d], \"grad00\": %.17g, \"grad_last\": %.17g", ss,
           INTEGER(Rf_getAttrib(grad, R_DimSymbol))[0], INTEGER(Rf_getAttrib(grad, R_DimSymbol))[1], REAL(grad)[0],
           REAL(grad)[3 * n - 1]);
    fr(VECTOR_ELT(r, 14));
    printf(", \"released\": %d}\n", R_ExternalPtrAddr(VECTOR_ELT(r, 14)) == NULL);
    r_stub_run_finalizers(); /* a second release through the finalizer must be harmless */
    return 0;
}
