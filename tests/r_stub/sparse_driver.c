/* tests/r_stub/sparse_driver.c -- TEST INFRASTRUCTURE: plays the part of R/nls_large_cuda_sparse.R for the compiled
 * shim r-package/src/nls_large_cuda_sparse.c.  Builds the two blocks of the Penalty function I
 * (inst/unit_tests/unit_tests_gslnls.R:316-346; README Example 4 with p = 500) as the R front-end would, looks
 * C_nls_large_cuda_sparse up in the registration table, fits with cgst and prints the result as JSON.
 * usage: sparse_driver p start(0: rep(0.15, p) | 1: 1:p) maxiter */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

void R_init_gslnlscuda(DllInfo *dll);

static SEXP str1(const char *s) { return Rf_mkString(s); }
static SEXP named(int n, const char **names, SEXP *vals)
{
    SEXP l = Rf_allocVector(VECSXP, n), nm = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; ++i) {
        SET_VECTOR_ELT(l, i, vals[i]);
        SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    }
    Rf_setAttrib(l, R_NamesSymbol, nm);
    return l;
}
static SEXP block(const char *rhs, SEXP index_col, SEXP rows, int row0)
{
    static const char *nm[] = {"rhs", "pnames", "base", "index", "vnames", "cols", "rows", "row0"};
    SEXP index = Rf_allocVector(VECSXP, 1), base = Rf_allocVector(INTSXP, 1);
    SET_VECTOR_ELT(index, 0, index_col);
    INTEGER(base)[0] = 0;
    SEXP vals[] = {str1(rhs), str1("th"), base, index, Rf_allocVector(STRSXP, 0), Rf_allocVector(VECSXP, 0), rows,
                   Rf_ScalarInteger(row0)};
    return named(8, nm, vals);
}

int main(int argc, char **argv)
{
    const int p = argc > 1 ? atoi(argv[1]) : 10, st = argc > 2 ? atoi(argv[2]) : 0, maxiter = argc > 3 ? atoi(argv[3]) : 100;
    DllInfo dll = {NULL};
    R_init_gslnlscuda(&dll);
    typedef SEXP (*fit_fn)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    fit_fn fit = NULL;
    for (const R_CallMethodDef *d = dll.call; d && d->name; ++d)
        if (strcmp(d->name, "C_nls_large_cuda_sparse") == 0 && d->numArgs == 8)
            fit = (fit_fn)d->fun;
    if (!fit) {
        fprintf(stderr, "C_nls_large_cuda_sparse/8 not registered\n");
        return 4;
    }
    SEXP idx = Rf_allocVector(INTSXP, p), rows2 = Rf_allocVector(INTSXP, p), y = Rf_allocVector(REALSXP, p + 1);
    SEXP start = Rf_allocVector(REALSXP, p);
    for (int i = 0; i < p; ++i) {
        INTEGER(idx)[i] = i;
        INTEGER(rows2)[i] = p;
        REAL(y)[i] = 0.0;
        REAL(start)[i] = st ? (double)(i + 1) : 0.15;
    }
    REAL(y)[p] = 0.25;
    char rhs1[64];
    snprintf(rhs1, sizeof rhs1, "%.17g * (th - 1)", sqrt(1e-5));
    SEXP blocks = Rf_allocVector(VECSXP, 2);
    SET_VECTOR_ELT(blocks, 0, block(rhs1, idx, R_NilValue, 0));
    SET_VECTOR_ELT(blocks, 1, block("th^2", idx, rows2, 0));
    SEXP ci = Rf_allocVector(INTSXP, 7), cd = Rf_allocVector(REALSXP, 8), want = Rf_allocVector(INTSXP, 2);
    const int civ[7] = {maxiter, 1, 5, 0, 0, -2, 0};                     /* R/nls_large.R:383-391, algorithm cgst */
    const double cdv[8] = {2.0, 3.0, 0.75, 1.4901161193847656e-08, 0.02, /* R/nls.R:1187-1188 */
                           1.4901161193847656e-08, 1.4901161193847656e-08, 1.4901161193847656e-08};
    memcpy(INTEGER(ci), civ, sizeof civ);
    memcpy(REAL(cd), cdv, sizeof cdv);
    INTEGER(want)[0] = p <= 64;
    INTEGER(want)[1] = 1;
    SEXP r = fit(blocks, y, start, R_NilValue, ci, cd, want, Rf_ScalarInteger(0));
    if (r_stub_protect_depth() != 0) {
        fprintf(stderr, "unbalanced PROTECT: %d\n", r_stub_protect_depth());
        return 5;
    }
    SEXP names = Rf_getAttrib(r, R_NamesSymbol), par = VECTOR_ELT(r, 0), resid = VECTOR_ELT(r, 10);
    printf("{\"names\": [");
    for (int i = 0; i < LENGTH(names); ++i)
        printf("%s\"%s\"", i ? ", " : "", CHAR(STRING_ELT(names, i)));
    printf("], \"par\": [");
    for (int i = 0; i < p; ++i)
        printf("%s%.17g", i ? ", " : "", REAL(par)[i]);
    double ss = 0.0;
    for (int i = 0; i <= p; ++i)
        ss += REAL(resid)[i] * REAL(resid)[i];
    printf("], \"niter\": %d, \"status\": \"%s\", \"conv\": %d, \"ssr\": %.17g, \"resid_ss\": %.17g, \"ntrace\": %d",
           INTEGER(VECTOR_ELT(r, 1))[0], CHAR(STRING_ELT(VECTOR_ELT(r, 2), 0)), INTEGER(VECTOR_ELT(r, 3))[0],
           REAL(VECTOR_ELT(r, 4))[0], ss, LENGTH(VECTOR_ELT(r, 7)));
    printf(", \"jtj_is_null\": %d, \"cg_iters\": %.0f, \"nnz\": %.0f}\n", VECTOR_ELT(r, 9) == R_NilValue,
           REAL(VECTOR_ELT(r, 11))[0], REAL(VECTOR_ELT(r, 12))[0]);
    return 0;
}
