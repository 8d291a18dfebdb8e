/* nls_large_cuda_sparse.c -- .Call shim for gsl_nls_large() with a SPARSE Jacobian on the GPU.  Logic-free:
 * SEXP <-> gslnls_sparse_* (include/gslnls_b200.h) only.
 *
 * The reference receives the sparsity from an R closure returning a dgC/dgR/dgTMatrix (R/nls_large.R:397-404) and
 * rebuilds a gsl_spmatrix from it on every callback (src/nls_large.c:528-623).  A closure cannot run on a GPU, so the
 * structure crosses the boundary as data: `blocks` is a list of
 *   list(rhs = <chr 1>, pnames = <chr k>, base = <int k>, index = <list k of int nterms | NULL>,
 *        vnames = <chr v>, cols = <list v of dbl nterms>, rows = <int nterms | NULL>, row0 = <int 1>)
 * (row formula, where each of its parameters lives in the global vector, its data columns, the row of every term);
 * y, start, weights, control_int[7], control_dbl[8] cross as in C_nls_large (src/nls_large.c:66-75). */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "gslnls_b200.h"

#define SP_MAX_BLOCKS 64

static SEXP list_elt(SEXP list, const char *name)
{
    SEXP nms = Rf_getAttrib(list, R_NamesSymbol);
    for (int i = 0; i < LENGTH(list); ++i)
        if (strcmp(CHAR(STRING_ELT(nms, i)), name) == 0)
            return VECTOR_ELT(list, i);
    return R_NilValue;
}

SEXP C_nls_large_cuda_sparse(SEXP blocks, SEXP y, SEXP start, SEXP weights, SEXP control_int, SEXP control_dbl,
                             SEXP want, SEXP device)
{
    static const char *names[] = {"par", "niter", "status", "conv", "ssr", "ssrtol", "neval", "ssrtrace",
                                  "grad_vec", "jtj", "resid", "cg_iters", "nnz"};
    const int p = LENGTH(start), nb = LENGTH(blocks), *wt = INTEGER(want); /* want = c(jtj, resid) */
    const R_xlen_t nrows = XLENGTH(y);
    gslnls_model *models[SP_MAX_BLOCKS] = {0};
    gslnls_sparse_problem *sp = NULL;
    gslnls_sparse_result res;
    char err[4096] = "";
    memset(&res, 0, sizeof res);
    if (nb < 1 || nb > SP_MAX_BLOCKS)
        Rf_error("gsl_nls_large (CUDA, sparse): between 1 and %d blocks", SP_MAX_BLOCKS);
    int rc = gslnls_sparse_create(INTEGER(device)[0], p, (int64_t)nrows, &sp);
    for (int b = 0; b < nb && rc == GSLNLS_SUCCESS; ++b) {
        SEXP blk = VECTOR_ELT(blocks, b), pnames = list_elt(blk, "pnames"), vnames = list_elt(blk, "vnames");
        SEXP index = list_elt(blk, "index"), cols = list_elt(blk, "cols"), rows = list_elt(blk, "rows");
        const int k = LENGTH(pnames), nvar = LENGTH(vnames);
        const char **pn = (const char **)R_alloc(k, sizeof(char *));
        const char **vn = (const char **)R_alloc(nvar > 0 ? nvar : 1, sizeof(char *));
        const double **xs = (const double **)R_alloc(nvar > 0 ? nvar : 1, sizeof(double *));
        const int **ix = (const int **)R_alloc(k, sizeof(int *));
        R_xlen_t nterms = rows != R_NilValue ? XLENGTH(rows) : 0;
        for (int s = 0; s < k; ++s) {
            SEXP col = VECTOR_ELT(index, s);
            pn[s] = CHAR(STRING_ELT(pnames, s));
            ix[s] = col != R_NilValue ? INTEGER(col) : NULL;
            if (col != R_NilValue)
                nterms = XLENGTH(col);
        }
        for (int v = 0; v < nvar; ++v) {
            vn[v] = CHAR(STRING_ELT(vnames, v));
            xs[v] = REAL(VECTOR_ELT(cols, v));
            nterms = XLENGTH(VECTOR_ELT(cols, v));
        }
        rc = gslnls_model_compile(CHAR(STRING_ELT(list_elt(blk, "rhs"), 0)), pn, k, vn, nvar, GSLNLS_JAC_SYMBOLIC, 0,
                                  &models[b], err, sizeof err);
        if (rc == GSLNLS_SUCCESS)
            rc = gslnls_sparse_add_block(sp, models[b], (int64_t)nterms, xs, INTEGER(list_elt(blk, "base")), ix,
                                         rows != R_NilValue ? INTEGER(rows) : NULL,
                                         (int64_t)INTEGER(list_elt(blk, "row0"))[0]);
    }
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_sparse_set_response(sp, REAL(y), weights != R_NilValue ? REAL(weights) : NULL);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_sparse_finalize(sp);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_sparse_fit(sp, REAL(start), INTEGER(control_int), REAL(control_dbl), wt[0], wt[1], &res);
    if (rc >= 1000 || rc == GSLNLS_EINVAL) { /* library-level failure: nothing to return */
        strncpy(err, err[0] ? err : gslnls_last_error(), sizeof err - 1);
        gslnls_sparse_result_free(&res);
        gslnls_sparse_free(sp);
        for (int b = 0; b < nb; ++b)
            gslnls_model_free(models[b]);
        Rf_error("gsl_nls_large (CUDA, sparse): %s", err);
    }
    SEXP ans = PROTECT(Rf_allocVector(VECSXP, 13)), nms = PROTECT(Rf_allocVector(STRSXP, 13));
    for (int i = 0; i < 13; ++i)
        SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
    Rf_setAttrib(ans, R_NamesSymbol, nms);
    SEXP par = PROTECT(Rf_allocVector(REALSXP, p)), gv = PROTECT(Rf_allocVector(REALSXP, p));
    SEXP neval = PROTECT(Rf_allocVector(INTSXP, 4));
    memcpy(REAL(par), res.par, sizeof(double) * (size_t)p);
    memcpy(REAL(gv), res.grad_vec, sizeof(double) * (size_t)p);
    for (int i = 0; i < 4; ++i)
        INTEGER(neval)[i] = (int)res.neval[i];
    SET_VECTOR_ELT(ans, 0, par);
    SET_VECTOR_ELT(ans, 1, Rf_ScalarInteger(res.niter));
    SET_VECTOR_ELT(ans, 2, Rf_mkString(res.status));
    SET_VECTOR_ELT(ans, 3, Rf_ScalarInteger(res.conv));
    SET_VECTOR_ELT(ans, 4, Rf_ScalarReal(res.ssr));
    SET_VECTOR_ELT(ans, 5, Rf_ScalarReal(res.ssrtol));
    SET_VECTOR_ELT(ans, 6, neval);
    if (res.ntrace > 0) {
        SEXP st = PROTECT(Rf_allocVector(REALSXP, res.niter + 1));
        memcpy(REAL(st), res.ssrtrace, sizeof(double) * (size_t)(res.niter + 1));
        SET_VECTOR_ELT(ans, 7, st);
        UNPROTECT(1);
    }
    SET_VECTOR_ELT(ans, 8, gv);
    if (res.jtj) {
        SEXP m = PROTECT(Rf_allocMatrix(REALSXP, p, p));
        memcpy(REAL(m), res.jtj, sizeof(double) * (size_t)p * (size_t)p);
        SET_VECTOR_ELT(ans, 9, m);
        UNPROTECT(1);
    }
    if (res.resid) {
        SEXP r = PROTECT(Rf_allocVector(REALSXP, nrows));
        memcpy(REAL(r), res.resid, sizeof(double) * (size_t)nrows);
        SET_VECTOR_ELT(ans, 10, r);
        UNPROTECT(1);
    }
    SET_VECTOR_ELT(ans, 11, Rf_ScalarReal((double)res.cg_iters));
    SET_VECTOR_ELT(ans, 12, Rf_ScalarReal((double)res.nnz));
    gslnls_sparse_result_free(&res);
    gslnls_sparse_free(sp); /* before the models: the problem's kernels belong to them */
    for (int b = 0; b < nb; ++b)
        gslnls_model_free(models[b]);
    UNPROTECT(5);
    return ans;
}
