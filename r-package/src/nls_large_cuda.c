/* nls_large_cuda.c -- the .Call shim between R and libgslnls_b200.so.  Logic-free: SEXP <-> C ABI only.
 *
 * Replaces, for the GPU path, the registered routine {"C_nls_large", (DL_FUNC)&C_nls_large, 9} of the
 * reference (src/init.c:9,16; src/nls_large.c:66-75, called from R/nls_large.R:411 and :598).  The closures
 * (fn, jac, fvv, env) cannot run on a GPU, so the model crosses the boundary as text (formula right-hand side,
 * parameter names, predictor names + columns); y, start, weights, control_int[7], control_dbl[8] cross unchanged
 * and the returned list has the reference's names (src/nls_large.c:279-288).  `resid` and `grad` are NOT built by
 * the fit: the data stays resident behind the external pointer `handle` and C_nls_large_cuda_eval produces them
 * when R asks (lazy post-fit accessors, R/nls_large_cuda.R).
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "gslnls_b200.h"

typedef struct {
    gslnls_model *model;
    gslnls_session *session;
    int p;
    R_xlen_t n;
} cuda_fit;

static void cuda_fit_release(cuda_fit *h)
{
    if (!h)
        return;
    gslnls_session_free(h->session); /* before the model: the session's kernels belong to it */
    gslnls_model_free(h->model);
    R_Free(h);
}

static void cuda_fit_finalizer(SEXP ptr)
{
    cuda_fit_release((cuda_fit *)R_ExternalPtrAddr(ptr));
    R_ClearExternalPtr(ptr);
}

static SEXP named_list(int len, const char **names)
{
    SEXP ans = PROTECT(Rf_allocVector(VECSXP, len)), nms = PROTECT(Rf_allocVector(STRSXP, len));
    for (int i = 0; i < len; ++i)
        SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
    Rf_setAttrib(ans, R_NamesSymbol, nms);
    UNPROTECT(2);
    return ans;
}

static SEXP real_matrix(const double *src, R_xlen_t nr, int nc, SEXP colnames, int both)
{
    SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)nr, nc));
    if (src)
        memcpy(REAL(m), src, sizeof(double) * (size_t)nr * (size_t)nc);
    if (colnames != R_NilValue) {
        SEXP dn = PROTECT(Rf_allocVector(VECSXP, 2));
        SET_VECTOR_ELT(dn, 0, both ? colnames : R_NilValue);
        SET_VECTOR_ELT(dn, 1, colnames);
        Rf_setAttrib(m, R_DimNamesSymbol, dn);
        UNPROTECT(1);
    }
    UNPROTECT(1);
    return m;
}

/* modes = c(jac_mode, fvv_mode, weights_mode): GSLNLS_JAC_*, GSLNLS_FVV_*, GSLNLS_WEIGHTS_* */
SEXP C_nls_large_cuda(SEXP rhs, SEXP pnames, SEXP vnames, SEXP cols, SEXP y, SEXP start, SEXP weights,
                      SEXP control_int, SEXP control_dbl, SEXP modes, SEXP devices)
{
    static const char *names[] = {"par", "covar", "resid", "grad", "niter", "status", "conv", "ssr", "ssrtol",
                                  "algorithm", "neval", "partrace", "ssrtrace", "jtj", "handle"};
    const int p = LENGTH(pnames), nvar = LENGTH(vnames), *md = INTEGER(modes);
    const R_xlen_t n = XLENGTH(y);
    const char **pn = (const char **)R_alloc(p > 0 ? p : 1, sizeof(char *));
    const char **vn = (const char **)R_alloc(nvar > 0 ? nvar : 1, sizeof(char *));
    const double **xs = (const double **)R_alloc(nvar > 0 ? nvar : 1, sizeof(double *));
    char err[4096] = "";
    for (int j = 0; j < p; ++j)
        pn[j] = CHAR(STRING_ELT(pnames, j));
    for (int k = 0; k < nvar; ++k) {
        vn[k] = CHAR(STRING_ELT(vnames, k));
        xs[k] = REAL(VECTOR_ELT(cols, k));
    }
    cuda_fit *h = R_Calloc(1, cuda_fit);
    h->p = p;
    h->n = n;
    int rc = gslnls_model_compile(CHAR(STRING_ELT(rhs, 0)), pn, p, vn, nvar, md[0], md[1], &h->model, err, sizeof err);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_session_create(h->model, (int64_t)n, weights != R_NilValue, LENGTH(devices), INTEGER(devices),
                                   &h->session);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_session_set_weights_mode(h->session, md[2]);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_session_upload(h->session, xs, REAL(y), weights != R_NilValue ? REAL(weights) : NULL);
    gslnls_result res;
    memset(&res, 0, sizeof res);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_session_fit(h->session, REAL(start), INTEGER(control_int), REAL(control_dbl), 0, &res);
    if (rc >= 1000 || rc == GSLNLS_EINVAL) { /* library-level failure: nothing to return */
        strncpy(err, err[0] ? err : gslnls_last_error(), sizeof err - 1);
        gslnls_result_free(&res);
        cuda_fit_release(h);
        Rf_error("gsl_nls_large (CUDA): %s", err);
    }
    SEXP ans = PROTECT(named_list(15, names));
    SEXP par = PROTECT(Rf_allocVector(REALSXP, p));
    memcpy(REAL(par), res.par, sizeof(double) * (size_t)p);
    Rf_setAttrib(par, R_NamesSymbol, pnames);
    SET_VECTOR_ELT(ans, 0, par);
    SET_VECTOR_ELT(ans, 1, real_matrix(res.covar, p, p, pnames, 1));
    SET_VECTOR_ELT(ans, 2, R_NilValue); /* resid: lazy, C_nls_large_cuda_eval */
    SET_VECTOR_ELT(ans, 3, R_NilValue); /* grad:  lazy */
    SET_VECTOR_ELT(ans, 4, Rf_ScalarInteger(res.niter));
    SET_VECTOR_ELT(ans, 5, Rf_mkString(res.status));
    SET_VECTOR_ELT(ans, 6, Rf_ScalarInteger(res.conv));
    SET_VECTOR_ELT(ans, 7, Rf_ScalarReal(res.ssr));
    SET_VECTOR_ELT(ans, 8, Rf_ScalarReal(res.ssrtol));
    SET_VECTOR_ELT(ans, 9, Rf_mkString(res.algorithm));
    static const char *ev[] = {"f", "dfu", "df2", "fvv"};
    SEXP neval = PROTECT(Rf_allocVector(INTSXP, 4)), evn = PROTECT(Rf_allocVector(STRSXP, 4));
    for (int i = 0; i < 4; ++i) {
        INTEGER(neval)[i] = (int)res.neval[i];
        SET_STRING_ELT(evn, i, Rf_mkChar(ev[i]));
    }
    Rf_setAttrib(neval, R_NamesSymbol, evn);
    SET_VECTOR_ELT(ans, 10, neval);
    if (res.ntrace > 0) {
        SET_VECTOR_ELT(ans, 11, real_matrix(res.partrace, res.ntrace, p, pnames, 0));
        SEXP st = PROTECT(Rf_allocVector(REALSXP, res.ntrace));
        memcpy(REAL(st), res.ssrtrace, sizeof(double) * (size_t)res.ntrace);
        SET_VECTOR_ELT(ans, 12, st);
        UNPROTECT(1);
    }
    SET_VECTOR_ELT(ans, 13, real_matrix(res.jtj, p, p, pnames, 1));
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, cuda_fit_finalizer, TRUE);
    SET_VECTOR_ELT(ans, 14, ptr);
    gslnls_result_free(&res);
    UNPROTECT(5);
    return ans;
}

/* list(resid = f - y (weighted, n), grad = n x p Jacobian or NULL) at `par` from the resident data:
 * the arrays of src/nls_large.c:339-385, on demand */
SEXP C_nls_large_cuda_eval(SEXP handle, SEXP par, SEXP want_grad)
{
    static const char *names[] = {"resid", "grad"};
    cuda_fit *h = (cuda_fit *)R_ExternalPtrAddr(handle);
    if (!h || !h->session)
        Rf_error("gsl_nls_large (CUDA): the device data of this fit has been released");
    if (LENGTH(par) != h->p)
        Rf_error("gsl_nls_large (CUDA): 'par' must have length %d", h->p);
    const int wg = Rf_asLogical(want_grad) == TRUE;
    SEXP ans = PROTECT(named_list(2, names));
    SEXP resid = PROTECT(Rf_allocVector(REALSXP, h->n));
    SEXP grad = PROTECT(wg ? Rf_allocMatrix(REALSXP, (int)h->n, h->p) : R_NilValue);
    const int rc = gslnls_session_residuals(h->session, REAL(par), REAL(resid), wg ? REAL(grad) : NULL);
    if (rc)
        Rf_error("gsl_nls_large (CUDA): %s", gslnls_last_error());
    SET_VECTOR_ELT(ans, 0, resid);
    SET_VECTOR_ELT(ans, 1, grad);
    UNPROTECT(3);
    return ans;
}

/* release the device memory now instead of at garbage collection */
SEXP C_nls_large_cuda_free(SEXP handle)
{
    cuda_fit_finalizer(handle);
    return R_NilValue;
}
