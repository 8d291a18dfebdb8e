/* init.c -- native routine registration of the CUDA shim (the reference registers C_nls_large with 9 arguments
 * at src/init.c:9,16; these are the entries a maintainer adds next to it) */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

extern SEXP C_nls_large_cuda(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP C_nls_large_cuda_eval(SEXP, SEXP, SEXP);
extern SEXP C_nls_large_cuda_free(SEXP);
extern SEXP C_nls_large_cuda_sparse(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static const R_CallMethodDef CallEntries[] = {
    {"C_nls_large_cuda", (DL_FUNC)&C_nls_large_cuda, 11},
    {"C_nls_large_cuda_eval", (DL_FUNC)&C_nls_large_cuda_eval, 3},
    {"C_nls_large_cuda_free", (DL_FUNC)&C_nls_large_cuda_free, 1},
    {"C_nls_large_cuda_sparse", (DL_FUNC)&C_nls_large_cuda_sparse, 8},
    {NULL, NULL, 0}};

void R_init_gslnlscuda(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
