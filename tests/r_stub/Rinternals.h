/* tests/r_stub/Rinternals.h -- TEST INFRASTRUCTURE: the slice of R's C API that the files under r-package/src use, so the
 * .Call shim is compiled (and, on the GPU box, executed) by the test-suite although R is not installed here.
 * Names and signatures follow R's public headers; the implementation (r_stub.c) is a toy heap that never frees. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>

typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
typedef struct SEXPREC *SEXP;
enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, CHARSXP = 9, EXTPTRSXP = 22 };

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol, R_DimNamesSymbol;

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkChar(const char *s);
SEXP Rf_mkString(const char *s);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
SEXP Rf_getAttrib(SEXP x, SEXP name);
int Rf_asLogical(SEXP x);
int LENGTH(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int TYPEOF(SEXP x);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
const char *CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP PROTECT(SEXP x);
void UNPROTECT(int n);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
void Rf_warning(const char *fmt, ...);
char *R_alloc(size_t n, int size);
/* test helpers (not part of R) */
void r_stub_run_finalizers(void);
int r_stub_protect_depth(void);
#endif
