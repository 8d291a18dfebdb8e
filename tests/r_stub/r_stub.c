/* tests/r_stub/r_stub.c -- TEST INFRASTRUCTURE: toy implementation of the R API slice in Rinternals.h */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
    int type;
    R_xlen_t len;
    void *data;
    SEXP names, dim, dimnames;
    R_CFinalizer_t fin;
};
static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC sym_names, sym_dim, sym_dimnames;
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &sym_names, R_DimSymbol = &sym_dim, R_DimNamesSymbol = &sym_dimnames;
static int protect_depth = 0;
static SEXP finalizable[64];
static int nfinalizable = 0;

static SEXP node(int type, R_xlen_t n, size_t elt)
{
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = type;
    s->len = n;
    s->data = calloc((size_t)(n > 0 ? n : 1), elt);
    s->names = s->dim = s->dimnames = R_NilValue;
    return s;
}
SEXP Rf_allocVector(int type, R_xlen_t n)
{
    switch (type) {
    case REALSXP: return node(type, n, sizeof(double));
    case INTSXP: case LGLSXP: return node(type, n, sizeof(int));
    case STRSXP: case VECSXP: {
        SEXP s = node(type, n, sizeof(SEXP));
        for (R_xlen_t i = 0; i < n; ++i)
            ((SEXP *)s->data)[i] = R_NilValue;
        return s;
    }
    default: Rf_error("r_stub: unsupported type %d", type);
    }
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol), d = Rf_allocVector(INTSXP, 2);
    INTEGER(d)[0] = nrow;
    INTEGER(d)[1] = ncol;
    s->dim = d;
    return s;
}
SEXP Rf_mkChar(const char *str)
{
    SEXP s = node(CHARSXP, (R_xlen_t)strlen(str), 1);
    free(s->data);
    s->data = strdup(str);
    return s;
}
SEXP Rf_mkString(const char *str)
{
    SEXP s = Rf_allocVector(STRSXP, 1);
    SET_STRING_ELT(s, 0, Rf_mkChar(str));
    return s;
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val)
{
    if (name == R_NamesSymbol) x->names = val;
    else if (name == R_DimSymbol) x->dim = val;
    else if (name == R_DimNamesSymbol) x->dimnames = val;
    return val;
}
SEXP Rf_getAttrib(SEXP x, SEXP name)
{
    return name == R_NamesSymbol ? x->names : name == R_DimSymbol ? x->dim : name == R_DimNamesSymbol ? x->dimnames : R_NilValue;
}
int Rf_asLogical(SEXP x) { return x->len > 0 ? ((int *)x->data)[0] != 0 : 0; }
int LENGTH(SEXP x) { return (int)x->len; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int TYPEOF(SEXP x) { return x->type; }
double *REAL(SEXP x) { if (x->type != REALSXP) Rf_error("r_stub: REAL() of a non-double"); return (double *)x->data; }
int *INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) Rf_error("r_stub: INTEGER() of a non-integer"); return (int *)x->data; }
int *LOGICAL(SEXP x) { return INTEGER(x); }
const char *CHAR(SEXP x) { return (const char *)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
SEXP PROTECT(SEXP x) { ++protect_depth; return x; }
void UNPROTECT(int n) { protect_depth -= n; if (protect_depth < 0) Rf_error("r_stub: protect stack underflow"); }
int r_stub_protect_depth(void) { return protect_depth; }
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = node(EXTPTRSXP, 0, 1);
    free(s->data);
    s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit)
{
    (void)onexit;
    s->fin = fun;
    if (nfinalizable < 64)
        finalizable[nfinalizable++] = s;
}
void r_stub_run_finalizers(void)
{
    for (int i = 0; i < nfinalizable; ++i)
        if (finalizable[i]->fin)
            finalizable[i]->fin(finalizable[i]);
    nfinalizable = 0;
}
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(3); /* R would longjmp to top level; the driver treats it as "the call raised an R error" */
}
void Rf_warning(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Warning: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}
char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }
static const R_CallMethodDef *registered = NULL;
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)c; (void)f; (void)e;
    registered = call;
    if (info) info->call = call;
    return 1;
}
int R_useDynamicSymbols(DllInfo *info, int value) { (void)info; (void)value; return 1; }
