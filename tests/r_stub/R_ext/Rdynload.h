/* tests/r_stub/R_ext/Rdynload.h -- TEST INFRASTRUCTURE, see ../Rinternals.h */
#ifndef R_STUB_RDYNLOAD_H
#define R_STUB_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const R_CallMethodDef *call; } DllInfo;
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e);
int R_useDynamicSymbols(DllInfo *info, int value);
#endif
