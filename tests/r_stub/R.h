/* tests/r_stub/R.h -- TEST INFRASTRUCTURE, see Rinternals.h */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stdlib.h>
#define R_Calloc(n, t) ((t *)calloc((size_t)(n), sizeof(t)))
#define R_Free(p) (free((void *)(p)), (p) = NULL)
#endif
