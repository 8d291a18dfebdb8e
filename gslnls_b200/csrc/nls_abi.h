// nls_abi.h -- plain-C structs shared by the host library (nvcc/g++) and the NVRTC-compiled
// kernels.  Keep it free of includes so that NVRTC can consume it as an in-memory header.
#pragma once

#define NLS_MAX_VARS 16

// Parameters of one fused pass (K1).  Passed by value to the kernel.
struct NlsPassParams {
    const double *vars[NLS_MAX_VARS]; // nvar predictor columns, n doubles each (device)
    const double *y;                  // responses
    const double *w;                  // raw weights or nullptr
    long long n;                      // local observations
    const double *req;                // [ncand][req_stride]: mode, theta[p], v[p]
    double *partials;                 // [ncand][gridDim.x][pk_stride] per-CTA partial packets
    double *packet;                   // [ncand][pk_stride] reduced packet
    unsigned int *ticket;             // [ncand] arrival counters (zero between launches)
    double h_df, h_fvv;               // finite-difference steps (control_dbl[3], [4])
    int req_stride, pk_stride;
    int force_mode;                   // >0: ignore req[0] and run this mode (test hooks)
    int pad_;
};

struct NlsMaterialiseParams {
    const double *vars[NLS_MAX_VARS];
    const double *y;
    const double *w;
    long long n;
    const double *theta;  // p doubles (device)
    double *resid;        // n or nullptr
    double *grad;         // n*p column-major or nullptr
    double h_df;
};
