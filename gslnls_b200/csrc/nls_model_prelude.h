// nls_model_prelude.h -- included by every generated model source (see expr.cpp).  Works for
// NVRTC / nvcc device compilation and, for the CPU-side code-generation tests, plain host C++.
#pragma once
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define NLS_FN static __device__ __forceinline__
#define NLS_INF (__longlong_as_double(0x7ff0000000000000LL))
#define NLS_NAN (__longlong_as_double(0x7ff8000000000000LL))
#else
#include <cmath>
using std::exp; using std::log; using std::log2; using std::log10; using std::log1p; using std::expm1;
using std::sqrt; using std::sin; using std::cos; using std::tan; using std::asin; using std::acos;
using std::atan; using std::sinh; using std::cosh; using std::tanh; using std::fabs; using std::pow;
using std::erfc;
#define NLS_FN static inline
#define NLS_INF (HUGE_VAL)
#define NLS_NAN (NAN)
#endif

// x^n for a literal integer n: repeated squaring, unrolled at compile time once n is known
NLS_FN double nls_powi(double a, int n)
{
    const bool inv = n < 0;
    unsigned m = inv ? (unsigned)(-n) : (unsigned)n;
    double r = 1.0, b = a;
    while (m) {
        if (m & 1u)
            r *= b;
        m >>= 1;
        if (m)
            b *= b;
    }
    return inv ? 1.0 / r : r;
}
NLS_FN double nls_sign(double a) { return (double)((a > 0.0) - (a < 0.0)); }
NLS_FN double nls_pnorm(double a) { return 0.5 * erfc(-a * 0.70710678118654752440); }
NLS_FN double nls_dnorm(double a) { return exp(-0.5 * a * a) * 0.39894228040143267794; }
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
NLS_FN double nls_sinpi(double a) { return sinpi(a); }
NLS_FN double nls_cospi(double a) { return cospi(a); }
#else
NLS_FN double nls_sinpi(double a) { return sin(3.14159265358979323846 * a); }
NLS_FN double nls_cospi(double a) { return cos(3.14159265358979323846 * a); }
#endif
