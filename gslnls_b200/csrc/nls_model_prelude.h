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

// exp() of the generated models.  The pass kernels are co-limited by the FP64 pipe (DMMA included: on
// B200 the tensor-core FP64 path issues to the same units), and CUDA's exp() is 16 of the 33 FP64
// instructions per observation of the 3-parameter exponential model.  nls_exp() spends 10: k =
// rint(x 64/ln2), r = x - k ln2/64 (two FMAs), a degree-5 polynomial on |r| <= ln2/128 and one
// table entry 2^(j/64) from shared memory, the power of two added to the exponent field by the
// integer pipe.  Maximum error 1.24 ulp over |x| <= 700 (checked against mpmath); larger |x|,
// infinities and NaN take the library exp().
#ifndef NLS_FAST_EXP
#define NLS_FAST_EXP 0
#endif
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__device__ const unsigned long long nls_exp_tab_bits[64] = {
    0x3ff0000000000000ULL, 0x3ff02c9a3e778061ULL, 0x3ff059b0d3158574ULL, 0x3ff0874518759bc8ULL,
    0x3ff0b5586cf9890fULL, 0x3ff0e3ec32d3d1a2ULL, 0x3ff11301d0125b51ULL, 0x3ff1429aaea92de0ULL,
    0x3ff172b83c7d517bULL, 0x3ff1a35beb6fcb75ULL, 0x3ff1d4873168b9aaULL, 0x3ff2063b88628cd6ULL,
    0x3ff2387a6e756238ULL, 0x3ff26b4565e27cddULL, 0x3ff29e9df51fdee1ULL, 0x3ff2d285a6e4030bULL,
    0x3ff306fe0a31b715ULL, 0x3ff33c08b26416ffULL, 0x3ff371a7373aa9cbULL, 0x3ff3a7db34e59ff7ULL,
    0x3ff3dea64c123422ULL, 0x3ff4160a21f72e2aULL, 0x3ff44e086061892dULL, 0x3ff486a2b5c13cd0ULL,
    0x3ff4bfdad5362a27ULL, 0x3ff4f9b2769d2ca7ULL, 0x3ff5342b569d4f82ULL, 0x3ff56f4736b527daULL,
    0x3ff5ab07dd485429ULL, 0x3ff5e76f15ad2148ULL, 0x3ff6247eb03a5585ULL, 0x3ff6623882552225ULL,
    0x3ff6a09e667f3bcdULL, 0x3ff6dfb23c651a2fULL, 0x3ff71f75e8ec5f74ULL, 0x3ff75feb564267c9ULL,
    0x3ff7a11473eb0187ULL, 0x3ff7e2f336cf4e62ULL, 0x3ff82589994cce13ULL, 0x3ff868d99b4492edULL,
    0x3ff8ace5422aa0dbULL, 0x3ff8f1ae99157736ULL, 0x3ff93737b0cdc5e5ULL, 0x3ff97d829fde4e50ULL,
    0x3ff9c49182a3f090ULL, 0x3ffa0c667b5de565ULL, 0x3ffa5503b23e255dULL, 0x3ffa9e6b5579fdbfULL,
    0x3ffae89f995ad3adULL, 0x3ffb33a2b84f15fbULL, 0x3ffb7f76f2fb5e47ULL, 0x3ffbcc1e904bc1d2ULL,
    0x3ffc199bdd85529cULL, 0x3ffc67f12e57d14bULL, 0x3ffcb720dcef9069ULL, 0x3ffd072d4a07897cULL,
    0x3ffd5818dcfba487ULL, 0x3ffda9e603db3285ULL, 0x3ffdfc97337b9b5fULL, 0x3ffe502ee78b3ff6ULL,
    0x3ffea4afa2a490daULL, 0x3ffefa1bee615a27ULL, 0x3fff50765b6e4540ULL, 0x3fffa7c1819e90d8ULL,
};
static __device__ __forceinline__ double *nls_exp_tab()
{
    __shared__ double tab[64];
    return tab;
}
// every kernel that evaluates a model calls this once, with all threads of the CTA
static __device__ __forceinline__ void nls_exp_init()
{
#if NLS_FAST_EXP == 1
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
        nls_exp_tab()[i] = __longlong_as_double((long long)nls_exp_tab_bits[i]);
    __syncthreads();
#endif
}
static __device__ __forceinline__ int nls_selp(int a, int b, bool c)
{
    int r;
    asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.b32 %0, %1, %2, p;\n}" : "=r"(r) : "r"(a), "r"(b), "r"((int)c));
    return r;
}
// Branch-free: a branch inside exp() ends the basic block, and the instruction scheduler can then no
// longer interleave the independent exp() chains of neighbouring observations (or of the 16 terms of
// a Gaussian mixture), which leaves them latency-bound.  |x| >= 708, infinities and NaN are resolved
// by integer selects on the result words: exp(x >= 708) = +Inf (the true value is finite up to
// 709.78, but such a model value overflows the sum of squares anyway), exp(x <= -708) = 0 (true value
// below 3.4e-308), NaN -> +Inf or NaN bit pattern with the sign cleared (non-finite either way, which
// is all the residual rule of src/nls_large.c:464-465 looks at).
NLS_FN double nls_exp(double x)
{
    const int hi = __double2hiint(x);
    const int ahi = hi & 0x7fffffff;
    const bool special = ahi >= 0x40862000;                     // |x| >= 708, Inf, NaN
    const bool to_zero = (hi < 0) && (ahi <= 0x7ff00000);       // large negative, -Inf
#if NLS_FAST_EXP == 2
    // polynomial only: k = rint(x / ln 2), degree-11 near-minimax polynomial on |r| <= ln2 / 2 (0.91 ulp)
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -0.6931471805599453, x);
    r = fma(kf, -2.3190468138462996e-17, r);
    double q = fma(2.5110049204818658e-08, r, 2.763265472252779e-07);
    q = fma(q, r, 2.755724088722987e-06);
    q = fma(q, r, 2.4801485441561313e-05);
    q = fma(q, r, 0.00019841269890076403);
    q = fma(q, r, 0.0013888888952352863);
    q = fma(q, r, 0.008333333333319589);
    q = fma(q, r, 0.04166666666648795);
    q = fma(q, r, 0.1666666666666668);
    q = fma(q, r, 0.5000000000000019);
    q = fma(q, r, 1.0);
    const double res = fma(q, r, 1.0);
    const int m = k;
#else
    // table: k = rint(x 64 / ln 2), degree-5 polynomial on |r| <= ln2 / 128, 2^(j/64) from shared memory
    // (1.24 ulp)
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -0.010830424696249145, x);
    r = fma(kf, -3.623510646634843e-19, r);
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    const double r2 = r * r;
    const double pr = fma(q, r2, r);
    const double T = nls_exp_tab()[k & 63];
    const double res = fma(T, pr, T);
    const int m = k >> 6;
#endif
    // opaque selects: written as a C conditional the compiler would branch around the arithmetic above
    const int shi = to_zero ? 0 : (ahi > 0x7ff00000 ? ahi : 0x7ff00000);
    const int ohi = nls_selp(shi, __double2hiint(res) + (m << 20), special);
    const int olo = nls_selp(0, __double2loint(res), special);
    return __hiloint2double(ohi, olo);
}
#else
NLS_FN double nls_exp(double x) { return exp(x); }
#endif
// NLS_FAST_EXP: 0 library exp(), 1 table variant, 2 polynomial variant of nls_exp()
#ifndef NLS_FAST_EXP
#define NLS_FAST_EXP 0
#endif
#if NLS_FAST_EXP
#define NLS_EXP(x) nls_exp(x)
#else
#define NLS_EXP(x) exp(x)
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
