// nls_pass_kernel.cuh -- K1: the fused residual / Jacobian / normal-equation pass, and K4: the
// final materialisation of residuals and Jacobian.  sm_100a, FP64.
//
// This text is compiled at run time by NVRTC, appended to the model source that expr.cpp generates
// (which defines GSLNLS_P, GSLNLS_NVAR, GSLNLS_JAC_MODE, GSLNLS_FVV_MODE and nls_model_*), so the
// model is inlined into the loop and a Jacobian row only ever exists in registers.
//
// What one launch replaces in the reference (per evaluation at a parameter vector theta):
//   gsl_f_large   src/nls_large.c:426-472  -> f_i = fn(theta)_i - y_i, non-finite fn -> +Inf
//   gsl_df_large  src/nls_large.c:474-653  -> J (n x p), NaN scan, transposing copy,
//                                            cblas_dgemv  g = J^T f   (:629)
//                                            cblas_dsyrk  J^T J lower (:633)
//   gsl_blas_ddot f.f (src/nls_fit.c:178)
//   gsl_fvv_large src/nls_large.c:655-713 + second df call for J^T fvv (mode 2)
// None of f, J, fvv is written to memory: the kernel reads 8*(nvar+1) (+8 with weights) bytes per
// observation and writes one packet per CTA.
//
// Determinism: fixed grid, fixed per-thread observation sets, fixed-shape shuffle trees, CTA
// partials summed in CTA order by whichever CTA arrives last.  No floating-point atomics.
//
// Tunables (NVRTC -D): NLS_BLOCK threads per CTA, NLS_UNROLL independent loads in flight per
// thread and array, NLS_MINB CTAs per SM for __launch_bounds__, NLS_VEC 2 = 16-byte loads (needs
// 16-byte aligned columns) or 1, NLS_HAS_W weights present, NLS_STREAM 1 = L1::no_allocate loads,
// NLS_PREFETCH software-pipelined loads, NLS_TILED 2 + NLS_STAGES the TMA ring, NLS_FAST_EXP the exp() flavour.
#include "nls_abi.h"

#ifndef NLS_BLOCK
#define NLS_BLOCK 256
#endif
#ifndef NLS_UNROLL
#define NLS_UNROLL 4
#endif
#ifndef NLS_MINB
#define NLS_MINB 2
#endif
#ifndef NLS_VEC
#define NLS_VEC 2
#endif
#ifndef NLS_HAS_W
#define NLS_HAS_W 0
#endif
#ifndef NLS_W_GSL
#define NLS_W_GSL 0 /* 1: weights the way GSL's multilarge applies them (reference-compatible): sqrt(w) scales
                       f and fvv only, the Jacobian rows -- hence J^T J -- stay unweighted.  0: rows of J are
                       scaled too, so that g = J^T W f and J^T W J belong to the same weighted problem. */
#endif
#define NLS_W_ROWS (NLS_HAS_W && !NLS_W_GSL) /* Jacobian rows carry sqrt(w) */
#ifndef NLS_STREAM
#define NLS_STREAM 1
#endif
#ifndef NLS_PREFETCH
#define NLS_PREFETCH 0 /* 1: register kernel issues the next trip's loads before this trip's arithmetic */
#endif
#ifndef NLS_TILED
#define NLS_TILED 0 /* 1: shared-memory J tiles + FP64 DMMA SYRK (nls_pass_tiled.cuh), for p > 8;
                       2: register accumulators fed by a TMA bulk-copy shared-memory pipeline */
#endif

#define NLS_P GSLNLS_P
#define NLS_NV (GSLNLS_NVAR > 0 ? GSLNLS_NVAR : 1)
#define NLS_NPK (NLS_P * (NLS_P + 1) / 2)
// packet slots: [J^T J lower | J^T f | f^T f | number of non-finite residuals]
#define NLS_PK (NLS_NPK + NLS_P + 2)
#define NLS_NW (NLS_BLOCK / 32)

enum { NLS_MODE_IDLE = 0, NLS_MODE_FJ = 1, NLS_MODE_FVV = 2, NLS_MODE_JVP = 3 };

// ------------------------------------------------------------------------------------ loads
// L2 eviction policies.  A pass reads each byte once, but the NEXT pass reads the same bytes again and 126 MB of
// L2 sit in front of HBM, so round 2 tried to keep the head of a shard resident: rows below prm.keep_rows
// loaded evict_last, the rest evict_first, on both load paths, with and without a persisting set-aside
// (cudaLimitPersistingL2CacheSize, 79 MB max on B200).  Measured on 100 / 200 / 400 MB shards: no effect
// (profiles/r02_summary.md), so keep_rows defaults to 0 and only the bulk-copy producer still selects a policy
// (one scalar select per tile; evict_first keeps the trust-region server's lines in L2).  The per-thread LDG
// loads carry no L2 hint: a per-lane policy operand makes ptxas wrap every load in a warp-collective loop.
struct NlsPolicy {
    unsigned long long keep, stream;
    long long keep_rows;
};
static __device__ __forceinline__ NlsPolicy nls_policy(long long keep_rows)
{
    NlsPolicy P;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(P.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(P.stream));
    P.keep_rows = keep_rows;
    return P;
}
static __device__ __forceinline__ double2 nls_ld2(const double *p)
{
    double2 r;
#if NLS_STREAM
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
#else
    asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
#endif
    return r;
}
#define NLS_LD2(ptr, o) nls_ld2((ptr) + (o))
static __device__ __forceinline__ double nls_ld1(const double *p)
{
    double r;
#if NLS_STREAM
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
#else
    asm("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
#endif
    return r;
}

static __device__ __forceinline__ bool nls_finite(double v)
{
    // exponent field all ones <=> Inf or NaN; integer test keeps the FP64 pipe free
    return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000;
}

// ------------------------------------------------------------------------------------ channel words
static __device__ __forceinline__ unsigned long long nls_ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void nls_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
static __device__ __forceinline__ void nls_st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long nls_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// developer hook (gslnls_problem_trace): phase stamps of every CTA -- 0 entry, 1 request seen, 2 thread 0 done
// streaming, 3 whole CTA done streaming, 4 CTA partial written, 5 (last CTA only) packet published
static __device__ __forceinline__ void nls_trace(const NlsPassParams &prm, int slot)
{
    if (prm.trace && threadIdx.x == 0 && blockIdx.y == 0)
        prm.trace[(size_t)blockIdx.x * 32 + slot] = nls_globaltimer();
}
// slots 8 + w: warp w is done streaming (stamped by its lane 0)
static __device__ __forceinline__ void nls_trace_warp(const NlsPassParams &prm)
{
    if (prm.trace && (threadIdx.x & 31) == 0 && blockIdx.y == 0 && (threadIdx.x >> 5) < 24)
        prm.trace[(size_t)blockIdx.x * 32 + 8 + (threadIdx.x >> 5)] = nls_globaltimer();
}
#define NLS_WATCHDOG_NS 60000000000ull /* default when the host passes no period (prm.watchdog_ns == 0) */

// ------------------------------------------------------------------------------------ model glue
struct NlsThread {
    double th[NLS_P];  // parameters
    double vv[NLS_P];  // velocity / direction (modes 2, 3)
    double dl[NLS_P];  // finite-difference steps  (src/fdjac.c:39-41, :96-98)
    double idl[NLS_P]; // 1 / step
    double h_fvv;
};

static __device__ __forceinline__ void nls_fj(const NlsThread &T, const double *x, double &f, double *J)
{
#if GSLNLS_JAC_MODE == 0
    nls_model_fj(T.th, x, f, J);
#elif GSLNLS_JAC_MODE == 1
    // forward differences: (f(theta + delta e_j) - f(theta)) * (1/delta), src/fdjac.c:44-59
    f = nls_model_f(T.th, x);
#pragma unroll
    for (int j = 0; j < NLS_P; ++j) {
        double tp[NLS_P];
#pragma unroll
        for (int k = 0; k < NLS_P; ++k)
            tp[k] = T.th[k];
        tp[j] = T.th[j] + T.dl[j];
        J[j] = (nls_model_f(tp, x) - f) * T.idl[j];
    }
#else
    // centred differences at +-delta/2, src/fdjac.c:101-124
    f = nls_model_f(T.th, x);
#pragma unroll
    for (int j = 0; j < NLS_P; ++j) {
        double tp[NLS_P];
#pragma unroll
        for (int k = 0; k < NLS_P; ++k)
            tp[k] = T.th[k];
        tp[j] = T.th[j] + 0.5 * T.dl[j];
        const double f1 = nls_model_f(tp, x);
        tp[j] = T.th[j] - 0.5 * T.dl[j];
        const double f0 = nls_model_f(tp, x);
        J[j] = (f1 - f0) * T.idl[j];
    }
#endif
}

// one observation, accumulated into the thread-private packet
template <int MODE>
static __device__ __forceinline__ void nls_observe(const NlsThread &T, const double *x, double y, double w,
                                                   double *acc, int &nbad)
{
    double f, J[NLS_P];
    nls_fj(T, x, f, J);
#if NLS_HAS_W
    const double sw = sqrt(w); // sqrt_wts_i = sqrt(w_i), src/fdf.c:60-64
#else
    (void)w;
#endif
    if (MODE == NLS_MODE_FJ) {
        // non-finite model value -> residual +Inf (src/nls_large.c:464-465), and counted: gslcblas
        // dnrm2 turns a vector with two or more Inf entries into NaN, which changes the reference's
        // accept/reject decision (see trs_core.h, norm_of).  Selects and an integer counter, no
        // branch: the observations of one loop trip stay in one basic block.
        const bool bad = !nls_finite(f);
        double r = bad ? NLS_INF : f - y;
        nbad += bad ? 1 : 0;
#if NLS_HAS_W
        r *= sw;
#endif
#if NLS_W_ROWS
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            J[j] *= sw;
#endif
        int e = 0;
#pragma unroll
        for (int i = 0; i < NLS_P; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j, ++e)
                acc[e] = fma(J[i], J[j], acc[e]);
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            acc[NLS_NPK + j] = fma(J[j], r, acc[NLS_NPK + j]);
        acc[NLS_NPK + NLS_P] = fma(r, r, acc[NLS_NPK + NLS_P]);
    } else if (MODE == NLS_MODE_FVV) {
        double h;
#if GSLNLS_FVV_MODE == 1
        h = nls_model_fvv(T.th, T.vv, x);
#elif GSLNLS_FVV_MODE == 2
        {
            // fvv = (2/h) ((f(x + h v) - f(x)) / h - J v), src/fdfvv.c:47-74
            double tp[NLS_P];
#pragma unroll
            for (int k = 0; k < NLS_P; ++k)
                tp[k] = T.th[k] + T.h_fvv * T.vv[k];
            const double fp = nls_model_f(tp, x);
            const double hinv = 1.0 / T.h_fvv;
            double u = 0.0;
#pragma unroll
            for (int k = 0; k < NLS_P; ++k)
                u += J[k] * T.vv[k];
            h = (2.0 * hinv) * ((fp - f) * hinv - u);
        }
#else
        h = 0.0;
#endif
#if NLS_HAS_W
        h *= sw;
#endif
#if NLS_W_ROWS
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            J[j] *= sw;
#endif
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            acc[j] = fma(J[j], h, acc[j]);
        acc[NLS_P] = fma(h, h, acc[NLS_P]);
    } else { // NLS_MODE_JVP: u = J d ; J^T u ; u^T u   (matrix-free products for Steihaug CG)
#if NLS_W_ROWS
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            J[j] *= sw;
#endif
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < NLS_P; ++k)
            u += J[k] * T.vv[k];
#pragma unroll
        for (int j = 0; j < NLS_P; ++j)
            acc[j] = fma(J[j], u, acc[j]);
        acc[NLS_P] = fma(u, u, acc[NLS_P]);
    }
}

// stream this CTA's share of the observations [lo, n) through nls_observe<MODE>; lo is even
template <int MODE>
static __device__ __forceinline__ void nls_stream(const NlsPassParams &prm, const NlsThread &T, double *acc,
                                                  long long lo = 0, int nbad = 0)
{
    const long long n = prm.n;
    const long long stride = (long long)gridDim.x * NLS_BLOCK;
    long long i = (long long)blockIdx.x * NLS_BLOCK + threadIdx.x;
#if NLS_VEC == 2
    const long long nv = n >> 1;
    i += lo >> 1;
#if NLS_PREFETCH
    // software pipeline: the loads of trip k+1 are issued before the arithmetic of trip k, so every
    // warp has NLS_UNROLL 16-byte loads per column in flight while it computes (without this a warp's
    // loads and its arithmetic alternate and only the other warps of the scheduler cover the latency)
    if (i + (NLS_UNROLL - 1) * stride < nv) {
        double2 xv[NLS_UNROLL][NLS_NV], yv[NLS_UNROLL], wv[NLS_UNROLL];
        double2 xn[NLS_UNROLL][NLS_NV], yn[NLS_UNROLL], wn[NLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            const long long o = 2 * (i + u * stride);
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k)
                xv[u][k] = NLS_LD2(prm.vars[k], o);
            yv[u] = NLS_LD2(prm.y, o);
#if NLS_HAS_W
            wv[u] = NLS_LD2(prm.w, o);
#else
            wv[u] = make_double2(1.0, 1.0);
#endif
        }
        for (;;) {
            const long long j = i + NLS_UNROLL * stride;
            const bool more = j + (NLS_UNROLL - 1) * stride < nv;
            if (more) {
#pragma unroll
                for (int u = 0; u < NLS_UNROLL; ++u) {
                    const long long o = 2 * (j + u * stride);
#pragma unroll
                    for (int k = 0; k < GSLNLS_NVAR; ++k)
                        xn[u][k] = NLS_LD2(prm.vars[k], o);
                    yn[u] = NLS_LD2(prm.y, o);
#if NLS_HAS_W
                    wn[u] = NLS_LD2(prm.w, o);
#else
                    wn[u] = make_double2(1.0, 1.0);
#endif
                }
            }
#pragma unroll
            for (int u = 0; u < NLS_UNROLL; ++u) {
                double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k) {
                    xa[k] = xv[u][k].x;
                    xb[k] = xv[u][k].y;
                }
                nls_observe<MODE>(T, xa, yv[u].x, wv[u].x, acc, nbad);
                nls_observe<MODE>(T, xb, yv[u].y, wv[u].y, acc, nbad);
            }
            i = j;
            if (!more)
                break;
#pragma unroll
            for (int u = 0; u < NLS_UNROLL; ++u) {
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k)
                    xv[u][k] = xn[u][k];
                yv[u] = yn[u];
                wv[u] = wn[u];
            }
        }
    }
#else
    for (; i + (NLS_UNROLL - 1) * stride < nv; i += NLS_UNROLL * stride) {
        double2 xv[NLS_UNROLL][NLS_NV], yv[NLS_UNROLL], wv[NLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            const long long o = 2 * (i + u * stride);
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k)
                xv[u][k] = NLS_LD2(prm.vars[k], o);
            yv[u] = NLS_LD2(prm.y, o);
#if NLS_HAS_W
            wv[u] = NLS_LD2(prm.w, o);
#else
            wv[u] = make_double2(1.0, 1.0);
#endif
        }
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k) {
                xa[k] = xv[u][k].x;
                xb[k] = xv[u][k].y;
            }
            nls_observe<MODE>(T, xa, yv[u].x, wv[u].x, acc, nbad);
            nls_observe<MODE>(T, xb, yv[u].y, wv[u].y, acc, nbad);
        }
    }
#endif // NLS_PREFETCH
#if NLS_UNROLL > 1
    // what is left is less than one full trip: at most NLS_UNROLL - 1 strided slots per thread.  All their
    // loads go out together (a one-slot-at-a-time loop would pay one memory round trip per slot, which is
    // what a 200 MB shard of an 8-GPU run spent 10 % of its pass on)
    if (i < nv) {
        double2 xt[NLS_UNROLL - 1][NLS_NV], yt[NLS_UNROLL - 1], wt[NLS_UNROLL - 1];
#pragma unroll
        for (int u = 0; u < NLS_UNROLL - 1; ++u) {
            const long long o = 2 * (i + u * stride);
            if (o < 2 * nv) {
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k)
                    xt[u][k] = NLS_LD2(prm.vars[k], o);
                yt[u] = NLS_LD2(prm.y, o);
#if NLS_HAS_W
                wt[u] = NLS_LD2(prm.w, o);
#endif
            }
        }
#pragma unroll
        for (int u = 0; u < NLS_UNROLL - 1; ++u) {
            if (i + u * stride < nv) {
                double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k) {
                    xa[k] = xt[u][k].x;
                    xb[k] = xt[u][k].y;
                }
#if NLS_HAS_W
                const double2 ww = wt[u];
#else
                const double2 ww = make_double2(1.0, 1.0);
#endif
                nls_observe<MODE>(T, xa, yt[u].x, ww.x, acc, nbad);
                nls_observe<MODE>(T, xb, yt[u].y, ww.y, acc, nbad);
            }
        }
    }
#else
    for (; i < nv; i += stride) {
        const long long o = 2 * i;
        double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k) {
            const double2 t = NLS_LD2(prm.vars[k], o);
            xa[k] = t.x;
            xb[k] = t.y;
        }
        const double2 yy = NLS_LD2(prm.y, o);
#if NLS_HAS_W
        const double2 ww = NLS_LD2(prm.w, o);
#else
        const double2 ww = make_double2(1.0, 1.0);
#endif
        nls_observe<MODE>(T, xa, yy.x, ww.x, acc, nbad);
        nls_observe<MODE>(T, xb, yy.y, ww.y, acc, nbad);
    }
#endif // NLS_UNROLL > 1
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long long o = n - 1;
        double xa[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = nls_ld1(prm.vars[k] + o);
#if NLS_HAS_W
        const double ww = nls_ld1(prm.w + o);
#else
        const double ww = 1.0;
#endif
        nls_observe<MODE>(T, xa, nls_ld1(prm.y + o), ww, acc, nbad);
    }
#else
    i += lo;
    for (; i + (NLS_UNROLL - 1) * stride < n; i += NLS_UNROLL * stride) {
        double xv[NLS_UNROLL][NLS_NV], yv[NLS_UNROLL], wv[NLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            const long long o = i + u * stride;
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k)
                xv[u][k] = nls_ld1(prm.vars[k] + o);
            yv[u] = nls_ld1(prm.y + o);
#if NLS_HAS_W
            wv[u] = nls_ld1(prm.w + o);
#else
            wv[u] = 1.0;
#endif
        }
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u)
            nls_observe<MODE>(T, xv[u], yv[u], wv[u], acc, nbad);
    }
    for (; i < n; i += stride) {
        double xa[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = nls_ld1(prm.vars[k] + i);
#if NLS_HAS_W
        const double ww = nls_ld1(prm.w + i);
#else
        const double ww = 1.0;
#endif
        nls_observe<MODE>(T, xa, nls_ld1(prm.y + i), ww, acc, nbad);
    }
#endif
    if (MODE == NLS_MODE_FJ)
        acc[NLS_NPK + NLS_P + 1] = (double)nbad;
}

#if NLS_TILED == 2
// ------------------------------------------------------------------------------------ K1, TMA-staged stream
// The columns arrive in shared memory through the bulk-copy engine (cp.async.bulk, the 1-D TMA form)
// instead of through per-thread LDG: one producer lane keeps NLS_STAGES tiles of NLS_TILE
// observations per column in flight per CTA, completion is signalled on a "full" mbarrier by the
// transaction count, the 8 consumer warps read their observations with conflict-free LDS.128 and
// release the slot on an "empty" mbarrier as soon as the values sit in registers.  Bytes in flight
// per SM no longer depend on registers per thread (NLS_STAGES x tile bytes per CTA), and the loop
// body is pure FP64 work.  Tiles go to CTAs round-robin; the ragged end of the shard (< one tile)
// runs through the register path above.  Needs 16-byte aligned columns (NLS_VEC == 2).
#ifndef NLS_STAGES
#define NLS_STAGES 4
#endif
#define NLS_NCONS (NLS_BLOCK - 32)             /* consumer threads; the last warp is the producer */
#define NLS_NCW (NLS_NCONS / 32)
#define NLS_TILE (NLS_NCONS * 2 * NLS_UNROLL)  /* observations per stage                          */
#define NLS_NARR (GSLNLS_NVAR + 1 + NLS_HAS_W)
#define NLS_STAGE_DOUBLES (NLS_NARR * NLS_TILE)

static __device__ __forceinline__ unsigned nls_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void nls_bar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nls_saddr(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void nls_bar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nls_saddr(bar)) : "memory");
}
static __device__ __forceinline__ void nls_bar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nls_saddr(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void nls_bar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "NLS_WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra NLS_DONE_%=;\n"
                 "bra NLS_WAIT_%=;\n"
                 "NLS_DONE_%=:\n"
                 "}" ::"r"(nls_saddr(bar)), "r"(parity) : "memory");
}
// NLS_TMA_HINT 1: the bulk copies carry an L2 evict_first policy.  The pass itself runs equally fast with or without (240.0 vs 240.4 us at n = 1e8), but without
// the hint 1.6 GB of streamed columns push the trust-region server's state, packet and request lines out
// of L2 and every step then starts with DRAM round trips: measured step latency 9.8 us (lm) / 18.7 us
// (dogleg) without the hint, 6.3 / 9.0 us with it.
#ifndef NLS_TMA_HINT
#define NLS_TMA_HINT 1
#endif
static __device__ __forceinline__ void nls_bulk_g2s(double *dst, const double *src, unsigned bytes,
                                                    unsigned long long *bar, unsigned long long policy)
{
#if NLS_TMA_HINT
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(nls_saddr(dst)), "l"(src), "r"(bytes), "r"(nls_saddr(bar)), "l"(policy) : "memory");
#else
    (void)policy;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(nls_saddr(dst)), "l"(src), "r"(bytes), "r"(nls_saddr(bar)) : "memory");
#endif
}

template <int MODE>
static __device__ __forceinline__ void nls_stream_tma(const NlsPassParams &prm, const NlsThread &T, double *acc)
{
    extern __shared__ __align__(128) unsigned char nls_dyn[];
    double *buf = reinterpret_cast<double *>(nls_dyn); // [NLS_STAGES][NLS_NARR][NLS_TILE]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(buf + (size_t)NLS_STAGES * NLS_STAGE_DOUBLES);
    unsigned long long *empty = full + NLS_STAGES;
    const long long ntile = prm.n / NLS_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NLS_STAGES; ++s) {
            nls_bar_init(full + s, 1);
            nls_bar_init(empty + s, NLS_NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int nbad = 0;
    if (warp == NLS_NCW) {
        if (lane == 0) {
            const NlsPolicy PL = nls_policy(prm.keep_rows);
            int s = 0;
            unsigned ph = 1u; // a fresh barrier lets a wait on the "previous" phase through
            for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
                nls_bar_wait(empty + s, ph);
                nls_bar_expect_tx(full + s, (unsigned)(NLS_STAGE_DOUBLES * sizeof(double)));
                double *dst = buf + (size_t)s * NLS_STAGE_DOUBLES;
                const long long o = t * NLS_TILE;
                const unsigned long long policy = (o < PL.keep_rows) ? PL.keep : PL.stream;
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k)
                    nls_bulk_g2s(dst + k * NLS_TILE, prm.vars[k] + o, NLS_TILE * 8u, full + s, policy);
                nls_bulk_g2s(dst + GSLNLS_NVAR * NLS_TILE, prm.y + o, NLS_TILE * 8u, full + s, policy);
#if NLS_HAS_W
                nls_bulk_g2s(dst + (GSLNLS_NVAR + 1) * NLS_TILE, prm.w + o, NLS_TILE * 8u, full + s, policy);
#endif
                if (++s == NLS_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        int s = 0;
        unsigned ph = 0u;
        for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
            nls_bar_wait(full + s, ph);
            const double *b = buf + (size_t)s * NLS_STAGE_DOUBLES;
            double2 xv[NLS_UNROLL][NLS_NV], yv[NLS_UNROLL], wv[NLS_UNROLL];
#pragma unroll
            for (int u = 0; u < NLS_UNROLL; ++u) {
                const int o = 2 * (u * NLS_NCONS + (int)threadIdx.x);
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k)
                    xv[u][k] = *reinterpret_cast<const double2 *>(b + k * NLS_TILE + o);
                yv[u] = *reinterpret_cast<const double2 *>(b + GSLNLS_NVAR * NLS_TILE + o);
#if NLS_HAS_W
                wv[u] = *reinterpret_cast<const double2 *>(b + (GSLNLS_NVAR + 1) * NLS_TILE + o);
#else
                wv[u] = make_double2(1.0, 1.0);
#endif
            }
            __syncwarp();
            if (lane == 0)
                nls_bar_arrive(empty + s); // the slot can be refilled while we compute
#pragma unroll
            for (int u = 0; u < NLS_UNROLL; ++u) {
                double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
                for (int k = 0; k < GSLNLS_NVAR; ++k) {
                    xa[k] = xv[u][k].x;
                    xb[k] = xv[u][k].y;
                }
                nls_observe<MODE>(T, xa, yv[u].x, wv[u].x, acc, nbad);
                nls_observe<MODE>(T, xb, yv[u].y, wv[u].y, acc, nbad);
            }
            if (++s == NLS_STAGES) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
    // rows past the last full tile: register path, all warps
    nls_stream<MODE>(prm, T, acc, ntile * NLS_TILE, nbad);
}
#endif // NLS_TILED == 2

// ------------------------------------------------------------------------------------ K1 head / tail
// Shared by the register-accumulator kernel below and the tiled DMMA kernel (nls_pass_tiled.cuh).

// Start of a launch: in server mode wait for this pass's request; returns the mode to run
// (NLS_MODE_IDLE: leave) and the pass sequence number.  Uniform across the CTA.
// the pass sequence number lives in shared memory, not in a register across the streaming loop
static __device__ __forceinline__ unsigned long long *nls_seq_slot()
{
    __shared__ unsigned long long s_seq;
    return &s_seq;
}

static __device__ __forceinline__ int nls_begin(const NlsPassParams &prm, const double *req)
{
    unsigned long long &s_seq = *nls_seq_slot();
    if (prm.channel) {
        // server mode: this launch is pass number k = (passes completed so far) + 1; its request is
        // published by the resident trust-region warp (trs_server) as soon as it has digested
        // pass k-1, typically while this grid is still being dispatched.
        if (threadIdx.x == 0) {
            const unsigned long long k = __ldcg((const unsigned long long *)(prm.channel + NLS_CH_PASS_CTR)) + 1ull;
            const unsigned long long *rs = (const unsigned long long *)(prm.channel + NLS_CH_REQ_SEQ);
            // start-of-fit handshake: tell the server that pass k is executing (i.e. that the two kernels
            // run concurrently); a server that never sees this gives up and the host falls back to
            // launch-ordered stepping instead of deadlocking under a serialising tool
            if (blockIdx.x == 0 && blockIdx.y == 0)
                nls_st_release_gpu((unsigned long long *)(prm.channel + NLS_CH_PASS_SEEN), k);
            const unsigned long long wd = prm.watchdog_ns ? prm.watchdog_ns : NLS_WATCHDOG_NS;
            unsigned long long t0 = 0ull, spins = 0ull, kk = k;
            while (nls_ld_acquire_gpu(rs) < k) {
                if ((++spins & 1023ull) == 0ull) {
                    const unsigned long long t = nls_globaltimer();
                    if (t0 == 0ull)
                        t0 = t;
                    else if (t - t0 > wd) {
                        kk = 0ull; // give up: behave like an idle launch
                        break;
                    }
                }
            }
            s_seq = kk;
            if (blockIdx.x == 0 && blockIdx.y == 0)
                *(unsigned long long *)(prm.channel + NLS_CH_TIMER) = nls_globaltimer();
        }
        __syncthreads();
        if (s_seq == 0ull)
            return NLS_MODE_IDLE;
    }
    const int mode = prm.force_mode > 0 ? prm.force_mode : (int)__ldcg(req);
    if (mode != NLS_MODE_IDLE && prm.prof_flag && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        *prm.prof_flag = 1;
    return mode;
}

static __device__ __forceinline__ void nls_load_request(const NlsPassParams &prm, const double *req, NlsThread &T)
{
#pragma unroll
    for (int j = 0; j < NLS_P; ++j) {
        T.th[j] = __ldcg(req + 1 + j);
        T.vv[j] = __ldcg(req + 1 + NLS_P + j);
        double d = prm.h_df * fabs(T.th[j]);
        if (d == 0.0)
            d = prm.h_df;
        T.dl[j] = d;
        T.idl[j] = 1.0 / d;
    }
    T.h_fvv = prm.h_fvv;
}

// hand the reduced packet on: prm.packet (launch-ordered mode) or slot [seq parity][this rank] of
// every rank's mailbox (server mode; peer memory over NVLink), then publish
static __device__ __forceinline__ void nls_packet_out(const NlsPassParams &prm, int cand, unsigned long long seq,
                                                      int e, double s)
{
    if (prm.channel) {
        const size_t slot = ((size_t)(seq & 1ull) * NLS_MAX_RANKS + (size_t)prm.rank) * NLS_CH_MAXPK;
        for (int q = 0; q < prm.nranks; ++q)
            ((double *)(prm.peer_channel[q] + NLS_CH_DATA))[slot + e] = s;
    } else {
        prm.packet[(size_t)cand * prm.pk_stride + e] = s;
    }
}
static __device__ __forceinline__ void nls_packet_publish(const NlsPassParams &prm, int cand, unsigned long long seq)
{
    nls_trace(prm, 5);
    if (prm.channel) {
        if (prm.nranks > 1)
            __threadfence_system();
        else
            __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            prm.ticket[cand] = 0u;
            unsigned long long *tm = (unsigned long long *)(prm.channel + NLS_CH_TIMER);
            // time from "request seen" (CTA 0) to "packet out", summed over passes
            tm[1] = __ldcg(tm + 1) + (nls_globaltimer() - __ldcg(tm));
            tm[2] = __ldcg(tm + 2) + 1ull;
            *(unsigned long long *)(prm.channel + NLS_CH_PASS_CTR) = seq;
        }
        if (prm.nranks > 1) {
            // one lane per peer: the release stores (each is a system-scope fence + store, i.e. one NVLink
            // round trip) go out side by side.  Issued one after the other by a single thread they cost
            // ~2 us per rank: 17 us of every 79 us step on 8 GPUs.
            if ((int)threadIdx.x < prm.nranks)
                nls_st_release_sys((unsigned long long *)(prm.peer_channel[threadIdx.x] + NLS_CH_FLAGS + 128 * prm.rank), seq);
        } else if (threadIdx.x == 0) {
            nls_st_release_gpu((unsigned long long *)(prm.channel + NLS_CH_FLAGS + 128 * prm.rank), seq);
        }
        return;
    }
    if (threadIdx.x == 0)
        prm.ticket[cand] = 0u; // ready for the next launch
}

// End of a launch, after this CTA's partial packet (NLS_PK doubles) has been written to
// prm.partials.  No floating-point atomics anywhere: partials are summed in CTA order, by whichever
// CTA happens to arrive last, so the packet is bit-identical from run to run.
//   flat (multi-candidate launches): the last CTA of a candidate sums all its CTA partials
//   two-level (single candidate):    the last CTA of each group of NLS_RED_GROUP sums the group,
//                                    the last group to finish sums the group sums; every thread
//                                    owns packet entries, so a 1226-entry packet (p = 48) over
//                                    444 CTAs costs two short, fully parallel rounds
static __device__ __forceinline__ void nls_grid_finish(const NlsPassParams &prm, int cand)
{
    __shared__ int s_last;
    const unsigned long long seq = prm.channel ? *nls_seq_slot() : 0ull;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __threadfence();
    __syncthreads();
    if (prm.group_partials) {
        const int grp = blockIdx.x / NLS_RED_GROUP, ngrp = (gridDim.x + NLS_RED_GROUP - 1) / NLS_RED_GROUP;
        const int b0 = grp * NLS_RED_GROUP;
        const int gsize = min(NLS_RED_GROUP, (int)gridDim.x - b0);
        if (threadIdx.x == 0)
            s_last = (atomicAdd(prm.group_ticket + grp, 1u) == (unsigned)(gsize - 1));
        __syncthreads();
        if (!s_last)
            return;
        __threadfence();
        const double *parts = prm.partials + (size_t)b0 * prm.pk_stride;
        double *gp = prm.group_partials + (size_t)grp * prm.pk_stride;
        for (int e = threadIdx.x; e < NLS_PK; e += NLS_BLOCK) {
            double s = 0.0;
#pragma unroll 8
            for (int b = 0; b < gsize; ++b)
                s += __ldcg(parts + (size_t)b * prm.pk_stride + e);
            gp[e] = s;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            prm.group_ticket[grp] = 0u;
            s_last = (atomicAdd(prm.ticket + cand, 1u) == (unsigned)(ngrp - 1));
        }
        __syncthreads();
        if (!s_last)
            return;
        __threadfence();
        for (int e = threadIdx.x; e < NLS_PK; e += NLS_BLOCK) {
            double s = 0.0;
#pragma unroll 4
            for (int g = 0; g < ngrp; ++g)
                s += __ldcg(prm.group_partials + (size_t)g * prm.pk_stride + e);
            nls_packet_out(prm, cand, seq, e, s);
        }
        nls_packet_publish(prm, cand, seq);
        return;
    }
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(prm.ticket + cand, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    const double *parts = prm.partials + (size_t)cand * gridDim.x * prm.pk_stride;
    for (int e = warp; e < NLS_PK; e += NLS_NW) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32)
            s += __ldcg(parts + (size_t)b * prm.pk_stride + e);
        s += __shfl_down_sync(0xffffffffu, s, 16);
        s += __shfl_down_sync(0xffffffffu, s, 8);
        s += __shfl_down_sync(0xffffffffu, s, 4);
        s += __shfl_down_sync(0xffffffffu, s, 2);
        s += __shfl_down_sync(0xffffffffu, s, 1);
        if (lane == 0)
            nls_packet_out(prm, cand, seq, e, s);
    }
    nls_packet_publish(prm, cand, seq);
}

#if NLS_TILED == 1
#include "nls_pass_tiled.cuh"
#else
// ------------------------------------------------------------------------------------ K1 (register accumulators)
extern "C" __global__ void __launch_bounds__(NLS_BLOCK, NLS_MINB) nls_pass(const NlsPassParams prm)
{
    const int cand = blockIdx.y;
    const double *req = prm.req + (size_t)cand * prm.req_stride;
    nls_trace(prm, 0);
    const int mode = nls_begin(prm, req);
    if (mode == NLS_MODE_IDLE)
        return; // this candidate has finished; uniform for the whole CTA
    nls_trace(prm, 1);
    nls_exp_init();

    NlsThread T;
    nls_load_request(prm, req, T);

    double acc[NLS_PK];
#pragma unroll
    for (int e = 0; e < NLS_PK; ++e)
        acc[e] = 0.0;

#if NLS_TILED == 2
    if (mode == NLS_MODE_FJ)
        nls_stream_tma<NLS_MODE_FJ>(prm, T, acc);
    else if (mode == NLS_MODE_FVV)
        nls_stream_tma<NLS_MODE_FVV>(prm, T, acc);
    else
        nls_stream_tma<NLS_MODE_JVP>(prm, T, acc);
#else
    if (mode == NLS_MODE_FJ)
        nls_stream<NLS_MODE_FJ>(prm, T, acc);
    else if (mode == NLS_MODE_FVV)
        nls_stream<NLS_MODE_FVV>(prm, T, acc);
    else
        nls_stream<NLS_MODE_JVP>(prm, T, acc);
#endif

    // ---- CTA reduction: fixed shuffle tree, then warps summed in warp order ----
    __shared__ double sred[NLS_NW][NLS_PK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    nls_trace(prm, 2);
    nls_trace_warp(prm);
#pragma unroll
    for (int e = 0; e < NLS_PK; ++e) {
        double v = acc[e];
        v += __shfl_down_sync(0xffffffffu, v, 16);
        v += __shfl_down_sync(0xffffffffu, v, 8);
        v += __shfl_down_sync(0xffffffffu, v, 4);
        v += __shfl_down_sync(0xffffffffu, v, 2);
        v += __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 0)
            sred[warp][e] = v;
    }
    __syncthreads();
    nls_trace(prm, 3);
    double *part = prm.partials + ((size_t)cand * gridDim.x + blockIdx.x) * prm.pk_stride;
    for (int e = threadIdx.x; e < NLS_PK; e += NLS_BLOCK) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NLS_NW; ++w)
            s += sred[w][e];
        part[e] = s;
    }
    nls_trace(prm, 4);
    nls_grid_finish(prm, cand);
}
#endif // NLS_TILED == 1

#if NLS_TILED == 2
// ------------------------------------------------------------------------------------ K1, persistent
// One launch per FIT instead of one per pass (resident-server mode, TMA ring).  The grid stays on the SMs and
// every CTA loops over the passes of the fit: wait for request k from the trust-region server, stream the
// shard, reduce, the last CTA to arrive deposits the packet; the server steps and publishes request k + 1.
// What this removes from every pass of a small shard (200 MB = 31 us of HBM time on an 8-GPU split):
//   * the launch and its dispatch, and the cold instruction / constant lines of the epilogue: code that ran
//     35 us ago is still in the SM's instruction cache, code of a fresh launch is fetched from DRAM behind
//     20 MB of queued streaming requests (measured: 11 us between "all warps done" and "partial written");
//   * the empty ring at the start of a pass: x and y do not depend on theta, so the producer lane does not
//     stop at the end of a pass -- it refills the ring with the FIRST tiles of the next pass while the
//     consumers reduce and the server steps (4 stages x 36 KB x 148 SMs = 21 MB are in shared memory when
//     request k + 1 lands).
// Consumers synchronise among themselves on named barrier 1 (the producer warp never joins a CTA barrier
// after the set-up).  Same arithmetic per observation as nls_pass; the thread -> row mapping of the ragged
// tail differs, so packets agree with the per-launch kernels to rounding, not bitwise.
static __device__ __forceinline__ void nls_cons_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NLS_NCONS) : "memory");
}
static __device__ __forceinline__ void nls_l2_prefetch(const double *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
static __device__ __forceinline__ bool nls_bar_try(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}" : "=r"(ok) : "r"(nls_saddr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}

template <int MODE>
static __device__ __forceinline__ void nls_persistent_stream(const NlsPassParams &prm, const NlsThread &T, double *acc,
                                                             int &nbad, const double *buf, unsigned long long *full,
                                                             unsigned long long *empty, long long my_tiles,
                                                             unsigned long long &gc)
{
    const int lane = threadIdx.x & 31;
    for (long long i = 0; i < my_tiles; ++i, ++gc) {
        const int s = (int)(gc % NLS_STAGES);
        nls_bar_wait(full + s, (unsigned)((gc / NLS_STAGES) & 1ull));
        const double *b = buf + (size_t)s * NLS_STAGE_DOUBLES;
        double2 xv[NLS_UNROLL][NLS_NV], yv[NLS_UNROLL], wv[NLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            const int o = 2 * (u * NLS_NCONS + (int)threadIdx.x);
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k)
                xv[u][k] = *reinterpret_cast<const double2 *>(b + k * NLS_TILE + o);
            yv[u] = *reinterpret_cast<const double2 *>(b + GSLNLS_NVAR * NLS_TILE + o);
#if NLS_HAS_W
            wv[u] = *reinterpret_cast<const double2 *>(b + (GSLNLS_NVAR + 1) * NLS_TILE + o);
#else
            wv[u] = make_double2(1.0, 1.0);
#endif
        }
        __syncwarp();
        if (lane == 0)
            nls_bar_arrive(empty + s); // the slot can be refilled while we compute
#pragma unroll
        for (int u = 0; u < NLS_UNROLL; ++u) {
            double xa[NLS_NV], xb[NLS_NV];
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k) {
                xa[k] = xv[u][k].x;
                xb[k] = xv[u][k].y;
            }
            nls_observe<MODE>(T, xa, yv[u].x, wv[u].x, acc, nbad);
            nls_observe<MODE>(T, xb, yv[u].y, wv[u].y, acc, nbad);
        }
    }
    // rows past the last full tile: one scalar load per consumer thread and grid stride
    for (long long r = (prm.n / NLS_TILE) * NLS_TILE + (long long)blockIdx.x * NLS_NCONS + threadIdx.x; r < prm.n;
         r += (long long)gridDim.x * NLS_NCONS) {
        double xa[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = nls_ld1(prm.vars[k] + r);
#if NLS_HAS_W
        const double ww = nls_ld1(prm.w + r);
#else
        const double ww = 1.0;
#endif
        nls_observe<MODE>(T, xa, nls_ld1(prm.y + r), ww, acc, nbad);
    }
    if (MODE == NLS_MODE_FJ)
        acc[NLS_NPK + NLS_P + 1] = (double)nbad;
}

extern "C" __global__ void __launch_bounds__(NLS_BLOCK, 1) nls_pass_persistent(const NlsPassParams prm)
{
    extern __shared__ __align__(128) unsigned char nls_dyn[];
    double *buf = reinterpret_cast<double *>(nls_dyn); // [NLS_STAGES][NLS_NARR][NLS_TILE]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(buf + (size_t)NLS_STAGES * NLS_STAGE_DOUBLES);
    unsigned long long *empty = full + NLS_STAGES;
    __shared__ double sred[NLS_NCW][NLS_PK];
    __shared__ unsigned long long s_seq, s_consumed;
    __shared__ int s_mode, s_last;
    __shared__ double s_req[2 * NLS_P]; // theta, v of the current request (fetched once per CTA)
    __shared__ volatile int s_stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntile = prm.n / NLS_TILE;
    const long long my_tiles = ntile > (long long)blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NLS_STAGES; ++s) {
            nls_bar_init(full + s, 1);
            nls_bar_init(empty + s, NLS_NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_stop = 0;
        s_consumed = 0ull;
    }
    nls_exp_init(); // contains a CTA-wide barrier: the last one the producer warp takes part in
    __syncthreads();

    if (warp == NLS_NCW) {
        // ---- producer lane: keeps the ring full, across pass boundaries ----
        if (lane == 0 && my_tiles > 0) {
            const NlsPolicy PL = nls_policy(prm.keep_rows);
            unsigned long long g = 0ull; // tiles issued since the kernel started
            // L2 run-ahead: beyond the ring, the next prm.l2_ahead tiles of this CTA are requested into L2.
            // Inside a pass that only moves HBM reads a few microseconds earlier; between passes -- while
            // the consumers reduce, the packet travels and the server steps, 9 us with the ring already full
            // -- it keeps HBM busy: 8 tiles x 36 KB x 147 SMs = 43 MB of the next pass are L2 hits.
            unsigned long long pf = NLS_STAGES; // tiles covered so far (the ring itself needs no prefetch)
            long long pfi = NLS_STAGES % my_tiles;
            bool stop = false;
            while (!stop) {
                for (long long i = 0; i < my_tiles && !stop; ++i) {
                    while (pf < g + NLS_STAGES + (unsigned long long)prm.l2_ahead) {
                        const long long po = ((long long)blockIdx.x + pfi * gridDim.x) * NLS_TILE;
#pragma unroll
                        for (int k = 0; k < GSLNLS_NVAR; ++k)
                            nls_l2_prefetch(prm.vars[k] + po, NLS_TILE * 8u);
                        nls_l2_prefetch(prm.y + po, NLS_TILE * 8u);
#if NLS_HAS_W
                        nls_l2_prefetch(prm.w + po, NLS_TILE * 8u);
#endif
                        ++pf;
                        if (++pfi == my_tiles)
                            pfi = 0;
                    }
                    const int s = (int)(g % NLS_STAGES);
                    const unsigned ph = (unsigned)(((g / NLS_STAGES) & 1ull) ^ 1ull);
                    while (!nls_bar_try(empty + s, ph)) {
                        if (s_stop) {
                            stop = true;
                            break;
                        }
                    }
                    if (stop)
                        break;
                    nls_bar_expect_tx(full + s, (unsigned)(NLS_STAGE_DOUBLES * sizeof(double)));
                    double *dst = buf + (size_t)s * NLS_STAGE_DOUBLES;
                    const long long o = ((long long)blockIdx.x + i * gridDim.x) * NLS_TILE;
                    const unsigned long long policy = (o < PL.keep_rows) ? PL.keep : PL.stream;
#pragma unroll
                    for (int k = 0; k < GSLNLS_NVAR; ++k)
                        nls_bulk_g2s(dst + k * NLS_TILE, prm.vars[k] + o, NLS_TILE * 8u, full + s, policy);
                    nls_bulk_g2s(dst + GSLNLS_NVAR * NLS_TILE, prm.y + o, NLS_TILE * 8u, full + s, policy);
#if NLS_HAS_W
                    nls_bulk_g2s(dst + (GSLNLS_NVAR + 1) * NLS_TILE, prm.w + o, NLS_TILE * 8u, full + s, policy);
#endif
                    ++g;
                }
            }
            // tiles requested for a pass that will not happen: their bytes must land before the CTA may go
            __threadfence_block();
            for (unsigned long long j = s_consumed; j < g; ++j)
                nls_bar_wait(full + (int)(j % NLS_STAGES), (unsigned)((j / NLS_STAGES) & 1ull));
        }
        return;
    }

    // ---- consumers ----
    unsigned long long gc = 0ull; // tiles consumed since the kernel started
    int done_passes = 0;
    unsigned long long k = 0ull;
    for (int pass = 0; pass < prm.max_passes; ++pass) {
        // wait for this pass's request (published by the resident trust-region server)
        if (threadIdx.x == 0) {
            if (pass == 0)
                k = __ldcg((const unsigned long long *)(prm.channel + NLS_CH_PASS_CTR)) + 1ull;
            else
                ++k;
            const unsigned long long *rs = (const unsigned long long *)(prm.channel + NLS_CH_REQ_SEQ);
            const unsigned long long *ab = (const unsigned long long *)(prm.channel + NLS_CH_ABORT);
            if (pass == 0 && blockIdx.x == 0) // start-of-fit handshake, see nls_begin
                nls_st_release_gpu((unsigned long long *)(prm.channel + NLS_CH_PASS_SEEN), k);
            const unsigned long long wd = prm.watchdog_ns ? prm.watchdog_ns : NLS_WATCHDOG_NS;
            unsigned long long t0 = 0ull, spins = 0ull;
            int mode = -1;
            while (nls_ld_acquire_gpu(rs) < k) {
                if ((++spins & 1023ull) == 0ull) {
                    const unsigned long long t = nls_globaltimer();
                    if (t0 == 0ull)
                        t0 = t;
                    else if (t - t0 > wd || nls_ld_acquire_gpu(ab) != 0ull) {
                        mode = NLS_MODE_IDLE; // give up (watchdog), or the host ended the fit early
                        break;
                    }
                }
            }
            if (mode < 0) {
                // the request record travels with the sequence number: fetch it here, once per CTA, instead of one
                // L2 round trip per thread after the barrier
                mode = (int)__ldcg(prm.req);
#pragma unroll
                for (int j = 0; j < 2 * NLS_P; ++j)
                    s_req[j] = __ldcg(prm.req + 1 + j);
            }
            s_mode = mode;
            s_seq = k;
            if (blockIdx.x == 0 && mode != NLS_MODE_IDLE)
                *(unsigned long long *)(prm.channel + NLS_CH_TIMER) = nls_globaltimer();
        }
        nls_cons_sync();
        const int mode = s_mode;
        if (mode == NLS_MODE_IDLE)
            break;
        nls_trace(prm, 1);
        NlsThread T;
#pragma unroll
        for (int j = 0; j < NLS_P; ++j) { // as nls_load_request, from the CTA's copy
            T.th[j] = s_req[j];
            T.vv[j] = s_req[NLS_P + j];
            double d = prm.h_df * fabs(T.th[j]);
            if (d == 0.0)
                d = prm.h_df;
            T.dl[j] = d;
            T.idl[j] = 1.0 / d;
        }
        T.h_fvv = prm.h_fvv;
        double acc[NLS_PK];
#pragma unroll
        for (int e = 0; e < NLS_PK; ++e)
            acc[e] = 0.0;
        int nbad = 0;
        if (mode == NLS_MODE_FJ)
            nls_persistent_stream<NLS_MODE_FJ>(prm, T, acc, nbad, buf, full, empty, my_tiles, gc);
        else if (mode == NLS_MODE_FVV)
            nls_persistent_stream<NLS_MODE_FVV>(prm, T, acc, nbad, buf, full, empty, my_tiles, gc);
        else
            nls_persistent_stream<NLS_MODE_JVP>(prm, T, acc, nbad, buf, full, empty, my_tiles, gc);
        nls_trace(prm, 2);
        nls_trace_warp(prm);

        // ---- CTA reduction among the consumer warps: fixed shuffle tree, warps summed in warp order ----
#pragma unroll
        for (int e = 0; e < NLS_PK; ++e) {
            double v = acc[e];
            v += __shfl_down_sync(0xffffffffu, v, 16);
            v += __shfl_down_sync(0xffffffffu, v, 8);
            v += __shfl_down_sync(0xffffffffu, v, 4);
            v += __shfl_down_sync(0xffffffffu, v, 2);
            v += __shfl_down_sync(0xffffffffu, v, 1);
            if (lane == 0)
                sred[warp][e] = v;
        }
        nls_cons_sync();
        nls_trace(prm, 3);
        double *part = prm.partials + (size_t)blockIdx.x * prm.pk_stride;
        for (int e = threadIdx.x; e < NLS_PK; e += NLS_NCONS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NLS_NCW; ++w)
                s += sred[w][e];
            part[e] = s;
        }
        nls_trace(prm, 4);
        if (prm.server_reduce) {
            // ---- grid: the resident server sums the CTA partials itself.  A CTA only counts itself in: no
            // ticket round trip, no last-CTA stage, no second hop for the packet (3.3 us -> ~1 us measured) ----
#if NLS_PK <= 32
            if (warp == 0) { // every partial entry was written by this warp
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(prm.channel + NLS_CH_CTA_COUNT) : "memory");
                    if (prm.prof_flag && blockIdx.x == 0)
                        *prm.prof_flag += 1;
                }
            }
#else
            __threadfence();
            nls_cons_sync();
            if (threadIdx.x == 0) {
                asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(prm.channel + NLS_CH_CTA_COUNT) : "memory");
                if (prm.prof_flag && blockIdx.x == 0)
                    *prm.prof_flag += 1;
            }
#endif
            ++done_passes;
            continue;
        }
        // ---- grid: the last CTA to arrive sums the CTA partials in CTA order and hands the packet on ----
        __threadfence();
        nls_cons_sync();
        if (threadIdx.x == 0)
            s_last = (atomicAdd(prm.ticket, 1u) == gridDim.x - 1);
        nls_cons_sync();
        if (s_last) {
            __threadfence();
            const unsigned long long seq = s_seq;
            for (int e = warp; e < NLS_PK; e += NLS_NCW) {
                double s = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32)
                    s += __ldcg(prm.partials + (size_t)b * prm.pk_stride + e);
                s += __shfl_down_sync(0xffffffffu, s, 16);
                s += __shfl_down_sync(0xffffffffu, s, 8);
                s += __shfl_down_sync(0xffffffffu, s, 4);
                s += __shfl_down_sync(0xffffffffu, s, 2);
                s += __shfl_down_sync(0xffffffffu, s, 1);
                if (lane == 0)
                    nls_packet_out(prm, 0, seq, e, s);
            }
            if (prm.nranks > 1)
                __threadfence_system();
            else
                __threadfence();
            nls_cons_sync();
            nls_trace(prm, 5);
            if (threadIdx.x == 0) {
                prm.ticket[0] = 0u;
                unsigned long long *tm = (unsigned long long *)(prm.channel + NLS_CH_TIMER);
                tm[1] = __ldcg(tm + 1) + (nls_globaltimer() - __ldcg(tm));
                tm[2] = __ldcg(tm + 2) + 1ull;
                *(unsigned long long *)(prm.channel + NLS_CH_PASS_CTR) = seq;
                if (prm.prof_flag)
                    *prm.prof_flag += 1;
            }
            if (prm.nranks > 1) {
                if ((int)threadIdx.x < prm.nranks)
                    nls_st_release_sys((unsigned long long *)(prm.peer_channel[threadIdx.x] + NLS_CH_FLAGS + 128 * prm.rank), seq);
            } else if (threadIdx.x == 0) {
                nls_st_release_gpu((unsigned long long *)(prm.channel + NLS_CH_FLAGS + 128 * prm.rank), seq);
            }
        }
        ++done_passes;
    }
    // ---- leave: tell the producer how far the ring was drained ----
    if (threadIdx.x == 0) {
        s_consumed = gc;
        __threadfence_block();
        s_stop = 1;
        if (blockIdx.x == 0 && prm.host_passes) {
            *prm.host_passes = done_passes;
            __threadfence_system();
        }
    }
}
#endif // NLS_TILED == 2 (persistent)

// ------------------------------------------------------------------------------------ K4
// resid_i = sqrt(w_i) (fn_i - y_i) and grad[i + n j] = sqrt(w_i) J_ij at the final parameters:
// the arrays C_nls_large returns at src/nls_large.c:339-385, produced once, after the fit.
extern "C" __global__ void __launch_bounds__(256) nls_materialise(const NlsMaterialiseParams prm)
{
    nls_exp_init();
    NlsThread T;
#pragma unroll
    for (int j = 0; j < NLS_P; ++j) {
        T.th[j] = prm.theta[j];
        T.vv[j] = 0.0;
        double d = prm.h_df * fabs(T.th[j]);
        if (d == 0.0)
            d = prm.h_df;
        T.dl[j] = d;
        T.idl[j] = 1.0 / d;
    }
    T.h_fvv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += stride) {
        double xa[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = prm.vars[k][i];
        double f, J[NLS_P];
        nls_fj(T, xa, f, J);
        double r = f - prm.y[i];
        if (!nls_finite(f))
            r = NLS_INF;
        const double sw = prm.w ? sqrt(prm.w[i]) : 1.0;
        if (prm.resid)
            prm.resid[i] = r * sw;
        if (prm.grad) {
#if NLS_W_GSL
            const double swj = 1.0; // the reference returns params.J as the callback left it: unweighted
#else
            const double swj = sw;
#endif
#pragma unroll
            for (int j = 0; j < NLS_P; ++j)
                prm.grad[i + prm.n * (long long)j] = J[j] * swj;
        }
    }
}

// ------------------------------------------------------------------------------------ IRLS (robust losses)
// The reference's iteratively reweighted least squares, src/nls_irls.c:412-546: after every weighted fit the
// unweighted residuals give the scale sigma = 1.4826 median |r| and the weights w_i = psi(r_i / sigma) /
// (r_i / sigma) of the chosen loss.  psi / psi' below restate src/nls_irls.c:10-330 (themselves credited to
// robustbase's lmrob.c).  The median comes from a radix select over the bit patterns of |r_i|, 8 bits per
// pass, each pass re-deriving r_i from the resident columns (16 bytes per observation, nothing stored).
static __device__ __forceinline__ double nls_irls_resid(const NlsThread &T, const NlsIrlsParams &prm, long long i)
{
    double xa[NLS_NV];
#pragma unroll
    for (int k = 0; k < GSLNLS_NVAR; ++k)
        xa[k] = prm.vars[k][i];
    const double f = nls_model_f(T.th, xa);
    return nls_finite(f) ? f - prm.y[i] : NLS_INF; // src/nls_large.c:464-467
}
static __device__ __forceinline__ void nls_irls_theta(const NlsIrlsParams &prm, NlsThread &T)
{
#pragma unroll
    for (int j = 0; j < NLS_P; ++j) {
        T.th[j] = prm.theta[j];
        T.vv[j] = 0.0;
        T.dl[j] = T.idl[j] = 0.0;
    }
    T.h_fvv = 0.0;
}

extern "C" __global__ void __launch_bounds__(256) nls_irls_hist(const NlsIrlsParams prm)
{
    __shared__ unsigned int sh[256];
    nls_exp_init();
    sh[threadIdx.x] = 0u;
    __syncthreads();
    NlsThread T;
    nls_irls_theta(prm, T);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += stride) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(nls_irls_resid(T, prm, i)));
        if (prm.shift >= 56 || (key >> (prm.shift + 8)) == prm.prefix)
            atomicAdd(&sh[(unsigned)(key >> prm.shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x])
        atomicAdd(prm.hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// #{|r| <= pivot} and min{|r| > pivot}: the upper middle element of an even-length median
extern "C" __global__ void __launch_bounds__(256) nls_irls_above(const NlsIrlsParams prm)
{
    nls_exp_init();
    NlsThread T;
    nls_irls_theta(prm, T);
    unsigned long long cnt = 0ull, mn = ~0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += stride) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(nls_irls_resid(T, prm, i)));
        if (key <= prm.pivot)
            ++cnt;
        else if (key < mn)
            mn = key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        const unsigned long long m2 = __shfl_down_sync(0xffffffffu, mn, o);
        mn = m2 < mn ? m2 : mn;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(prm.cnt_min, cnt);
        atomicMin(prm.cnt_min + 1, mn);
    }
}

// ---- psi = rho' and psi' = rho'' of the eight losses (src/nls_irls.c:10-330) ----
static __device__ double nls_psi(double x, const double *c, int which, double *dpsi)
{
    const double ax = fabs(x);
    switch (which) {
    default:
    case 1: { // huber
        *dpsi = ax >= c[0] ? 0.0 : 1.0;
        return x <= -c[0] ? -c[0] : (x < c[0] ? x : c[0]);
    }
    case 2: { // barron
        const double alpha = c[0], c2 = c[1] * c[1], x2 = x * x, xc2 = x2 / c2;
        const double se = 1.4901161193847656e-08;
        if (fabs(alpha - 2.0) < se) {
            *dpsi = 1.0 / c2;
            return x / c2;
        } else if (fabs(alpha) < se) {
            *dpsi = -2.0 * (x2 - 2.0 * c2) / ((2.0 * c2 + x2) * (2.0 * c2 + x2));
            return 2.0 * x / (x2 + 2.0 * c2);
        } else if (alpha > -1e8) {
            const double denom = x2 - (alpha - 2.0) * c2;
            *dpsi = (alpha - 2.0) * ((alpha - 2.0) * c2 - (alpha - 1.0) * x2) *
                    pow(1.0 - x2 / ((alpha - 2.0) * c2), 0.5 * alpha) / (denom * denom);
            return x / c2 * pow(xc2 / fabs(alpha - 2.0) + 1.0, 0.5 * alpha - 1.0);
        }
        *dpsi = exp(-x2 / (2.0 * c2)) * (c2 - x2) / (c2 * c2);
        return x / c2 * exp(-0.5 * xc2);
    }
    case 3: { // bisquare
        if (ax > c[0]) {
            *dpsi = 0.0;
            return 0.0;
        }
        const double a = x / c[0], u = 1.0 - a * a, a2 = a * a;
        *dpsi = (1.0 - a2) * (1.0 - 5.0 * a2);
        return x * u * u;
    }
    case 4: { // welsh / Gauss weight
        const double a = x / c[0];
        if (fabs(a) > 37.7) {
            *dpsi = 0.0;
            return 0.0;
        }
        const double e = exp(-(a * a) / 2.0);
        *dpsi = e * (1.0 - a * a);
        return x * e;
    }
    case 5: { // optimal
        const double R1 = -1.944, R2 = 1.728, R3 = -0.312, R4 = 0.016;
        const double ac = x / c[0], aa = fabs(ac);
        if (aa > 3.0) {
            *dpsi = 0.0;
            return 0.0;
        } else if (aa > 2.0) {
            const double a2 = ac * ac;
            *dpsi = R1 + a2 * (3.0 * R2 + a2 * (5.0 * R3 + a2 * 7.0 * R4));
            const double v = c[0] * ((((R4 * a2 + R3) * a2 + R2) * a2 + R1) * ac);
            return ac > 0.0 ? (v > 0.0 ? v : 0.0) : -fabs(v);
        }
        *dpsi = 1.0;
        return x;
    }
    case 6: { // hampel
        const double a = 1.5 * c[0], b = 3.5 * c[0], r = 8.0 * c[0];
        const double sx = x < 0.0 ? -1.0 : 1.0;
        if (ax <= a) {
            *dpsi = 1.0;
            return x;
        } else if (ax <= b) {
            *dpsi = 0.0;
            return sx * a;
        } else if (ax <= r) {
            *dpsi = a / (b - r);
            return sx * a * (r - ax) / (r - b);
        }
        *dpsi = 0.0;
        return 0.0;
    }
    case 7: { // ggw
        if (ax < c[2]) {
            *dpsi = 1.0;
            return x;
        }
        const double ea = -pow(ax - c[2], c[1]) / (2.0 * c[0]);
        if (ea < -708.4) {
            *dpsi = 0.0;
            return 0.0;
        }
        *dpsi = exp(ea) * (1.0 - c[1] / (2.0 * c[0]) * ax * pow(ax - c[2], c[1] - 1.0));
        return x * exp(ea);
    }
    case 8: { // lqq
        if (ax <= c[1]) {
            *dpsi = 1.0;
            return x;
        }
        const double k01 = c[0] + c[1], sg = x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0);
        if (ax <= k01) {
            *dpsi = 1.0 - c[2] / c[0] * (ax - c[1]);
            return sg * (ax - c[2] * (ax - c[1]) * (ax - c[1]) / c[0] / 2.0);
        }
        const double s5 = c[2] - 1.0, s6 = -2.0 * k01 + c[0] * c[2];
        const double aa = (c[0] * c[2] - 2.0 * k01) / (1.0 - c[2]);
        *dpsi = ax < k01 + aa ? -(1.0 - c[2]) * ((ax - k01) / aa - 1.0) : 0.0;
        if (ax < k01 - s6 / s5)
            return (x > 0.0 ? 1.0 : -1.0) * (-s6 / 2.0 - s5 * s5 / s6 * ((ax - k01) * (ax - k01) / 2.0 + s6 / s5 * (ax - k01)));
        return 0.0;
    }
    }
}

extern "C" __global__ void __launch_bounds__(256) nls_irls_weights(const NlsIrlsParams prm)
{
    __shared__ double sw[8];
    nls_exp_init();
    NlsThread T;
    nls_irls_theta(prm, T);
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += stride) {
        const double rs = nls_irls_resid(T, prm, i) / prm.sigma;
        double dpsi;
        const double ps = nls_psi(rs, prm.cc, prm.loss, &dpsi);
        const double q = ps / rs;
        const double wt = q > 2.2204460492503131e-16 ? q : 2.2204460492503131e-16; // gsl_max(psi / r, eps); 0 / 0 -> eps
        prm.wout[i] = wt;
        if (prm.psi)
            prm.psi[i] = ps;
        if (prm.psip)
            prm.psip[i] = dpsi;
        acc += wt;
    }
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w)
            s += sw[w];
        prm.partial[blockIdx.x] = s;
    }
}

extern "C" __global__ void __launch_bounds__(256) nls_irls_scale(const NlsIrlsParams prm)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += stride)
        prm.wout[i] = prm.wout[i] * prm.scale * (prm.userw ? prm.userw[i] : 1.0);
}

// ---------------------------------------------------------------- sparse-row problems: term evaluation
// (csrc/sparse.cu holds the model-independent part: row sums, J d, J^T u, the matrix-free cgst solver)
#if GSLNLS_JAC_MODE == 0 && GSLNLS_P <= NLS_SP_MAXSLOT
#ifdef GSLNLS_JCONST_MASK
#define NLS_JCONST_MASK ((unsigned)GSLNLS_JCONST_MASK)
#else
#define NLS_JCONST_MASK 0u
#endif
extern "C" __global__ void __launch_bounds__(256) nls_sparse_eval(const NlsSparseEvalParams prm)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int badf = 0, badj = 0;
    nls_exp_init();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < prm.nterms; t += stride) {
        double th[NLS_P], xa[NLS_NV], f, J[NLS_P];
#pragma unroll
        for (int s = 0; s < NLS_P; ++s) {
            const int *ix = prm.slot_index[s];
            th[s] = prm.theta[prm.slot_base[s] + (ix ? ix[t] : 0)];
        }
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = prm.vars[k][t];
        nls_model_fj(th, xa, f, J);
        const bool okf = nls_finite(f);
        prm.tv[t] = okf ? f : NLS_INF; // a non-finite value makes the residual +Inf, src/nls_large.c:464-465
        bool okj = true;
#pragma unroll
        for (int s = 0; s < NLS_P; ++s) {
            okj = okj && nls_finite(J[s]);
            if (!(NLS_JCONST_MASK >> s & 1u)) // constant partials were written once when the problem was built
                prm.jv[(long long)s * prm.nterms + t] = J[s];
        }
        badf += okf ? 0u : 1u;
        badj += okj ? 0u : 1u; // -> GSL_EBADFUNC when this Jacobian is adopted, src/nls_large.c:560-566
    }
    if (badf)
        atomicAdd(prm.nbad, (unsigned long long)badf);
    if (badj)
        atomicAdd(prm.nbad + 1, (unsigned long long)badj);
}
#endif
