/*
 * gslnls_b200.h -- C ABI of libgslnls_b200.so: the B200-native gsl_nls_large() hot path.
 *
 * This library replaces exactly one reference entry point and what runs under it:
 *
 *     SEXP C_nls_large(SEXP fn, SEXP y, SEXP jac, SEXP fvv, SEXP env, SEXP start, SEXP weights,
 *                      SEXP control_int, SEXP control_dbl)          -- src/nls_large.c:66
 *     registered at src/init.c:9,16, called from R/nls_large.R:411 and :598,
 *
 * i.e. C_nls_large_internal (src/nls_large.c:77-424), its three callbacks gsl_f_large (:426),
 * gsl_df_large (:474), gsl_fvv_large (:655), callback_large (:715), the driver
 * gsl_multilarge_nlinear_driver2 (src/nls_fit.c:153-224) and libgsl's multilarge_nlinear trust
 * solver underneath.  R closures cannot execute on a GPU, so the (fn, jac, fvv, env) quadruple
 * is replaced by a compiled model (formula text -> symbolic Jacobian / directional second
 * derivative -> NVRTC device code); every other argument crosses the boundary unchanged:
 * y, weights (raw, not sqrt), start, control_int[7], control_dbl[8] exactly as packed at
 * R/nls_large.R:383-407, and the result carries the fields of the list built at
 * src/nls_large.c:276-416.  Status codes are GSL errno values (0 success, 9 EBADFUNC,
 * 11 EMAXITER, 27 ENOPROG, ...), strings are gsl_strerror()'s.
 *
 * Plain C types only; host pointers unless a name says "device".  The library owns all device
 * memory it allocates.  One host thread per problem handle.  There is no CPU compute path:
 * every entry point that needs the GPU fails with GSLNLS_ENODEVICE if none is usable.
 */
#ifndef GSLNLS_B200_H
#define GSLNLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSLNLS_API __attribute__((visibility("default")))

/* return codes beyond GSL's errno range */
enum {
    GSLNLS_SUCCESS = 0,
    GSLNLS_FAILURE = -1,
    GSLNLS_CONTINUE = -2,
    GSLNLS_EDOM = 1,
    GSLNLS_EINVAL = 4,
    GSLNLS_ENOMEM = 8,
    GSLNLS_EBADFUNC = 9,
    GSLNLS_EMAXITER = 11,
    GSLNLS_EBADLEN = 19,
    GSLNLS_ENOPROG = 27,
    GSLNLS_ETOLF = 29,
    GSLNLS_ETOLX = 30,
    GSLNLS_ETOLG = 31,
    GSLNLS_EPARSE = 1001,   /* formula could not be parsed / differentiated */
    GSLNLS_ECOMPILE = 1002, /* NVRTC rejected the generated kernel */
    GSLNLS_ECUDA = 1003,    /* CUDA runtime error (message via gslnls_last_error) */
    GSLNLS_ENODEVICE = 1004,/* no usable CUDA device: there is no CPU fallback */
    GSLNLS_ECOMM = 1005     /* multi-GPU exchange failed */
};

/* jac_mode of gslnls_model_compile */
enum {
    GSLNLS_JAC_SYMBOLIC = 0, /* deriv()-style symbolic Jacobian (R/nls_large.R:296-306) */
    GSLNLS_JAC_FORWARD = 1,  /* forward differences, step rule of src/fdjac.c:24-64 */
    GSLNLS_JAC_CENTER = 2    /* centred differences, src/fdjac.c:81-128 */
};
/* fvv_mode */
enum {
    GSLNLS_FVV_NONE = 0,
    GSLNLS_FVV_SYMBOLIC = 1, /* deriv(hessian=TRUE) contracted with v (R/nls_large.R:333-343) */
    GSLNLS_FVV_FD = 2        /* Transtrum-Sethna Eq. 19, src/fdfvv.c:35-77 */
};

/* how raw `weights` enter the normal equations */
enum {
    /* default: rows of f AND J are scaled by sqrt(w_i) (what src/fdf.c:135-177 does on the multifit path): the
     * packet is that of the weighted problem, g = J^T W f and J^T W J, covar = (J^T W J)^-1 */
    GSLNLS_WEIGHTS_CONSISTENT = 0,
    /* reference-compatible: what gsl_nls_large() computes through gsl_multilarge_nlinear_winit
     * (src/nls_large.c:216-222): libgsl scales f and fvv by sqrt(w_i) but never sees J, and the reference's
     * callback gsl_df_large (:474-653) forms J^T u and J^T J from the UNWEIGHTED Jacobian.  Same inputs, same
     * numbers as the reference for non-unit weights, including the returned `grad` (unweighted, :354-363). */
    GSLNLS_WEIGHTS_GSL = 1
};

typedef struct gslnls_model gslnls_model;     /* compiled model: generated CUDA + cubin, opaque */
typedef struct gslnls_problem gslnls_problem; /* device-resident data + solver workspace, opaque */
typedef struct gslnls_comm gslnls_comm;       /* multi-GPU exchange context, opaque */
typedef struct gslnls_session gslnls_session; /* the data of a fit resident on 1..8 GPUs of this process, opaque */

/* The list returned by C_nls_large (src/nls_large.c:279-288), as a C struct.
 * All pointers are owned by the struct; release with gslnls_result_free. */
typedef struct gslnls_result {
    int64_t n;           /* global number of observations */
    int p;
    double *par;         /* [p]   final parameters, or start on failure (:293-302) */
    double *covar;       /* [p*p] column-major (J^T J)^-1, NaN-filled on failure (:311-326) */
    double ssr;          /* chisq1 (:391) */
    double ssrtol;       /* chisq0 - chisq1 (:392) */
    double chisq_init;   /* ssr at start (:230-235, printed by trace) */
    int niter;           /* gsl_multilarge_nlinear_niter (:248) */
    int conv;            /* status code (:390); R: isConv = !conv */
    int info;            /* convergence reason 1 = step, 2 = gradient (src/nls_fit.c:135-144) */
    const char *status;  /* gsl_strerror(conv) (:389) */
    const char *algorithm; /* trs name (:393) */
    int64_t neval[4];    /* f, dfu, df2, fvv (:395-401): logical GSL counts */
    int64_t npass;       /* physical fused passes over the data that were launched */
    int ntrace;          /* maxiter + 1 if trace else 0 */
    double *partrace;    /* [(maxiter+1) * p] column-major, rows 0..niter valid (:177-182,:719-727) */
    double *ssrtrace;    /* [maxiter+1] */
    double *condtrace;   /* [maxiter+1] cond(J) per iteration as printed by callback_large (:733-738) */
    double *resid;       /* [n_local] weighted f - y (:339-343) if requested, else NULL */
    double *grad;        /* [n_local * p] column-major weighted Jacobian (:354-363) if requested */
    int64_t n_local;
    double *jtj;         /* [p*p] column-major J^T J at par (so R = chol(JTJ)^T replaces the O(n p^2) QR) */
    double *grad_vec;    /* [p] J^T f at par */
    double *x_final;     /* [p] the solver's last accepted iterate w->x, also when conv reports a failure and `par`
                            therefore carries the start values (read by the IRLS driver, src/nls_irls.c:454,515) */
} gslnls_result;

/* ---- model compilation ------------------------------------------------------------------- */

/* Translate the right-hand side of an R model formula into device code.
 *   rhs_expr     e.g. "A * exp(-lam * x) + b"  (R arithmetic: + - * / ^ ** unary minus, parentheses,
 *                exp log log2 log10 log1p expm1 sqrt sin cos tan asin acos atan sinh cosh tanh abs
 *                pnorm dnorm sinpi cospi, the constant pi)
 *   param_names  p names, in the order of `start`
 *   var_names    nvar predictor names, in the order the data columns are passed later
 * Replaces the closures built at R/nls_large.R:273 (.fn), :296-309 (.jac), :333-347 (.fvv). */
GSLNLS_API int gslnls_model_compile(const char *rhs_expr, const char *const *param_names, int p,
                                    const char *const *var_names, int nvar, int jac_mode, int fvv_mode,
                                    gslnls_model **out, char *errbuf, size_t errlen);
GSLNLS_API void gslnls_model_free(gslnls_model *m);
GSLNLS_API int gslnls_model_p(const gslnls_model *m);
GSLNLS_API int gslnls_model_nvar(const gslnls_model *m);
/* generated model source (the device functions, also valid host C++ for CPU-side codegen tests) */
GSLNLS_API const char *gslnls_model_source(const gslnls_model *m);

/* ---- one-shot fit: the drop-in for .Call(C_nls_large, ...) ---------------------------------- */

/* vars: nvar host pointers of n doubles each; y: n; weights: n raw weights or NULL;
 * control_int[7] = {maxiter, trace, algorithm 0..5, scale 0..2, fdtype 0..1, jacclass, jacnz}
 * control_dbl[8] = {factor_up, factor_down, avmax, h_df, h_fvv, xtol, ftol, gtol}
 * Host->device copies of the data happen inside this call. Returns the GSL status (== out->conv)
 * or a GSLNLS_E* library error (out is then zeroed). */
GSLNLS_API int gslnls_fit_large(const gslnls_model *m, const double *const *vars, const double *y,
                                const double *weights, int64_t n, const double *start,
                                const int *control_int, const double *control_dbl, int device,
                                int want_resid_grad, gslnls_result *out);
/* The same call for one rank of a multi-GPU job (one process per GPU): vars/y/weights hold this rank's
 * contiguous shard of n_local rows, `comm` is the exchange context every rank created with
 * gslnls_comm_create (NULL or a 1-rank comm = single GPU).  Every rank calls it with the same start and
 * control and receives the same result; resid/grad, when requested, cover the local rows only. */
GSLNLS_API int gslnls_fit_large_sharded(const gslnls_model *m, const double *const *vars, const double *y,
                                        const double *weights, int64_t n_local, const double *start,
                                        const int *control_int, const double *control_dbl, int device,
                                        gslnls_comm *comm, int want_resid_grad, gslnls_result *out);
/* The same call over several GPUs of the box from ONE process (what an R session is): the rows are split
 * into ngpu contiguous shards, one host thread per GPU uploads its shard over that GPU's own PCIe link and
 * runs the sharded fit; packets cross NVLink peer memory.  devices == NULL means 0..ngpu-1.  The result is
 * the single-GPU result (resid/grad stitched back to n rows when requested).  This is the
 * `int ngpu, const int *devices` form of the replacement for src/nls_large.c:66. */
GSLNLS_API int gslnls_fit_large_multi(const gslnls_model *m, const double *const *vars, const double *y,
                                      const double *weights, int64_t n, const double *start,
                                      const int *control_int, const double *control_dbl, int ngpu,
                                      const int *devices, int want_resid_grad, gslnls_result *out);
GSLNLS_API void gslnls_result_free(gslnls_result *r);

/* ---- sessions: what the host language keeps behind its fitted-model object ------------------------
 * A session is the data of one fit resident on ngpu GPUs of this process (rows split into contiguous shards,
 * one host thread and one PCIe link per GPU, packets over NVLink peer memory).  gslnls_fit_large_multi is
 * create + upload + fit + free; a host object that keeps the session can fit again from other start values
 * or methods without a new upload, and obtains the O(n) outputs of C_nls_large -- `resid` and `grad`,
 * src/nls_large.c:339-385 -- lazily, only when residuals() / the gradient are actually asked for
 * (R/nls.R:1231-1484 builds them eagerly: 3.2 GB of host arrays at n = 1e8, p = 3). */
GSLNLS_API int gslnls_session_create(const gslnls_model *m, int64_t n, int has_weights, int ngpu, const int *devices,
                                     gslnls_session **out);
GSLNLS_API void gslnls_session_free(gslnls_session *s);
GSLNLS_API int gslnls_session_ngpu(const gslnls_session *s); /* GPUs actually used (tiny n uses fewer) */
GSLNLS_API int gslnls_session_set_weights_mode(gslnls_session *s, int mode);
GSLNLS_API int gslnls_session_upload(gslnls_session *s, const double *const *vars, const double *y,
                                     const double *weights);
GSLNLS_API int gslnls_session_fit(gslnls_session *s, const double *start, const int *control_int,
                                  const double *control_dbl, int want_resid_grad, gslnls_result *out);
/* weighted residuals f - y (n) and Jacobian (n x p column-major) at theta from the resident shards; either
 * output may be NULL */
GSLNLS_API int gslnls_session_residuals(gslnls_session *s, const double *theta, double *resid, double *grad_colmajor);

/* gslnls_fit_large[_sharded] keeps the device buffers, workspace and streams of the last call per device
 * for the next one (same model); this returns them (also done for a model by gslnls_model_free).
 * GSLNLS_CACHE=0 in the environment disables the cache. */
GSLNLS_API void gslnls_cache_clear(void);

/* ---- resident-data API (data stays in HBM across fits; used by benchmarks and multi-GPU) ---- */

GSLNLS_API int gslnls_problem_create(const gslnls_model *m, int64_t n_local, int has_weights, int device,
                                     gslnls_problem **out);
GSLNLS_API void gslnls_problem_free(gslnls_problem *pb);
/* copy host data into library-owned device buffers.  Pageable host memory (what R passes) is pinned by the
 * library: worker threads stage slices through pinned buffers while the copy engine drains them (upload.cpp);
 * already pinned / registered memory goes to the copy engine directly.  The host arrays may be released as
 * soon as the call returns. */
GSLNLS_API int gslnls_problem_upload(gslnls_problem *pb, const double *const *vars, const double *y,
                                     const double *weights);
/* use caller-owned device buffers (plain device pointers); they must outlive the problem */
GSLNLS_API int gslnls_problem_bind_device(gslnls_problem *pb, const double *const *dev_vars,
                                          const double *dev_y, const double *dev_weights);
/* weights mode (GSLNLS_WEIGHTS_*) of this problem; the process-wide default for problems created later and for
 * the one-shot calls is set by gslnls_set_weights_mode() or the environment (GSLNLS_WEIGHTS_MODE=gsl) */
GSLNLS_API int gslnls_problem_set_weights_mode(gslnls_problem *pb, int mode);
GSLNLS_API int gslnls_set_weights_mode(int mode);
/* attach an exchange context: this problem holds one shard of a global problem */
GSLNLS_API int gslnls_problem_set_comm(gslnls_problem *pb, gslnls_comm *comm);
GSLNLS_API int gslnls_problem_fit(gslnls_problem *pb, const double *start, const int *control_int,
                                  const double *control_dbl, int want_resid_grad, gslnls_result *out);
/* test / benchmark hooks ------------------------------------------------------------------------ */
/* one fused pass at theta: packet = [J^T J lower packed row-major | J^T f | f^T f], p(p+1)/2+p+1 doubles
 * (summed over all shards when a comm is attached) */
GSLNLS_API int gslnls_problem_eval_packet(gslnls_problem *pb, const double *theta, double *packet);
/* one geodesic pass at theta with velocity v: out = J^T fvv (p doubles) */
GSLNLS_API int gslnls_problem_eval_jtfvv(gslnls_problem *pb, const double *theta, const double *v, double *out);
/* enqueue `npass` fused passes at theta without host synchronisation in between and return the
 * device time per pass in milliseconds measured with CUDA events on the launch stream */
GSLNLS_API int gslnls_problem_time_passes(gslnls_problem *pb, const double *theta, int npass, float *ms_per_pass);
/* weighted residuals f - y and Jacobian at theta (K4); either output may be NULL */
GSLNLS_API int gslnls_problem_residuals(gslnls_problem *pb, const double *theta, double *resid, double *grad_colmajor);
/* run exactly `ntrial` trust-region trial iterations (pass + step) from the current solver state
 * created by gslnls_problem_fit_begin, timing them on the device; for bench.py */
GSLNLS_API int gslnls_problem_fit_begin(gslnls_problem *pb, const double *start, const int *control_int,
                                        const double *control_dbl);
GSLNLS_API int gslnls_problem_fit_run(gslnls_problem *pb, int max_passes, int *done, int64_t *passes_run,
                                      float *device_ms);
GSLNLS_API int gslnls_problem_fit_end(gslnls_problem *pb, int want_resid_grad, gslnls_result *out);
GSLNLS_API int64_t gslnls_problem_launch_count(const gslnls_problem *pb);
/* developer hook: per-CTA phase stamps (globaltimer ns) of the last pass kernel launch, [ctas][32]:
 * 0 entry, 1 request seen, 2 thread 0 done streaming, 3 CTA done streaming, 4 partial written, 5 packet published,
 * 8 + w: warp w done streaming */
GSLNLS_API int gslnls_problem_trace(gslnls_problem *pb, int enable, unsigned long long *out, int cap_ctas, int *nctas);
/* device timers on the solver's own stream (CUDA events): a region timer, and optional event
 * pairs around each fused-pass launch so a benchmark can report the pass kernel's mean duration
 * inside its timed region (roofline) */
GSLNLS_API int gslnls_problem_timer_start(gslnls_problem *pb);
GSLNLS_API int gslnls_problem_timer_stop(gslnls_problem *pb, float *ms);
GSLNLS_API int gslnls_problem_set_profile(gslnls_problem *pb, int max_passes);
GSLNLS_API int gslnls_problem_profile(gslnls_problem *pb, float *avg_pass_ms, int64_t *npasses_timed);

/* device-side clocks of the resident-server mode (globaltimer, averaged over the passes since the last
 * reset): time a pass kernel spends between seeing its request and depositing its packet, and time the
 * trust-region warp spends between a complete packet and the next published request */
GSLNLS_API int gslnls_problem_channel_stats(gslnls_problem *pb, int reset, double *avg_stream_us,
                                            double *avg_step_us, int64_t *npasses);

/* ---- batched multi-start inner kernels (src/nls_mstart.c:75-91: det(J^T J) screen + mstart_p LM
 *      iterations per start point), candidates ride blockIdx.y ----------------------------------- */
GSLNLS_API int gslnls_problem_fit_batch(gslnls_problem *pb, const double *starts /* S*p row-major */, int S,
                                        const int *control_int, const double *control_dbl,
                                        double *par_out /* S*p */, double *ssr_out /* S */,
                                        double *logdet_out /* S: log det(J^T J) at start, or NULL */,
                                        int *conv_out /* S */, int *niter_out /* S */);

/* ---- robust losses: iteratively reweighted least squares (gsl_multifit_nlinear_rho_driver,
 *      src/nls_irls.c:412-546; loss functions and default tuning constants R/nls_rho.R:101-144) -----------
 * loss  1 huber, 2 barron, 3 bisquare, 4 welsh, 5 optimal, 6 hampel, 7 ggw, 8 lqq; cc[3] tuning constants
 * The problem must have been created with has_weights = 1 (uploaded weights = the user's, or ones): the
 * weights column is the IRLS working vector and holds the final robustness weights on return.  Every IRLS
 * iteration is one weighted fit from `start`, then on the device: unweighted residuals -> sigma = 1.4826
 * median |r| by radix select -> w_i = max(psi(r_i / sigma) / (r_i / sigma), eps), normalised to sum n, times
 * the user's weights.  `out` is the last weighted fit. */
typedef struct gslnls_irls_info {
    double sigma;   /* irls_sigma */
    double delta;   /* irls_tol: max |x_k - x_{k-1}| of the last iteration (src/nls.c:589-595) */
    int niter;      /* irls_niter */
    int status;     /* irls_conv: 0 converged, 11 irls_maxiter reached, -1 not run */
} gslnls_irls_info;
GSLNLS_API int gslnls_problem_fit_irls(gslnls_problem *pb, const double *start, const int *control_int,
                                       const double *control_dbl, int loss, const double *cc, int irls_maxiter,
                                       double irls_xtol, gslnls_result *out, gslnls_irls_info *info);
/* the weights column as it stands on the device (after gslnls_problem_fit_irls: the final IRLS weights) */
GSLNLS_API int gslnls_problem_get_weights(gslnls_problem *pb, double *weights);
/* test hook: median of |fn(theta) - y| over the resident rows by the device radix select */
GSLNLS_API int gslnls_problem_median_abs_resid(gslnls_problem *pb, const double *theta, double *median);

/* ---- multi-start global search: the control logic of gsl_multistart_driver (src/nls_mstart.c:24-349) and its
 *      outer loop (src/nls.c:274-399) over the batched kernels above --------------------------------
 * range       2p doubles (lower, upper) per parameter: the sampling ranges (`start` given as ranges)
 * has_range   2p flags: 0 where the user gave no (finite) limit -> that side adapts dynamically; the caller
 *             substitutes the reference's defaults -0.1 / 0.75 there (R/nls.R:412-422)
 * mstart_int  {mstart_n, mstart_p, mstart_q, mstart_s, mstart_maxiter, mstart_maxstart, mstart_minsp}
 *             = control_int[6..12] of the reference's gsl_nls() call (src/nls.c:300-306)
 * mstart_dbl  {mstart_r, mstart_tol} = control_dbl[8..9] (:308-309)
 * The result is the start vector the final fit is launched from (src/nls.c:518-541). */
typedef struct gslnls_mstart_result {
    int p;
    double *par;     /* [p]  best stationary point found (or the fall-back sample) */
    double *range;   /* [2p] sampling ranges at the end */
    double ssr, ssrconv;
    int nsp, nwsp, mstarts; /* stationary points, worse stationary points since the last one, major iterations */
    int status;      /* 0: stopping rule nsp >= minsp && nwsp > r + sqrt(r) nsp met; 11: mstart_maxstart reached */
    int64_t searches; /* local searches run */
} gslnls_mstart_result;
GSLNLS_API int gslnls_problem_multistart(gslnls_problem *pb, const double *range, const int *has_range,
                                         const int *control_int, const double *control_dbl, const int *mstart_int,
                                         const double *mstart_dbl, gslnls_mstart_result *out);
GSLNLS_API void gslnls_mstart_result_free(gslnls_mstart_result *r);
/* test hook: first `count` points (row-major count x dim) of the quasi-random generator behind the sampler:
 * restatements of gsl_qrng_sobol (dim < 41) / gsl_qrng_halton (src/nls.c:277-280) */
GSLNLS_API int gslnls_qrng_points(int dim, int count, double *out);

/* ---- multi-GPU exchange (one process per GPU) ------------------------------------------------ */
#define GSLNLS_COMM_ID_BYTES 128
/* rank 0 creates an id and ships it to the other ranks by any means (MPI, files, a process-group broadcast) */
GSLNLS_API int gslnls_comm_get_unique_id(void *id_bytes /* GSLNLS_COMM_ID_BYTES */);
GSLNLS_API int gslnls_comm_create(const void *id_bytes, int rank, int nranks, int device, gslnls_comm **out);
/* every rank in ONE process (the shape an R session has): out[0..ndev-1] receive one context per device;
 * the mailboxes are reached through CUDA peer access.  Each context is then driven by its own host thread
 * (gslnls_fit_large_multi does that), or handed to gslnls_fit_large_sharded from ndev threads. */
GSLNLS_API int gslnls_comm_create_local(int ndev, const int *devices, gslnls_comm **out);
GSLNLS_API void gslnls_comm_free(gslnls_comm *c);
/* 1 when the ranks exchange packets through NVLink peer memory (each pass kernel deposits its packet
 * directly in every GPU's mailbox and a resident trust-region warp consumes it); 0 when the exchange
 * is an NCCL all-gather between the pass and the step kernel */
GSLNLS_API int gslnls_comm_has_peer_memory(const gslnls_comm *c);
GSLNLS_API int gslnls_comm_rank(const gslnls_comm *c);
GSLNLS_API int gslnls_comm_size(const gslnls_comm *c);

/* ---- measurement hooks (bench.py roofline denominators; not on the product path) -------------- */
/* FP64-pipe issue peak of the device in TFLOP/s: independent DFMA chains, and mma.sync.m8n8k4.f64 (DMMA) */
GSLNLS_API int gslnls_measure_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops);
/* read bandwidth of a `bytes`-sized device buffer (16-byte streaming loads, best of 6), GB/s */
GSLNLS_API int gslnls_measure_read_bandwidth(int device, size_t bytes, double *gb_per_s);

/* ---- misc -------------------------------------------------------------------------------------- */
GSLNLS_API const char *gslnls_strerror(int code);   /* gsl_strerror() strings + library errors */
GSLNLS_API const char *gslnls_trs_name(int algorithm);
/* ---- sparse-row problems: gsl_nls_large() with a sparse Jacobian ---------------------------------------
 * Replaces the dgT/dgC/dgRMatrix branches of gsl_df_large (src/nls_large.c:528-623: triplet rebuild per callback,
 * :635-648: gsl_spblas_dgemv and densify + dsyrk) for models with many parameters of which a row touches a few
 * (README Example 4, inst/unit_tests/unit_tests_gslnls.R:302-346).  R closures cannot run on the device, so the
 * sparsity structure is data:
 *   - the problem has p_total parameters and nrows residual rows;
 *   - a BLOCK is a compiled row formula (gslnls_model_compile, symbolic Jacobian, k <= 16 local parameters) over
 *     nterms terms.  Local parameter s of term t is theta[slot_base[s] + (slot_index[s] ? slot_index[s][t] : 0)];
 *     vars are the block's data columns (nterms doubles each);
 *   - term t of the block belongs to row rows[t] (rows == NULL: row0 + t); a row is the sum of its terms minus
 *     y[row], times sqrt(weights[row]).  A dense row such as sum(theta^2) - 0.25 is p one-parameter terms that
 *     share a row.
 * The solver is GSL's multilarge trust region, driver and convergence tests of src/nls_fit.c:153-224.  algorithm 5
 * (cgst, Steihaug-Toint) applies J d and J^T u from the stored nonzeros, never a dense J or J^T J, scaling more /
 * levenberg / marquardt, any number of parameters.  Algorithms 0, 2, 3, 4 (lm, dogleg, ddogleg, subspace2D) factor
 * a dense J^T J: for p_total <= 100 the library assembles the dense packet from the nonzeros and runs the dense
 * trust-region kernel (the reference densifies J for the same purpose, src/nls_large.c:641-648); algorithm 1
 * (lmaccel) is refused (blocks carry no fvv).  control_int / control_dbl are those of gslnls_fit_large. */
typedef struct gslnls_sparse_problem gslnls_sparse_problem;
typedef struct gslnls_sparse_result {
    int p;
    int64_t nrows, nterms, nnz; /* nnz: stored nonzeros of J */
    double *par;        /* [p] */
    double ssr, ssrtol, chisq_init;
    int niter, conv, info;
    const char *status;
    int64_t neval[4];   /* f, dfu, df2, fvv: logical GSL counts (dfu: every J d / J^T u product) */
    int64_t cg_iters;   /* Steihaug-Toint iterations over the whole fit */
    int64_t launches;   /* solver kernel launches (one per trial point) */
    double eval_ms, solver_ms; /* device time (CUDA events) of the term-evaluation and of the solver launches */
    int ntrace;
    double *ssrtrace;   /* [maxiter + 1] when control_int[1] */
    double *grad_vec;   /* [p] J^T f at par */
    double *jtj;        /* [p*p] column-major J^T J at par when want_jtj (R = chol(jtj) for summary / vcov) */
    double *resid;      /* [nrows] weighted residuals when want_resid */
} gslnls_sparse_result;
GSLNLS_API int gslnls_sparse_create(int device, int p_total, int64_t nrows, gslnls_sparse_problem **out);
GSLNLS_API void gslnls_sparse_free(gslnls_sparse_problem *sp);
GSLNLS_API int gslnls_sparse_add_block(gslnls_sparse_problem *sp, const gslnls_model *m, int64_t nterms,
                                       const double *const *vars, const int *slot_base,
                                       const int *const *slot_index, const int *rows, int64_t row0);
GSLNLS_API int gslnls_sparse_set_response(gslnls_sparse_problem *sp, const double *y, const double *weights);
/* builds the row / column gather lists and the device workspace; no blocks can be added afterwards */
GSLNLS_API int gslnls_sparse_finalize(gslnls_sparse_problem *sp);
GSLNLS_API int64_t gslnls_sparse_nnz(const gslnls_sparse_problem *sp);
GSLNLS_API int gslnls_sparse_fit(gslnls_sparse_problem *sp, const double *start, const int *control_int,
                                 const double *control_dbl, int want_jtj, int want_resid, gslnls_sparse_result *out);
/* weighted residuals [nrows], J^T f [p], diag(J^T J) [p] and f^T f at theta; any output may be NULL */
GSLNLS_API int gslnls_sparse_eval(gslnls_sparse_problem *sp, const double *theta, double *resid, double *grad_vec,
                                  double *jtj_diag, double *ssr);
GSLNLS_API void gslnls_sparse_result_free(gslnls_sparse_result *r);

GSLNLS_API const char *gslnls_last_error(void);     /* thread-local detail of the last GSLNLS_E* */
GSLNLS_API int gslnls_device_count(void);
GSLNLS_API const char *gslnls_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSLNLS_B200_H */
