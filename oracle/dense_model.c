/*
 * oracle/dense_model.c -- restatement of the callbacks and the .Call body of
 * /root/reference/src/nls_large.c on top of a plain "row evaluator".
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 *   orc_dense_f    <- gsl_f_large    src/nls_large.c:426-472
 *   orc_dense_df   <- gsl_df_large   src/nls_large.c:474-653 (dense branch :504-527, :625-634)
 *   orc_dense_fvv  <- gsl_fvv_large  src/nls_large.c:655-713
 *   orc_nls_large  <- C_nls_large_internal src/nls_large.c:77-424
 *
 * The data flow is the reference's on purpose (it is also the timed CPU baseline): f is
 * materialised as an n-vector, J as an n x p row-major matrix, then J^T u and J^T J are
 * formed by sequential-order dgemv / dsyrk loops (gslcblas ordering).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int is_bad(double v) { return isnan(v) || !isfinite(v); }

/* ---- model evaluation, optionally split over threads (rows are independent) ------------- */

typedef struct {
    const double *x;
} xonly;

static int rows_call(const orc_dense_model *m, const double *theta, const double *v, double *fval, double *J,
                     double *fvv)
{
#ifdef _OPENMP
    if (m->threads > 1 && m->data && m->rows != NULL) {
        /* only the built-in evaluators (data = orc_xdata) are safe to slice */
        if (m->rows == orc_rows_exp3 || m->rows == orc_rows_gauss || m->rows == orc_rows_gaussmix ||
            m->rows == orc_rows_expmix2) {
            const orc_xdata *xd = (const orc_xdata *)m->data;
            int bad = 0;
#pragma omp parallel num_threads(m->threads) reduction(| : bad)
            {
                const int T = omp_get_num_threads(), t = omp_get_thread_num();
                const size_t lo = m->n * (size_t)t / (size_t)T, hi = m->n * (size_t)(t + 1) / (size_t)T;
                orc_xdata sub;
                sub.x = xd->x + lo;
                bad |= m->rows(theta, v, hi - lo, m->p, &sub, fval ? fval + lo : NULL,
                               J ? J + lo * m->p : NULL, fvv ? fvv + lo : NULL);
            }
            return bad ? ORC_EBADFUNC : ORC_SUCCESS;
        }
    }
#endif
    return m->rows(theta, v, m->n, m->p, m->data, fval, J, fvv);
}

/* ---- gsl_f_large -------------------------------------------------------------------------- */

int orc_dense_f(const double *x, void *params, double *f)
{
    orc_dense_model *m = (orc_dense_model *)params;
    int s = rows_call(m, x, NULL, f, NULL, NULL);
    if (s)
        return ORC_EBADFUNC;
    /* src/nls_large.c:462-468 */
    for (size_t i = 0; i < m->n; ++i) {
        if (is_bad(f[i]))
            f[i] = INFINITY;
        else
            f[i] = f[i] - m->y[i];
    }
    return ORC_SUCCESS;
}

/* ---- Jacobian: analytic or src/fdjac.c rules ---------------------------------------------- */

static int fill_jacobian(orc_dense_model *m, const double *x, int count)
{
    const size_t n = m->n, p = m->p;
    if (m->fd_jac == 0)
        return rows_call(m, x, NULL, NULL, m->J, NULL) ? ORC_EBADFUNC : ORC_SUCCESS;

    double *xp = (double *)malloc(p * sizeof(double));
    double *f0 = (double *)malloc(n * sizeof(double));
    double *f1 = (double *)malloc(n * sizeof(double));
    int s = 0;
    memcpy(xp, x, p * sizeof(double));
    if (m->fd_jac == 1) {
        s = rows_call(m, x, NULL, f0, NULL, NULL); /* the stored f(x) of multifit; not an extra eval */
        for (size_t j = 0; j < p && !s; ++j) {
            const double xj = x[j];
            double delta = m->h_df * fabs(xj);
            if (delta == 0.0)
                delta = m->h_df;
            xp[j] = xj + delta;
            s = rows_call(m, xp, NULL, f1, NULL, NULL);
            if (count)
                ++m->fd_nevalf;
            xp[j] = xj;
            delta = 1.0 / delta;
            for (size_t i = 0; i < n; ++i)
                m->J[i * p + j] = (f1[i] - f0[i]) * delta;
        }
    } else {
        for (size_t j = 0; j < p && !s; ++j) {
            const double xj = x[j];
            double delta = m->h_df * fabs(xj);
            if (delta == 0.0)
                delta = m->h_df;
            xp[j] = xj + 0.5 * delta;
            s = rows_call(m, xp, NULL, f1, NULL, NULL);
            xp[j] = xj - 0.5 * delta;
            s |= rows_call(m, xp, NULL, f0, NULL, NULL);
            if (count)
                m->fd_nevalf += 2;
            xp[j] = xj;
            delta = 1.0 / delta;
            for (size_t i = 0; i < n; ++i)
                m->J[i * p + j] = (f1[i] - f0[i]) * delta;
        }
    }
    free(xp); free(f0); free(f1);
    return s ? ORC_EBADFUNC : ORC_SUCCESS;
}

/* ---- gsl_df_large ---------------------------------------------------------------------------- */

int orc_dense_df(int TransJ, const double *x, const double *u, void *params, double *v, double *JTJ)
{
    orc_dense_model *m = (orc_dense_model *)params;
    const size_t n = m->n, p = m->p;
    double *J = m->J;

    /* multifit keeps J(x); only a (g, J^T J) refresh is a new Jacobian evaluation */
    if (fill_jacobian(m, x, JTJ != NULL))
        return ORC_EBADFUNC;

    /* src/nls_large.c:515-522: missing/infinite values are not allowed in jac */
    for (size_t i = 0; i < n; ++i)
        for (size_t k = 0; k < p; ++k)
            if (is_bad(J[i * p + k]))
                return ORC_EBADFUNC;

    /* weights: rows scaled by sqrt(w_i) exactly like src/fdf.c:153-167 does for multifit.
       (GSL's multilarge leaves J to the callback; the reference only ever tests unit weights.) */
    if (m->sqrt_wts)
        for (size_t i = 0; i < n; ++i)
            for (size_t k = 0; k < p; ++k)
                J[i * p + k] *= m->sqrt_wts[i];

    if (v) {
        if (TransJ == ORC_TRANS) {
            /* cblas_dgemv(RowMajor, Trans): y_j += (alpha x_i) A[i][j], i outer */
            if (m->longdouble) {
                long double *acc = (long double *)calloc(p, sizeof(long double));
                for (size_t i = 0; i < n; ++i)
                    for (size_t j = 0; j < p; ++j)
                        acc[j] += (long double)u[i] * (long double)J[i * p + j];
                for (size_t j = 0; j < p; ++j)
                    v[j] = (double)acc[j];
                free(acc);
            } else {
#ifdef _OPENMP
                if (m->threads > 1) {
                    const int T = m->threads;
                    double *part = (double *)calloc((size_t)T * p, sizeof(double));
#pragma omp parallel num_threads(T)
                    {
                        const int t = omp_get_thread_num(), TT = omp_get_num_threads();
                        const size_t lo = n * (size_t)t / (size_t)TT, hi = n * (size_t)(t + 1) / (size_t)TT;
                        double *a = part + (size_t)t * p;
                        for (size_t i = lo; i < hi; ++i) {
                            const double temp = u[i];
                            for (size_t j = 0; j < p; ++j)
                                a[j] += temp * J[i * p + j];
                        }
                    }
                    for (size_t j = 0; j < p; ++j) {
                        double s = 0.0;
                        for (int t = 0; t < T; ++t)
                            s += part[(size_t)t * p + j];
                        v[j] = s;
                    }
                    free(part);
                } else
#endif
                {
                    for (size_t j = 0; j < p; ++j)
                        v[j] = 0.0;
                    for (size_t i = 0; i < n; ++i) {
                        const double temp = u[i];
                        for (size_t j = 0; j < p; ++j)
                            v[j] += temp * J[i * p + j];
                    }
                }
            }
        } else {
            for (size_t i = 0; i < n; ++i) {
                double temp = 0.0;
                for (size_t j = 0; j < p; ++j)
                    temp += u[j] * J[i * p + j];
                v[i] = temp;
            }
        }
    }

    if (JTJ) {
        /* cblas_dsyrk(RowMajor, Lower, Trans): C_ij = sum_k A[k][i] A[k][j], j <= i, k sequential */
        const size_t np = p * (p + 1) / 2;
        if (m->longdouble) {
            long double *acc = (long double *)calloc(np, sizeof(long double));
            for (size_t k = 0; k < n; ++k) {
                const double *row = J + k * p;
                size_t e = 0;
                for (size_t i = 0; i < p; ++i)
                    for (size_t j = 0; j <= i; ++j)
                        acc[e++] += (long double)row[i] * (long double)row[j];
            }
            size_t e = 0;
            for (size_t i = 0; i < p; ++i)
                for (size_t j = 0; j <= i; ++j)
                    JTJ[i * p + j] = (double)acc[e++];
            free(acc);
        } else {
            int T = 1;
#ifdef _OPENMP
            if (m->threads > 1)
                T = m->threads;
#endif
            double *part = (double *)calloc((size_t)T * np, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(T) if (T > 1)
#endif
            {
#ifdef _OPENMP
                const int t = omp_get_thread_num(), TT = omp_get_num_threads();
#else
                const int t = 0, TT = 1;
#endif
                const size_t lo = n * (size_t)t / (size_t)TT, hi = n * (size_t)(t + 1) / (size_t)TT;
                double *a = part + (size_t)t * np;
                for (size_t k = lo; k < hi; ++k) {
                    const double *row = J + k * p;
                    size_t e = 0;
                    for (size_t i = 0; i < p; ++i)
                        for (size_t j = 0; j <= i; ++j)
                            a[e++] += row[i] * row[j];
                }
            }
            size_t e = 0;
            for (size_t i = 0; i < p; ++i)
                for (size_t j = 0; j <= i; ++j, ++e) {
                    double s = 0.0;
                    for (int t = 0; t < T; ++t)
                        s += part[(size_t)t * np + e];
                    JTJ[i * p + j] = s;
                }
            free(part);
        }
    }
    return ORC_SUCCESS;
}

/* ---- gsl_fvv_large (analytic) or src/fdfvv.c:35-77 (finite difference) ----------------------- */

int orc_dense_fvv(const double *x, const double *v, void *params, double *fvv)
{
    orc_dense_model *m = (orc_dense_model *)params;
    const size_t n = m->n, p = m->p;
    if (!m->fd_fvv) {
        if (rows_call(m, x, v, NULL, NULL, fvv))
            return ORC_EBADFUNC;
        for (size_t i = 0; i < n; ++i)
            if (is_bad(fvv[i]))
                return ORC_EBADFUNC; /* src/nls_large.c:697-705 */
        return ORC_SUCCESS;
    }
    /* fvv_i = (2/h) ((f_i(x + h v) - f_i(x)) / h - (J v)_i) with the J in use at x */
    {
        const double h = m->h_fvv, hinv = 1.0 / h;
        double *xp = (double *)malloc(p * sizeof(double));
        double *f0 = m->tmp;
        int s;
        for (size_t k = 0; k < p; ++k)
            xp[k] = x[k] + h * v[k];
        s = rows_call(m, xp, NULL, fvv, NULL, NULL);
        ++m->fd_nevalf;
        s |= rows_call(m, x, NULL, f0, NULL, NULL); /* stored f(x) in multifit */
        s |= fill_jacobian(m, x, 0);                /* stored J(x) in multifit */
        free(xp);
        if (s)
            return ORC_EBADFUNC;
        for (size_t i = 0; i < n; ++i) {
            double u = 0.0;
            for (size_t k = 0; k < p; ++k)
                u += m->J[i * p + k] * v[k];
            fvv[i] = (2.0 * hinv) * ((fvv[i] - f0[i]) * hinv - u);
        }
    }
    return ORC_SUCCESS;
}

/* ---- C_nls_large_internal ---------------------------------------------------------------------- */

typedef struct {
    size_t p, nrow;
    double chisq;
    double *partrace, *ssrtrace, *condtrace;
} trace_ctx;

/* callback_large, src/nls_large.c:715-739 (printing dropped) */
static void trace_cb(size_t iter, void *params, const orc_workspace *w)
{
    trace_ctx *t = (trace_ctx *)params;
    const double *x = orc_position(w);
    const double *f = orc_residual(w);
    (void)f;
    t->ssrtrace[iter] = t->chisq;
    for (size_t k = 0; k < t->p; ++k)
        t->partrace[iter + t->nrow * k] = x[k];
    if (t->condtrace) { /* :733-738: cond(J) = 1 / rcond as printed; a failed factorisation leaves rcond = 0 */
        double rcond = 0.0;
        orc_rcond(&rcond, (orc_workspace *)w);
        t->condtrace[iter] = 1.0 / rcond;
    }
}

void orc_fit_result_free(orc_fit_result *r)
{
    if (!r)
        return;
    free(r->par); free(r->covar); free(r->partrace); free(r->ssrtrace); free(r->condtrace); free(r->diag); free(r->jtj); free(r->xfinal);
    memset(r, 0, sizeof(*r));
}

/* driver2 with the chisq hand-over of src/nls_fit.c:180-181 */
static int driver2_traced(size_t maxiter, double xtol, double gtol, double ftol, trace_ctx *tc, int *info,
                          double *chisq0, double *chisq1, orc_workspace *w, size_t n)
{
    int status = ORC_CONTINUE;
    size_t iter = 0;
    do {
        chisq0[0] = chisq1[0];
        status = orc_iterate(w);
        {
            const double *f = orc_residual(w);
            double s = 0.0;
            for (size_t i = 0; i < n; ++i)
                s += f[i] * f[i];
            chisq1[0] = s;
        }
        if (tc)
            tc->chisq = chisq1[0];
        if (status == ORC_EBADFUNC || (status == ORC_ENOPROG && iter == 0)) {
            *info = status;
            return status;
        }
        ++iter;
        if (tc)
            trace_cb(iter, tc, w);
        status = orc_test(xtol, gtol, ftol, info, w);
    } while (status == ORC_CONTINUE && iter < maxiter);
    if (status == ORC_ETOLF || status == ORC_ETOLX || status == ORC_ETOLG) {
        *info = status;
        status = ORC_SUCCESS;
    }
    if (iter >= maxiter && status != ORC_SUCCESS)
        status = ORC_EMAXITER;
    return status;
}

int orc_nls_large(orc_rows_fn rows, void *data, const double *y, const double *weights, size_t n,
                  const double *start, size_t p, int have_fvv, const int *control_int,
                  const double *control_dbl, const orc_opts *opts, orc_fit_result *out, double *resid,
                  double *grad)
{
    const size_t niter = (size_t)control_int[0];
    const int verbose = control_int[1];
    orc_params P = orc_default_parameters();
    orc_dense_model m;
    orc_fdf fdf;
    orc_workspace *w;
    trace_ctx tc;
    double xtol, ftol, gtol, chisq_init, chisq0, chisq1;
    int info = ORC_CONTINUE, status;

    memset(out, 0, sizeof(*out));
    /* src/nls_large.c:97-116 */
    switch (control_int[2]) {
    case 1: P.trs = ORC_TRS_LMACCEL; break;
    case 2: P.trs = ORC_TRS_DOGLEG; break;
    case 3: P.trs = ORC_TRS_DDOGLEG; break;
    case 4: P.trs = ORC_TRS_SUBSPACE2D; break;
    case 5: P.trs = ORC_TRS_CGST; break;
    default: P.trs = ORC_TRS_LM;
    }
    /* :119-129 */
    switch (control_int[3]) {
    case 1: P.scale = ORC_SCALE_LEVENBERG; break;
    case 2: P.scale = ORC_SCALE_MARQUARDT; break;
    default: P.scale = ORC_SCALE_MORE;
    }
    P.fdtype = control_int[4] ? 1 : 0;
    /* :135-142 */
    P.factor_up = control_dbl[0];
    P.factor_down = control_dbl[1];
    P.avmax = control_dbl[2];
    P.h_df = control_dbl[3];
    P.h_fvv = control_dbl[4];
    xtol = control_dbl[5];
    ftol = control_dbl[6];
    gtol = control_dbl[7];

    memset(&m, 0, sizeof(m));
    m.rows = rows; m.data = data; m.y = y; m.n = n; m.p = p;
    m.J = (double *)malloc(((n && p) ? n * p : 1) * sizeof(double)); /* :167 */
    m.tmp = (double *)malloc((n ? n : 1) * sizeof(double));
    m.longdouble = opts ? opts->longdouble : 0;
    m.threads = opts ? opts->threads : 0;
    m.fd_jac = opts ? opts->fd_jac : 0;
    m.fd_fvv = opts ? opts->fd_fvv : 0;
    m.h_df = P.h_df;
    m.h_fvv = P.h_fvv;

    fdf.f = orc_dense_f;
    fdf.df = orc_dense_df;
    fdf.fvv = (have_fvv || m.fd_fvv) ? orc_dense_fvv : NULL;
    fdf.n = n; fdf.p = p; fdf.params = &m;

    w = orc_alloc(&P, n, p);
    if (weights && !(opts && opts->weights_gsl)) { /* reference mode: gsl_df_large never weights J */
        double *sw = (double *)malloc(n * sizeof(double));
        for (size_t i = 0; i < n; ++i)
            sw[i] = sqrt(weights[i]);
        m.sqrt_wts = sw;
    }
    status = orc_winit(start, weights, &fdf, w);

    out->par = (double *)calloc(p ? p : 1, sizeof(double));
    out->covar = (double *)calloc(p ? p * p : 1, sizeof(double));
    if (verbose) {
        out->partrace = (double *)calloc((niter + 1) * p, sizeof(double));
        out->ssrtrace = (double *)calloc(niter + 1, sizeof(double));
        out->condtrace = (double *)calloc(niter + 1, sizeof(double));
    }

    if (status == ORC_SUCCESS) {
        /* :229-243 */
        const double *r = orc_residual(w);
        chisq_init = 0.0;
        for (size_t i = 0; i < n; ++i)
            chisq_init += r[i] * r[i];
        chisq0 = chisq1 = chisq_init;
        tc.p = p; tc.nrow = niter + 1; tc.chisq = chisq_init;
        tc.partrace = out->partrace; tc.ssrtrace = out->ssrtrace; tc.condtrace = out->condtrace;
        if (verbose) {
            out->ssrtrace[0] = chisq_init;
            for (size_t k = 0; k < p; ++k)
                out->partrace[(niter + 1) * k] = start[k];
        }
        status = driver2_traced(niter, xtol, gtol, ftol, verbose ? &tc : NULL, &info, &chisq0, &chisq1, w, n);
    } else {
        chisq_init = chisq0 = chisq1 = INFINITY;
        info = status;
    }

    out->chisq_init = chisq_init;
    out->diag = (double *)calloc(p ? p : 1, sizeof(double));
    out->jtj = (double *)calloc(p ? p * p : 1, sizeof(double));
    memcpy(out->diag, orc_diag(w), p * sizeof(double));
    memcpy(out->jtj, orc_JTJ(w), p * p * sizeof(double));
    out->xfinal = (double *)calloc(p ? p : 1, sizeof(double));
    memcpy(out->xfinal, orc_position(w), p * sizeof(double));
    out->niter = (int)orc_niter(w);
    out->conv = status;
    out->info = info;
    out->ssr = chisq1;
    out->ssrtol = chisq0 - chisq1;
    out->neval[0] = fdf.nevalf + m.fd_nevalf;
    out->neval[1] = fdf.nevaldfu;
    out->neval[2] = fdf.nevaldf2;
    out->neval[3] = fdf.nevalfvv;

    if (status == ORC_SUCCESS || status == ORC_EMAXITER) {
        double *cov = (double *)malloc(p * p * sizeof(double));
        const double *x = orc_position(w);
        const double *r = orc_residual(w);
        int cs = orc_covar(cov, w); /* :251-256 */
        for (size_t k = 0; k < p; ++k)
            out->par[k] = x[k];
        for (size_t k1 = 0; k1 < p; ++k1)
            for (size_t k2 = 0; k2 < p; ++k2)
                out->covar[k1 + p * k2] = cs ? NAN : cov[k1 * p + k2];
        free(cov);
        if (resid)
            for (size_t i = 0; i < n; ++i)
                resid[i] = r[i];
        if (grad) /* params.J holds the Jacobian of the last df call: n x p -> column-major, :360-362 */
            for (size_t i = 0; i < n; ++i)
                for (size_t k = 0; k < p; ++k)
                    grad[i + n * k] = m.J[i * p + k];
    } else {
        /* :298-302, :319-326, :345-349, :371-376 */
        for (size_t k = 0; k < p; ++k)
            out->par[k] = start[k];
        for (size_t k = 0; k < p * p; ++k)
            out->covar[k] = NAN;
        if (resid)
            for (size_t i = 0; i < n; ++i)
                resid[i] = NAN;
        if (grad)
            for (size_t i = 0; i < n * p; ++i)
                grad[i] = NAN;
    }

    orc_free(w);
    free(m.J); free(m.tmp);
    free((void *)m.sqrt_wts);
    return status;
}

int orc_eval_packet(orc_rows_fn rows, void *data, const double *y, const double *weights, size_t n,
                    size_t p, const double *theta, const orc_opts *opts, double *packet)
{
    orc_dense_model m;
    double *f = (double *)malloc((n ? n : 1) * sizeof(double));
    double *JTJ = (double *)calloc(p ? p * p : 1, sizeof(double));
    double *sw = NULL;
    int s;
    memset(&m, 0, sizeof(m));
    m.rows = rows; m.data = data; m.y = y; m.n = n; m.p = p;
    m.J = (double *)malloc(((n && p) ? n * p : 1) * sizeof(double));
    m.tmp = (double *)malloc((n ? n : 1) * sizeof(double));
    m.longdouble = opts ? opts->longdouble : 0;
    m.threads = opts ? opts->threads : 0;
    m.fd_jac = opts ? opts->fd_jac : 0;
    m.h_df = 1.4901161193847656e-08;
    if (weights) {
        sw = (double *)malloc(n * sizeof(double));
        for (size_t i = 0; i < n; ++i)
            sw[i] = sqrt(weights[i]);
        if (!(opts && opts->weights_gsl))
            m.sqrt_wts = sw;
    }
    s = orc_dense_f(theta, &m, f);
    if (!s && sw)
        for (size_t i = 0; i < n; ++i)
            f[i] *= sw[i];
    if (!s)
        s = orc_dense_df(ORC_TRANS, theta, f, &m, packet + p * (p + 1) / 2, JTJ);
    if (!s) {
        size_t e = 0;
        for (size_t i = 0; i < p; ++i)
            for (size_t j = 0; j <= i; ++j)
                packet[e++] = JTJ[i * p + j];
        if (m.longdouble) {
            long double acc = 0.0L;
            for (size_t i = 0; i < n; ++i)
                acc += (long double)f[i] * (long double)f[i];
            packet[p * (p + 1) / 2 + p] = (double)acc;
        } else {
            double acc = 0.0;
            for (size_t i = 0; i < n; ++i)
                acc += f[i] * f[i];
            packet[p * (p + 1) / 2 + p] = acc;
        }
    }
    free(f); free(JTJ); free(m.J); free(m.tmp); free(sw);
    return s;
}

/* ---- built-in row evaluators --------------------------------------------------------------------- */

int orc_rows_exp3(const double *th, const double *v, size_t n, size_t p, void *data, double *fval, double *J,
                  double *fvv)
{
    const double *x = ((const orc_xdata *)data)->x;
    const double A = th[0], lam = th[1], b = th[2];
    (void)p;
    for (size_t i = 0; i < n; ++i) {
        const double e = exp(-lam * x[i]);
        if (fval)
            fval[i] = A * e + b;
        if (J) {
            J[i * 3 + 0] = e;
            J[i * 3 + 1] = -(A * (e * x[i]));
            J[i * 3 + 2] = 1.0;
        }
        if (fvv) /* d2/dA dlam = -x e ; d2/dlam2 = A x^2 e */
            fvv[i] = 2.0 * v[0] * v[1] * (-(x[i] * e)) + v[1] * v[1] * (A * x[i] * x[i] * e);
    }
    return 0;
}

int orc_rows_gauss(const double *th, const double *v, size_t n, size_t p, void *data, double *fval, double *J,
                   double *fvv)
{
    const double *x = ((const orc_xdata *)data)->x;
    const double a = th[0], b = th[1], c = th[2];
    (void)p;
    for (size_t i = 0; i < n; ++i) {
        const double z = (x[i] - b) / c;
        const double e = exp(-0.5 * z * z);
        if (fval)
            fval[i] = a * e;
        if (J) {
            J[i * 3 + 0] = e;
            J[i * 3 + 1] = a * e * z / c;
            J[i * 3 + 2] = a * e * z * z / c;
        }
        if (fvv) /* README.md Example 2 Hessian */
            fvv[i] = 2.0 * v[0] * v[1] * z / c * e + 2.0 * v[0] * v[2] * z * z / c * e -
                     v[1] * v[1] * a / (c * c) * (1.0 - z * z) * e -
                     2.0 * v[1] * v[2] * a / (c * c) * z * (2.0 - z * z) * e -
                     v[2] * v[2] * a / (c * c) * z * z * (3.0 - z * z) * e;
    }
    return 0;
}

/* sum_k a_k exp(-(x - m_k)^2 / s_k^2), parameters ordered (a,m,s) per component */
int orc_rows_gaussmix(const double *th, const double *v, size_t n, size_t p, void *data, double *fval,
                      double *J, double *fvv)
{
    const double *x = ((const orc_xdata *)data)->x;
    const size_t K = p / 3;
    for (size_t i = 0; i < n; ++i) {
        double f = 0.0, h = 0.0;
        for (size_t k = 0; k < K; ++k) {
            const double a = th[3 * k], mu = th[3 * k + 1], s = th[3 * k + 2];
            const double d = x[i] - mu;
            const double e = exp(-(d * d) / (s * s));
            f += a * e;
            if (J) {
                J[i * p + 3 * k] = e;
                J[i * p + 3 * k + 1] = a * e * 2.0 * d / (s * s);
                J[i * p + 3 * k + 2] = a * e * 2.0 * d * d / (s * s * s);
            }
            if (fvv) {
                const double va = v[3 * k], vm = v[3 * k + 1], vs = v[3 * k + 2];
                const double s2 = s * s;
                const double q = d * d / s2;
                const double fam = e * 2.0 * d / s2;
                const double fas = e * 2.0 * d * d / (s2 * s);
                const double fmm = a * e * (4.0 * q - 2.0) / s2;
                const double fms = a * e * (4.0 * d / (s2 * s)) * (q - 1.0);
                const double fss = a * e * (d * d / (s2 * s2)) * (4.0 * q - 6.0);
                h += 2.0 * va * vm * fam + 2.0 * va * vs * fas + vm * vm * fmm + 2.0 * vm * vs * fms + vs * vs * fss;
            }
        }
        if (fval)
            fval[i] = f;
        if (fvv)
            fvv[i] = h;
    }
    return 0;
}

int orc_rows_expmix2(const double *th, const double *v, size_t n, size_t p, void *data, double *fval,
                     double *J, double *fvv)
{
    const double *x = ((const orc_xdata *)data)->x;
    (void)p;
    for (size_t i = 0; i < n; ++i) {
        const double e1 = exp(-th[1] * x[i]), e2 = exp(-th[3] * x[i]);
        if (fval)
            fval[i] = th[0] * e1 + th[2] * e2;
        if (J) {
            J[i * 4 + 0] = e1;
            J[i * 4 + 1] = -(th[0] * (e1 * x[i]));
            J[i * 4 + 2] = e2;
            J[i * 4 + 3] = -(th[2] * (e2 * x[i]));
        }
        if (fvv)
            fvv[i] = 2.0 * v[0] * v[1] * (-(x[i] * e1)) + v[1] * v[1] * (th[0] * x[i] * x[i] * e1) +
                     2.0 * v[2] * v[3] * (-(x[i] * e2)) + v[3] * v[3] * (th[2] * x[i] * x[i] * e2);
    }
    return 0;
}
