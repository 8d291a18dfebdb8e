/*
 * oracle/multilarge.c -- CPU restatement of GSL's multilarge_nlinear trust-region solver
 * as driven by gslnls' gsl_nls_large() path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * GSL itself (third-party, >= 2.3, Docker pin 2.8) is not vendored in /root/reference; the
 * algorithm is restated from its published sources with the in-tree multifit siblings as
 * the line-level specification.  Each function cites what it follows:
 *   reference src/trust.c, src/fdf.c, src/fdjac.c, src/fdfvv.c, src/nls_fit.c, src/nls_large.c
 *   GSL multilarge_nlinear/<file>.c (by name only; not in tree).
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define MINV(a, b) ((a) < (b) ? (a) : (b))

/* ----------------------------------------------------------------------------------------- */
/* small dense kernels with gslcblas' loop order                        */
/* ----------------------------------------------------------------------------------------- */

static double v_dot(size_t n, const double *a, const double *b)
{
    double r = 0.0;
    for (size_t i = 0; i < n; ++i)
        r += a[i] * b[i];
    return r;
}

/* reference-BLAS dnrm2 (scaled sum of squares) */
static double v_nrm2(size_t n, const double *x)
{
    double scale = 0.0, ssq = 1.0;
    if (n == 0)
        return 0.0;
    if (n == 1)
        return fabs(x[0]);
    for (size_t i = 0; i < n; ++i) {
        if (x[i] != 0.0) {
            const double ax = fabs(x[i]);
            if (scale < ax) {
                ssq = 1.0 + ssq * (scale / ax) * (scale / ax);
                scale = ax;
            } else {
                ssq += (ax / scale) * (ax / scale);
            }
        }
    }
    return scale * sqrt(ssq);
}

/* || diag(D) a ||  -- src/trust.c:48-65 */
static double scaled_enorm(size_t n, const double *D, const double *a)
{
    double e2 = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double u = D[i] * a[i];
        e2 += u * u;
    }
    return sqrt(e2);
}

/* z = alpha x + beta y -- src/trust.c:202-215 */
static void scaled_addition(size_t n, double alpha, const double *x, double beta, const double *y, double *z)
{
    for (size_t i = 0; i < n; ++i)
        z[i] = alpha * x[i] + beta * y[i];
}

/* y = A x, A symmetric with the lower triangle stored (row-major) -- cblas_dsymv(Lower) */
static void symv_lower(size_t p, const double *A, const double *x, double *y)
{
    for (size_t i = 0; i < p; ++i)
        y[i] = 0.0;
    for (size_t i = 0; i < p; ++i) {
        double t1 = x[i], t2 = 0.0;
        for (size_t j = 0; j < i; ++j) {
            y[j] += t1 * A[i * p + j];
            t2 += A[i * p + j] * x[j];
        }
        y[i] += t1 * A[i * p + i] + t2;
    }
}

/* gsl_linalg_cholesky_decomp1 (level-2 variant): lower factor in place, EDOM if not p.d. */
static int cholesky_decomp1(size_t p, double *A)
{
    for (size_t j = 0; j < p; ++j) {
        for (size_t i = j; i < p; ++i) {
            double s = A[i * p + j];
            for (size_t k = 0; k < j; ++k)
                s -= A[i * p + k] * A[j * p + k];
            A[i * p + j] = s;
        }
        double ajj = A[j * p + j];
        if (!(ajj > 0.0))
            return ORC_EDOM;
        ajj = sqrt(ajj);
        const double inv = 1.0 / ajj;
        for (size_t i = j; i < p; ++i)
            A[i * p + j] *= inv;
    }
    return ORC_SUCCESS;
}

/* gsl_linalg_cholesky_solve: L c = b, L^T x = c */
static void cholesky_solve(size_t p, const double *L, const double *b, double *x)
{
    if (x != b)
        memcpy(x, b, p * sizeof(double));
    for (size_t i = 0; i < p; ++i) {
        double t = x[i];
        for (size_t j = 0; j < i; ++j)
            t -= L[i * p + j] * x[j];
        x[i] = t / L[i * p + i];
    }
    for (size_t ii = p; ii-- > 0;) {
        x[ii] = x[ii] / L[ii * p + ii];
        for (size_t j = 0; j < ii; ++j)
            x[j] -= L[ii * p + j] * x[ii];
    }
}

/* gsl_linalg_cholesky_invert: full symmetric inverse from the factor */
static void cholesky_invert(size_t p, double *A)
{
    /* invert L in place */
    for (size_t j = 0; j < p; ++j) {
        A[j * p + j] = 1.0 / A[j * p + j];
        for (size_t i = j + 1; i < p; ++i) {
            double s = 0.0;
            for (size_t k = j; k < i; ++k)
                s += A[i * p + k] * A[k * p + j];
            A[i * p + j] = -s / A[i * p + i];
        }
    }
    /* A^{-1} = L^{-T} L^{-1}; lower triangle then mirror */
    for (size_t i = 0; i < p; ++i)
        for (size_t j = 0; j <= i; ++j) {
            double s = 0.0;
            for (size_t k = i; k < p; ++k)
                s += A[k * p + i] * A[k * p + j];
            A[i * p + j] = s;
        }
    /* rows were overwritten in increasing i while later rows (k >= i) were still L^{-1}:
       row i only reads rows k >= i of column i and j <= i, and writes row i cols <= i.
       Column i of rows k > i is still intact because those rows are written later. */
    for (size_t i = 0; i < p; ++i)
        for (size_t j = i + 1; j < p; ++j)
            A[i * p + j] = A[j * p + i];
}

/* ----------------------------------------------------------------------------------------- */
/* workspace                                                                                   */
/* ----------------------------------------------------------------------------------------- */

struct orc_workspace {
    orc_params params;
    size_t n, p;
    orc_fdf *fdf;
    /* gsl_multilarge_nlinear_workspace */
    double *x, *f, *dx, *g, *JTJ, *sqrt_wts_buf;
    const double *sqrt_wts;
    size_t niter;
    /* trust_state_t (src/gsl_nls.h:110-127) */
    double *diag, *x_trial, *f_trial, *workp, *workn;
    double mu, delta, avratio;
    long nu;
    /* cholesky solver state */
    double *sol_JTJ, *sol_work;
    /* lm state (src/gsl_nls.h:130-142) */
    double *vel, *acc, *JTfvv, *fvv, *lm_workp;
    /* dogleg / subspace2D */
    double *dx_gn, *dx_sd, *workp1, *workp2;
    double norm_Dgn, norm_Dsd, norm_Dinvg, norm_JDinv2g;
    double *W, tau[2], subg[2], subB[4], trB, detB, normg, term0, term1;
    size_t perm[2], rank;
    /* cgst */
    double *cg_z, *cg_r, *cg_d, cg_norm_g;
};

orc_params orc_default_parameters(void)
{
    /* gsl_multilarge_nlinear_default_parameters() */
    orc_params P;
    P.trs = ORC_TRS_LM;
    P.scale = ORC_SCALE_MORE;
    P.fdtype = 0;
    P.factor_up = 3.0;
    P.factor_down = 2.0;
    P.avmax = 0.75;
    P.h_df = 1.4901161193847656e-08; /* GSL_SQRT_DBL_EPSILON */
    P.h_fvv = 0.02;
    P.cg_max_iter = 0;
    P.cg_tol = 1.0e-6;
    return P;
}

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

orc_workspace *orc_alloc(const orc_params *params, size_t n, size_t p)
{
    orc_workspace *w = (orc_workspace *)calloc(1, sizeof(*w));
    if (!w)
        return NULL;
    w->params = *params;
    w->n = n;
    w->p = p;
    w->x = dalloc(p); w->f = dalloc(n); w->dx = dalloc(p); w->g = dalloc(p);
    w->JTJ = dalloc(p * p); w->sqrt_wts_buf = dalloc(n);
    w->diag = dalloc(p); w->x_trial = dalloc(p); w->f_trial = dalloc(n);
    w->workp = dalloc(p); w->workn = dalloc(n);
    w->sol_JTJ = dalloc(p * p); w->sol_work = dalloc(p * p);
    w->vel = dalloc(p); w->acc = dalloc(p); w->JTfvv = dalloc(p); w->fvv = dalloc(n); w->lm_workp = dalloc(p);
    w->dx_gn = dalloc(p); w->dx_sd = dalloc(p); w->workp1 = dalloc(p); w->workp2 = dalloc(p);
    w->W = dalloc(2 * p);
    w->cg_z = dalloc(p); w->cg_r = dalloc(p); w->cg_d = dalloc(p);
    return w;
}

void orc_free(orc_workspace *w)
{
    if (!w)
        return;
    free(w->x); free(w->f); free(w->dx); free(w->g); free(w->JTJ); free(w->sqrt_wts_buf);
    free(w->diag); free(w->x_trial); free(w->f_trial); free(w->workp); free(w->workn);
    free(w->sol_JTJ); free(w->sol_work);
    free(w->vel); free(w->acc); free(w->JTfvv); free(w->fvv); free(w->lm_workp);
    free(w->dx_gn); free(w->dx_sd); free(w->workp1); free(w->workp2); free(w->W);
    free(w->cg_z); free(w->cg_r); free(w->cg_d);
    free(w);
}

const double *orc_position(const orc_workspace *w) { return w->x; }
const double *orc_residual(const orc_workspace *w) { return w->f; }
const double *orc_step(const orc_workspace *w) { return w->dx; }
const double *orc_gradient(const orc_workspace *w) { return w->g; }
const double *orc_JTJ(const orc_workspace *w) { return w->JTJ; }
const double *orc_diag(const orc_workspace *w) { return w->diag; }
size_t orc_niter(const orc_workspace *w) { return w->niter; }
double orc_mu(const orc_workspace *w) { return w->mu; }
double orc_delta(const orc_workspace *w) { return w->delta; }
double orc_avratio(const orc_workspace *w) { return w->avratio; }

const char *orc_trs_name(const orc_workspace *w)
{
    /* gsl_multilarge_nlinear_trs_name(); the first two are confirmed by README.md:595,649 */
    switch (w->params.trs) {
    case ORC_TRS_LM: return "levenberg-marquardt";
    case ORC_TRS_LMACCEL: return "levenberg-marquardt+accel";
    case ORC_TRS_DOGLEG: return "dogleg";
    case ORC_TRS_DDOGLEG: return "double-dogleg";
    case ORC_TRS_SUBSPACE2D: return "2D-subspace";
    default: return "steihaug-toint";
    }
}

const char *orc_strerror(int e)
{
    /* gsl_strerror() strings for the codes that can surface on this path */
    switch (e) {
    case ORC_SUCCESS: return "success";
    case ORC_FAILURE: return "failure";
    case ORC_CONTINUE: return "the iteration has not converged yet";
    case ORC_EDOM: return "input domain error";
    case ORC_EINVAL: return "invalid argument supplied by user";
    case ORC_EBADFUNC: return "problem with user-supplied function";
    case ORC_EMAXITER: return "exceeded max number of iterations";
    case ORC_ENOPROG: return "iteration is not making progress towards solution";
    case ORC_ETOLF: return "cannot reach the specified tolerance in F";
    case ORC_ETOLX: return "cannot reach the specified tolerance in X";
    case ORC_ETOLG: return "cannot reach the specified tolerance in gradient";
    default: return "unknown error code";
    }
}

/* ----------------------------------------------------------------------------------------- */
/* eval wrappers -- GSL multilarge_nlinear/fdf.c, weighting as src/fdf.c:94-113,135-177,200-233 */
/* ----------------------------------------------------------------------------------------- */

static int eval_f(orc_workspace *w, const double *x, double *y)
{
    int s = w->fdf->f(x, w->fdf->params, y);
    ++w->fdf->nevalf;
    if (w->sqrt_wts)
        for (size_t i = 0; i < w->n; ++i)
            y[i] *= w->sqrt_wts[i];
    return s;
}

static int eval_df(orc_workspace *w, int TransJ, const double *x, const double *u, double *v, double *JTJ)
{
    int s = w->fdf->df(TransJ, x, u, w->fdf->params, v, JTJ);
    if (v)
        ++w->fdf->nevaldfu;
    if (JTJ)
        ++w->fdf->nevaldf2;
    return s;
}

static int eval_fvv(orc_workspace *w, const double *x, const double *v, double *yvv)
{
    int s = ORC_SUCCESS;
    if (w->fdf->fvv) {
        s = w->fdf->fvv(x, v, w->fdf->params, yvv);
        ++w->fdf->nevalfvv;
    }
    if (w->sqrt_wts)
        for (size_t i = 0; i < w->n; ++i)
            yvv[i] *= w->sqrt_wts[i];
    return s;
}

/* ----------------------------------------------------------------------------------------- */
/* scaling -- GSL multilarge_nlinear/scaling.c                                   */
/* ----------------------------------------------------------------------------------------- */

static void scale_init(const orc_workspace *w, const double *JTJ, double *diag)
{
    const size_t p = w->p;
    for (size_t j = 0; j < p; ++j) {
        const double Jjj = JTJ[j * p + j];
        const double norm = (Jjj <= 0.0) ? 1.0 : sqrt(Jjj);
        switch (w->params.scale) {
        case ORC_SCALE_LEVENBERG: diag[j] = 1.0; break;
        case ORC_SCALE_MARQUARDT: diag[j] = norm; break;
        default: diag[j] = MAXV(0.0, norm); break; /* more: D=0 then max */
        }
    }
}

static void scale_update(const orc_workspace *w, const double *JTJ, double *diag)
{
    const size_t p = w->p;
    for (size_t j = 0; j < p; ++j) {
        const double Jjj = JTJ[j * p + j];
        const double norm = (Jjj <= 0.0) ? 1.0 : sqrt(Jjj);
        switch (w->params.scale) {
        case ORC_SCALE_LEVENBERG: break;
        case ORC_SCALE_MARQUARDT: diag[j] = norm; break;
        default: diag[j] = MAXV(diag[j], norm); break;
        }
    }
}

/* ----------------------------------------------------------------------------------------- */
/* Nielsen updates -- src/trust.c:149-199 with J^T J in place of J                            */
/* ----------------------------------------------------------------------------------------- */

static void nielsen_init(const orc_workspace *w, double *mu, long *nu)
{
    const double mu0 = 1.0e-3;
    const size_t p = w->p;
    double max = -1.0;
    *nu = 2;
    for (size_t j = 0; j < p; ++j) {
        const double dj = w->diag[j];
        const double val = w->JTJ[j * p + j] / (dj * dj);
        max = MAXV(max, val);
    }
    *mu = mu0 * max;
}

static void nielsen_accept(double rho, double *mu, long *nu)
{
    double b;
    *nu = 2;
    b = 2.0 * rho - 1.0;
    b = 1.0 - b * b * b;
    *mu *= MAXV(0.333333333333333, b);
}

static void nielsen_reject(double *mu, long *nu)
{
    *mu *= (double)*nu;
    *nu <<= 1;
}

/* ----------------------------------------------------------------------------------------- */
/* Cholesky normal-equation solver -- GSL multilarge_nlinear/cholesky.c                       */
/* ----------------------------------------------------------------------------------------- */

static void solver_init(orc_workspace *w)
{
    const size_t p = w->p;
    for (size_t i = 0; i < p; ++i)
        for (size_t j = 0; j <= i; ++j)
            w->sol_JTJ[i * p + j] = w->JTJ[i * p + j];
}

static int solver_presolve(orc_workspace *w, double mu)
{
    const size_t p = w->p;
    for (size_t i = 0; i < p; ++i)
        for (size_t j = 0; j <= i; ++j)
            w->sol_work[i * p + j] = w->sol_JTJ[i * p + j];
    for (size_t i = 0; i < p; ++i)
        w->sol_work[i * p + i] += mu * w->diag[i] * w->diag[i];
    return cholesky_decomp1(p, w->sol_work);
}

static void solver_solve(orc_workspace *w, const double *g, double *x)
{
    cholesky_solve(w->p, w->sol_work, g, x);
    for (size_t i = 0; i < w->p; ++i)
        x[i] = -x[i];
}

/* ----------------------------------------------------------------------------------------- */
/* quadratic predicted reduction -- GSL multilarge_nlinear/common.c               */
/* ----------------------------------------------------------------------------------------- */

static double quadratic_preduction(orc_workspace *w, const double *dx)
{
    const double normf = v_nrm2(w->n, w->f);
    const double gTdx = v_dot(w->p, w->g, dx);
    double pred = -2.0 * gTdx / (normf * normf);
    double u;
    symv_lower(w->p, w->JTJ, dx, w->workp);
    u = v_dot(w->p, w->workp, dx);
    pred -= u / (normf * normf);
    return pred;
}

/* ----------------------------------------------------------------------------------------- */
/* LM / LM + geodesic acceleration -- GSL multilarge_nlinear/lm.c, cf. src/trust.c:223-292    */
/* ----------------------------------------------------------------------------------------- */

static int lm_preloop(orc_workspace *w)
{
    solver_init(w);
    return ORC_SUCCESS;
}

static int lm_step(orc_workspace *w, double delta, double *dx)
{
    const size_t p = w->p;
    const int accel = (w->params.trs == ORC_TRS_LMACCEL);
    int status;
    (void)delta;

    status = solver_presolve(w, w->mu);
    if (status)
        return status;
    solver_solve(w, w->g, w->vel);

    if (accel) {
        double anorm, vnorm;
        status = eval_fvv(w, w->x, w->vel, w->fvv);
        if (status)
            return status;
        /* J^T fvv: a full df callback in the reference (src/nls_large.c:629) */
        status = eval_df(w, ORC_TRANS, w->x, w->fvv, w->JTfvv, NULL);
        if (status)
            return status;
        solver_solve(w, w->JTfvv, w->acc);
        anorm = v_nrm2(p, w->acc);
        vnorm = v_nrm2(p, w->vel);
        w->avratio = anorm / vnorm;
    } else {
        for (size_t i = 0; i < p; ++i)
            w->acc[i] = 0.0;
    }
    scaled_addition(p, 1.0, w->vel, 0.5, w->acc, dx);
    return ORC_SUCCESS;
}

/* More' 1978 Eq. 4.4 on the velocity */
static int lm_preduction(orc_workspace *w, const double *dx, double *pred)
{
    const size_t p = w->p;
    const double norm_Dp = scaled_enorm(p, w->diag, w->vel);
    const double normf = v_nrm2(w->n, w->f);
    double norm_Jp, u, v;
    (void)dx;
    symv_lower(p, w->JTJ, w->vel, w->lm_workp);
    norm_Jp = sqrt(v_dot(p, w->lm_workp, w->vel));
    u = norm_Jp / normf;
    v = norm_Dp / normf;
    *pred = u * u + 2.0 * w->mu * v * v;
    return ORC_SUCCESS;
}

/* ----------------------------------------------------------------------------------------- */
/* dogleg / double dogleg -- GSL multilarge_nlinear/dogleg.c                      */
/* ----------------------------------------------------------------------------------------- */

static int dogleg_preloop(orc_workspace *w)
{
    const size_t p = w->p;
    double u, alpha;
    for (size_t i = 0; i < p; ++i)
        w->workp1[i] = w->g[i] / w->diag[i];
    w->norm_Dinvg = v_nrm2(p, w->workp1);
    for (size_t i = 0; i < p; ++i)
        w->workp1[i] /= w->diag[i];
    symv_lower(p, w->JTJ, w->workp1, w->workp2);
    u = v_dot(p, w->workp1, w->workp2);
    w->norm_JDinv2g = sqrt(u);
    u = w->norm_Dinvg / w->norm_JDinv2g;
    alpha = u * u;
    for (size_t i = 0; i < p; ++i)
        w->dx_sd[i] = -alpha * w->workp1[i];
    w->norm_Dsd = scaled_enorm(p, w->diag, w->dx_sd);
    w->norm_Dgn = -1.0;
    return ORC_SUCCESS;
}

static int dogleg_calc_gn(orc_workspace *w, double *dx)
{
    int status;
    solver_init(w);
    status = solver_presolve(w, 0.0);
    if (status)
        return status;
    solver_solve(w, w->g, dx);
    return ORC_SUCCESS;
}

static double dogleg_beta(orc_workspace *w, double t, double delta)
{
    const size_t p = w->p;
    double a, b, c, beta;
    scaled_addition(p, t, w->dx_gn, -1.0, w->dx_sd, w->workp1);
    a = scaled_enorm(p, w->diag, w->workp1);
    a *= a;
    for (size_t i = 0; i < p; ++i)
        w->workp1[i] *= w->diag[i] * w->diag[i];
    b = 2.0 * v_dot(p, w->dx_sd, w->workp1);
    c = (w->norm_Dsd + delta) * (w->norm_Dsd - delta);
    if (b > 0.0)
        beta = (-2.0 * c) / (b + sqrt(b * b - 4.0 * a * c));
    else
        beta = (-b + sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
    return beta;
}

static int dogleg_step(orc_workspace *w, double delta, double *dx, int dbl)
{
    const size_t p = w->p;
    if (w->norm_Dsd >= delta) {
        for (size_t i = 0; i < p; ++i)
            dx[i] = w->dx_sd[i] * (delta / w->norm_Dsd);
        return ORC_SUCCESS;
    }
    if (w->norm_Dgn < 0.0) {
        int status = dogleg_calc_gn(w, w->dx_gn);
        if (status)
            return status;
        w->norm_Dgn = scaled_enorm(p, w->diag, w->dx_gn);
    }
    if (w->norm_Dgn <= delta) {
        memcpy(dx, w->dx_gn, p * sizeof(double));
        return ORC_SUCCESS;
    }
    if (!dbl) {
        const double beta = dogleg_beta(w, 1.0, delta);
        scaled_addition(p, 1.0, w->dx_gn, -1.0, w->dx_sd, w->workp1);
        scaled_addition(p, beta, w->workp1, 1.0, w->dx_sd, dx);
    } else {
        const double alpha_fac = 0.8;
        double t, u, v, c;
        v = w->norm_Dinvg / w->norm_JDinv2g;
        u = v * v;
        v = v_dot(p, w->g, w->dx_gn);
        c = u * (w->norm_Dinvg / fabs(v)) * w->norm_Dinvg;
        t = 1.0 - alpha_fac * (1.0 - c);
        if (t * w->norm_Dgn <= delta) {
            for (size_t i = 0; i < p; ++i)
                dx[i] = w->dx_gn[i] * (delta / w->norm_Dgn);
        } else {
            const double beta = dogleg_beta(w, t, delta);
            scaled_addition(p, t, w->dx_gn, -1.0, w->dx_sd, w->workp1);
            scaled_addition(p, beta, w->workp1, 1.0, w->dx_sd, dx);
        }
    }
    return ORC_SUCCESS;
}

/* ----------------------------------------------------------------------------------------- */
/* 2D subspace -- GSL multilarge_nlinear/subspace2D.c                            */
/* ----------------------------------------------------------------------------------------- */

/* gsl_linalg_householder_transform on a strided vector */
static double householder_transform(size_t n, double *v, size_t stride)
{
    if (n <= 1)
        return 0.0;
    double ssq = 0.0, scale = 0.0;
    /* xnorm = dnrm2(v[1:]) */
    {
        double sc = 0.0, sq = 1.0;
        for (size_t i = 1; i < n; ++i) {
            const double xi = v[i * stride];
            if (xi != 0.0) {
                const double ax = fabs(xi);
                if (sc < ax) { sq = 1.0 + sq * (sc / ax) * (sc / ax); sc = ax; }
                else sq += (ax / sc) * (ax / sc);
            }
        }
        scale = sc; ssq = sq;
    }
    const double xnorm = scale * sqrt(ssq);
    if (xnorm == 0.0)
        return 0.0;
    const double alpha = v[0];
    const double beta = -(alpha >= 0.0 ? +1.0 : -1.0) * hypot(alpha, xnorm);
    const double tau = (beta - alpha) / beta;
    const double s = alpha - beta;
    if (fabs(s) > DBL_MIN) {
        for (size_t i = 1; i < n; ++i)
            v[i * stride] *= 1.0 / s;
    } else {
        for (size_t i = 1; i < n; ++i)
            v[i * stride] *= DBL_EPSILON / s;
        for (size_t i = 1; i < n; ++i)
            v[i * stride] *= 1.0 / DBL_EPSILON;
    }
    v[0] = beta;
    return tau;
}

/* w <- (I - tau v v^T) w with v[0] == 1 implied */
static void householder_hv(size_t n, double tau, const double *v, size_t vstride, double *wv)
{
    if (tau == 0.0)
        return;
    double d = wv[0];
    for (size_t i = 1; i < n; ++i)
        d += v[i * vstride] * wv[i];
    wv[0] -= tau * d;
    for (size_t i = 1; i < n; ++i)
        wv[i] -= tau * d * v[i * vstride];
}

/* QRPT of the p-by-2 matrix W (row-major, 2 columns) */
static void qrpt_decomp_p2(size_t p, double *W, double tau[2], size_t perm[2])
{
    double norms[2];
    perm[0] = 0; perm[1] = 1;
    for (size_t j = 0; j < 2; ++j) {
        double s = 0.0;
        for (size_t i = 0; i < p; ++i)
            s += W[i * 2 + j] * W[i * 2 + j];
        norms[j] = sqrt(s);
    }
    const size_t K = MINV(p, (size_t)2);
    tau[0] = tau[1] = 0.0;
    for (size_t i = 0; i < K; ++i) {
        /* pivot */
        size_t kmax = i;
        double max_norm = norms[i];
        for (size_t j = i + 1; j < 2; ++j)
            if (norms[j] > max_norm) { max_norm = norms[j]; kmax = j; }
        if (kmax != i) {
            for (size_t r = 0; r < p; ++r) {
                const double t = W[r * 2 + i]; W[r * 2 + i] = W[r * 2 + kmax]; W[r * 2 + kmax] = t;
            }
            { size_t t = perm[i]; perm[i] = perm[kmax]; perm[kmax] = t; }
            { double t = norms[i]; norms[i] = norms[kmax]; norms[kmax] = t; }
        }
        tau[i] = householder_transform(p - i, &W[i * 2 + i], 2);
        if (i + 1 < 2) {
            /* apply to the remaining column */
            const size_t j = i + 1;
            if (tau[i] != 0.0) {
                double wj = W[i * 2 + j];
                for (size_t r = i + 1; r < p; ++r)
                    wj += W[r * 2 + j] * W[r * 2 + i];
                W[i * 2 + j] -= tau[i] * wj;
                for (size_t r = i + 1; r < p; ++r)
                    W[r * 2 + j] -= tau[i] * W[r * 2 + i] * wj;
            }
            /* norm downdate as in gsl_linalg_QRPT_decomp */
            if (i + 1 < p) {
                double x = norms[j];
                if (x > 0.0) {
                    double y = 0.0;
                    const double temp = W[i * 2 + j] / x;
                    if (fabs(temp) >= 1.0)
                        y = 0.0;
                    else
                        y = x * sqrt(1.0 - temp * temp);
                    if (fabs(y / x) < sqrt(20.0) * sqrt(DBL_EPSILON)) {
                        double s = 0.0;
                        for (size_t r = i + 1; r < p; ++r)
                            s += W[r * 2 + j] * W[r * 2 + j];
                        y = sqrt(s);
                    }
                    norms[j] = y;
                }
            }
        }
    }
}

static size_t qrpt_rank_p2(size_t p, const double *W)
{
    const size_t K = MINV(p, (size_t)2);
    double mn = W[0], mx = W[0];
    for (size_t i = 0; i < K; ++i) {
        const double d = W[i * 2 + i];
        mn = MINV(mn, d); mx = MAXV(mx, d);
    }
    const double absmax = MAXV(fabs(mn), fabs(mx));
    int ee;
    (void)frexp(absmax, &ee);
    const double eps = 20.0 * (double)(p + 2) * ldexp(1.0, ee) * DBL_EPSILON;
    size_t r = 0;
    for (size_t i = 0; i < K; ++i)
        if (fabs(W[i * 2 + i]) > eps)
            ++r;
    return r;
}

static void qr_QTvec_p2(size_t p, const double *W, const double tau[2], double *v)
{
    const size_t K = MINV(p, (size_t)2);
    for (size_t i = 0; i < K; ++i)
        householder_hv(p - i, tau[i], &W[i * 2 + i], 2, v + i);
}

static void qr_Qvec_p2(size_t p, const double *W, const double tau[2], double *v)
{
    const size_t K = MINV(p, (size_t)2);
    for (size_t i = K; i-- > 0;)
        householder_hv(p - i, tau[i], &W[i * 2 + i], 2, v + i);
}

static int subspace2D_preloop(orc_workspace *w)
{
    const size_t p = w->p;
    int status;
    double u, alpha;

    /* Gauss-Newton step */
    status = dogleg_calc_gn(w, w->dx_gn);
    if (status)
        return status;

    /* steepest descent step (same as dogleg_preloop) */
    for (size_t i = 0; i < p; ++i)
        w->workp1[i] = w->g[i] / w->diag[i];
    w->norm_Dinvg = v_nrm2(p, w->workp1);
    for (size_t i = 0; i < p; ++i)
        w->workp1[i] /= w->diag[i];
    symv_lower(p, w->JTJ, w->workp1, w->workp2);
    u = v_dot(p, w->workp1, w->workp2);
    w->norm_JDinv2g = sqrt(u);
    u = w->norm_Dinvg / w->norm_JDinv2g;
    alpha = u * u;
    for (size_t i = 0; i < p; ++i)
        w->dx_sd[i] = -alpha * w->workp1[i];

    w->norm_Dgn = scaled_enorm(p, w->diag, w->dx_gn);
    w->norm_Dsd = scaled_enorm(p, w->diag, w->dx_sd);

    for (size_t i = 0; i < p; ++i) {
        double a = w->dx_sd[i] * w->diag[i];
        double b = w->dx_gn[i] * w->diag[i];
        if (w->norm_Dsd != 0.0) a *= 1.0 / w->norm_Dsd;
        if (w->norm_Dgn != 0.0) b *= 1.0 / w->norm_Dgn;
        w->W[i * 2 + 0] = a;
        w->W[i * 2 + 1] = b;
    }
    qrpt_decomp_p2(p, w->W, w->tau, w->perm);
    w->rank = qrpt_rank_p2(p, w->W);

    if (w->rank == 2) {
        double B00, B10, B11, g0, g1;
        /* subg = Q^T D^{-1} g */
        for (size_t i = 0; i < p; ++i)
            w->workp1[i] = w->g[i] / w->diag[i];
        qr_QTvec_p2(p, w->W, w->tau, w->workp1);
        g0 = w->workp1[0];
        g1 = w->workp1[1];
        w->subg[0] = g0;
        w->subg[1] = g1;
        /* subB = Q^T D^{-1} J^T J D^{-1} Q, first two columns of Q */
        double *q0 = w->workp1, *q1 = w->workp2;
        for (size_t i = 0; i < p; ++i) { q0[i] = 0.0; q1[i] = 0.0; }
        q0[0] = 1.0;
        q1[1] = 1.0;
        qr_Qvec_p2(p, w->W, w->tau, q0);
        qr_Qvec_p2(p, w->W, w->tau, q1);
        for (size_t i = 0; i < p; ++i) { q0[i] /= w->diag[i]; q1[i] /= w->diag[i]; }
        symv_lower(p, w->JTJ, q0, w->workp);
        B00 = v_dot(p, q0, w->workp);
        B10 = v_dot(p, q1, w->workp);
        symv_lower(p, w->JTJ, q1, w->workp);
        B11 = v_dot(p, q1, w->workp);
        w->subB[0] = B00; w->subB[1] = B10; w->subB[2] = B10; w->subB[3] = B11;
        w->trB = B00 + B11;
        w->detB = B00 * B11 - B10 * B10;
        w->normg = v_nrm2(2, w->subg);
        w->term0 = (B10 * B10 + B11 * B11) * g0 * g0 - 2.0 * B10 * (B00 + B11) * g0 * g1 +
                   (B10 * B10 + B00 * B00) * g1 * g1;
        w->term1 = 2.0 * (B11 * g0 * g0 - 2.0 * B10 * g0 * g1 + B00 * g1 * g1);
    }
    return ORC_SUCCESS;
}

/* 2x2 gsl_linalg_mcholesky_decomp + mcholesky_solve: (B + lambda I + E) x = -g */
static void subspace2D_solution(const orc_workspace *w, double lambda, double x[2])
{
    double A00 = w->subB[0] + lambda, A10 = w->subB[1], A11 = w->subB[3] + lambda;
    const double gamma = MAXV(fabs(A00), fabs(A11));
    const double xi = fabs(A10);
    const double nu = sqrt(2.0 * 2.0 - 1.0);
    const double beta = sqrt(MAXV(MAXV(gamma, xi / nu), DBL_EPSILON));
    const double delta = DBL_EPSILON;
    int swapped = 0;
    /* pivot: largest |diagonal| first */
    if (fabs(A11) > fabs(A00)) {
        const double t = A00; A00 = A11; A11 = t;
        swapped = 1;
    }
    const double theta0 = fabs(A10);
    double u = theta0 / beta;
    const double d0 = MAXV(MAXV(delta, fabs(A00)), u * u);
    const double l10 = A10 / d0;
    const double a11 = A11 - A10 * A10 / d0;
    const double d1 = MAXV(delta, fabs(a11));
    /* solve L D L^T (P x) = P b */
    double b0 = swapped ? w->subg[1] : w->subg[0];
    double b1 = swapped ? w->subg[0] : w->subg[1];
    /* forward */
    double y0 = b0, y1 = b1 - l10 * y0;
    y0 /= d0;
    y1 /= d1;
    /* backward */
    double z1 = y1, z0 = y0 - l10 * z1;
    if (swapped) { x[0] = -z1; x[1] = -z0; }
    else { x[0] = -z0; x[1] = -z1; }
}

static double subspace2D_objective(const orc_workspace *w, const double x[2])
{
    const double y0 = w->subg[0] + 0.5 * (w->subB[0] * x[0] + w->subB[1] * x[1]);
    const double y1 = w->subg[1] + 0.5 * (w->subB[2] * x[0] + w->subB[3] * x[1]);
    return x[0] * y0 + x[1] * y1;
}

/* all four complex roots of a[0] + a[1] z + ... + a[4] z^4 (Aberth-Ehrlich, long double) */
static int quartic_roots(const double a[5], double zr[4], double zi[4])
{
    long double cr[4], ci[4];
    long double c[5];
    for (int i = 0; i < 5; ++i)
        c[i] = (long double)a[i] / (long double)a[4];
    long double R = 0.0L;
    for (int i = 0; i < 4; ++i)
        R = MAXV(R, fabsl(c[i]));
    R = 1.0L + R;
    /* Cauchy-type tighter radius: max |c_i|^(1/(4-i)) * 2 */
    long double R2 = 0.0L;
    for (int i = 0; i < 4; ++i) {
        const long double t = powl(fabsl(c[i]), 1.0L / (long double)(4 - i));
        R2 = MAXV(R2, t);
    }
    R2 *= 2.0L;
    if (R2 > 0.0L && R2 < R)
        R = R2;
    for (int k = 0; k < 4; ++k) {
        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / 4.0L + 0.4L;
        cr[k] = 0.5L * R * cosl(ang);
        ci[k] = 0.5L * R * sinl(ang);
    }
    for (int it = 0; it < 500; ++it) {
        long double maxrel = 0.0L;
        for (int k = 0; k < 4; ++k) {
            /* Horner for p and p' */
            long double pr = 1.0L, pi = 0.0L, dr = 0.0L, di = 0.0L;
            for (int i = 3; i >= 0; --i) {
                /* d = d*z + p ; p = p*z + c[i] */
                long double ndr = dr * cr[k] - di * ci[k] + pr;
                long double ndi = dr * ci[k] + di * cr[k] + pi;
                long double npr = pr * cr[k] - pi * ci[k] + c[i];
                long double npi = pr * ci[k] + pi * cr[k];
                dr = ndr; di = ndi; pr = npr; pi = npi;
            }
            const long double dden = dr * dr + di * di;
            if (dden == 0.0L)
                continue;
            /* w = p / p' */
            long double wr = (pr * dr + pi * di) / dden;
            long double wi = (pi * dr - pr * di) / dden;
            /* s = sum 1/(z_k - z_j) */
            long double sr = 0.0L, si = 0.0L;
            for (int j = 0; j < 4; ++j) {
                if (j == k)
                    continue;
                const long double er = cr[k] - cr[j], ei = ci[k] - ci[j];
                const long double den = er * er + ei * ei;
                if (den == 0.0L)
                    continue;
                sr += er / den;
                si += -ei / den;
            }
            /* corr = w / (1 - w s) */
            const long double qr = 1.0L - (wr * sr - wi * si);
            const long double qi = -(wr * si + wi * sr);
            const long double qden = qr * qr + qi * qi;
            long double corr_r, corr_i;
            if (qden == 0.0L) { corr_r = wr; corr_i = wi; }
            else {
                corr_r = (wr * qr + wi * qi) / qden;
                corr_i = (wi * qr - wr * qi) / qden;
            }
            cr[k] -= corr_r;
            ci[k] -= corr_i;
            const long double mag = sqrtl(cr[k] * cr[k] + ci[k] * ci[k]);
            const long double cm = sqrtl(corr_r * corr_r + corr_i * corr_i);
            const long double rel = cm / MAXV(mag, (long double)DBL_MIN);
            maxrel = MAXV(maxrel, rel);
        }
        if (maxrel < 1.0e-18L)
            break;
    }
    for (int k = 0; k < 4; ++k) {
        zr[k] = (double)cr[k];
        zi[k] = (double)ci[k];
    }
    return ORC_SUCCESS;
}

static int subspace2D_step(orc_workspace *w, double delta, double *dx)
{
    const size_t p = w->p;
    if (w->norm_Dgn <= delta) {
        memcpy(dx, w->dx_gn, p * sizeof(double));
    } else if (w->rank < 2) {
        for (size_t i = 0; i < p; ++i)
            dx[i] = w->dx_sd[i] * (delta / w->norm_Dsd);
    } else {
        const double delta_sq = delta * delta;
        const double u = w->normg / delta;
        double a[5], zr[4], zi[4];
        double minc = 0.0;
        int mini = -1;
        double x[2];
        a[0] = w->detB * w->detB - w->term0 / delta_sq;
        a[1] = 2.0 * w->detB * w->trB - w->term1 / delta_sq;
        a[2] = w->trB * w->trB + 2.0 * w->detB - u * u;
        a[3] = 2.0 * w->trB;
        a[4] = 1.0;
        quartic_roots(a, zr, zi);
        /* GSL uses the REAL PART of every root as a candidate multiplier */
        for (int i = 0; i < 4; ++i) {
            double cost, normx;
            subspace2D_solution(w, zr[i], x);
            normx = v_nrm2(2, x);
            if (normx == 0.0)
                continue;
            x[0] *= delta / normx;
            x[1] *= delta / normx;
            cost = subspace2D_objective(w, x);
            if (mini < 0 || cost < minc) {
                mini = i;
                minc = cost;
            }
        }
        if (mini < 0)
            return ORC_FAILURE;
        subspace2D_solution(w, zr[mini], x);
        for (size_t i = 0; i < p; ++i)
            dx[i] = 0.0;
        dx[0] = x[0];
        dx[1] = x[1];
        qr_Qvec_p2(p, w->W, w->tau, dx);
        for (size_t i = 0; i < p; ++i)
            dx[i] /= w->diag[i];
    }
    return ORC_SUCCESS;
}

/* ----------------------------------------------------------------------------------------- */
/* Steihaug-Toint CG -- GSL multilarge_nlinear/cgst.c                            */
/* ----------------------------------------------------------------------------------------- */

static double cgst_calc_tau(size_t p, const double *pv, const double *q, double delta)
{
    const double norm_p = v_nrm2(p, pv);
    const double norm_q = v_nrm2(p, q);
    const double u = v_dot(p, pv, q);
    const double t1 = u / (norm_q * norm_q);
    const double t2 = t1 * u + (delta + norm_p) * (delta - norm_p);
    return -t1 + sqrt(t2) / norm_q;
}

static int cgst_step(orc_workspace *w, double delta, double *dx)
{
    const size_t p = w->p;
    const size_t cgmaxit = w->params.cg_max_iter ? w->params.cg_max_iter : w->n;
    const double cgtol = w->params.cg_tol;
    double alpha, beta, u, norm_Jd, norm_r, norm_rp1;
    int status;

    for (size_t i = 0; i < p; ++i) {
        const double gi = w->g[i], di = w->diag[i];
        w->cg_z[i] = 0.0;
        w->cg_r[i] = -gi / di;
        w->cg_d[i] = -gi / di;
        w->workp[i] = gi / di;
    }
    w->cg_norm_g = v_nrm2(p, w->workp);

    for (size_t it = 0; it < cgmaxit; ++it) {
        for (size_t i = 0; i < p; ++i)
            w->workp[i] = w->cg_d[i] / w->diag[i];
        /* workn = J D^{-1} d : a df callback (NoTrans) in the reference */
        status = eval_df(w, ORC_NOTRANS, w->x, w->workp, w->workn, NULL);
        if (status)
            return status;
        norm_Jd = v_nrm2(w->n, w->workn);
        if (norm_Jd == 0.0) {
            const double tau = cgst_calc_tau(p, w->cg_z, w->cg_d, delta);
            scaled_addition(p, 1.0, w->cg_z, tau, w->cg_d, dx);
            for (size_t i = 0; i < p; ++i)
                dx[i] /= w->diag[i];
            return ORC_SUCCESS;
        }
        norm_r = v_nrm2(p, w->cg_r);
        u = norm_r / norm_Jd;
        alpha = u * u;
        scaled_addition(p, 1.0, w->cg_z, alpha, w->cg_d, w->workp);
        u = v_nrm2(p, w->workp);
        if (u >= delta) {
            const double tau = cgst_calc_tau(p, w->cg_z, w->cg_d, delta);
            scaled_addition(p, 1.0, w->cg_z, tau, w->cg_d, dx);
            for (size_t i = 0; i < p; ++i)
                dx[i] /= w->diag[i];
            return ORC_SUCCESS;
        }
        memcpy(w->cg_z, w->workp, p * sizeof(double));
        /* workp = J^T (J D^{-1} d): second df callback (Trans) */
        status = eval_df(w, ORC_TRANS, w->x, w->workn, w->workp, NULL);
        if (status)
            return status;
        for (size_t i = 0; i < p; ++i)
            w->workp[i] = (w->workp[i] / w->diag[i]) * alpha;
        for (size_t i = 0; i < p; ++i)
            w->cg_r[i] -= w->workp[i];
        norm_rp1 = v_nrm2(p, w->cg_r);
        u = norm_rp1 / w->cg_norm_g;
        if (u < cgtol) {
            for (size_t i = 0; i < p; ++i)
                dx[i] = w->cg_z[i] / w->diag[i];
            return ORC_SUCCESS;
        }
        u = norm_rp1 / norm_r;
        beta = u * u;
        scaled_addition(p, 1.0, w->cg_r, beta, w->cg_d, w->cg_d);
    }
    for (size_t i = 0; i < p; ++i)
        dx[i] = w->cg_z[i] / w->diag[i];
    return ORC_EMAXITER;
}

/* ----------------------------------------------------------------------------------------- */
/* trs dispatch                                                                                */
/* ----------------------------------------------------------------------------------------- */

static int trs_preloop(orc_workspace *w)
{
    switch (w->params.trs) {
    case ORC_TRS_LM:
    case ORC_TRS_LMACCEL: return lm_preloop(w);
    case ORC_TRS_DOGLEG:
    case ORC_TRS_DDOGLEG: return dogleg_preloop(w);
    case ORC_TRS_SUBSPACE2D: return subspace2D_preloop(w);
    default: return ORC_SUCCESS;
    }
}

static int trs_step(orc_workspace *w, double delta, double *dx)
{
    switch (w->params.trs) {
    case ORC_TRS_LM:
    case ORC_TRS_LMACCEL: return lm_step(w, delta, dx);
    case ORC_TRS_DOGLEG: return dogleg_step(w, delta, dx, 0);
    case ORC_TRS_DDOGLEG: return dogleg_step(w, delta, dx, 1);
    case ORC_TRS_SUBSPACE2D: return subspace2D_step(w, delta, dx);
    default: return cgst_step(w, delta, dx);
    }
}

static int trs_preduction(orc_workspace *w, const double *dx, double *pred)
{
    switch (w->params.trs) {
    case ORC_TRS_LM:
    case ORC_TRS_LMACCEL: return lm_preduction(w, dx, pred);
    default: *pred = quadratic_preduction(w, dx); return ORC_SUCCESS;
    }
}

/* ----------------------------------------------------------------------------------------- */
/* trust region driver -- GSL multilarge_nlinear/trust.c == src/trust.c:311-372, 408-549       */
/* ----------------------------------------------------------------------------------------- */

int orc_winit(const double *x0, const double *wts, orc_fdf *fdf, orc_workspace *w)
{
    const size_t n = w->n, p = w->p;
    int status;
    double Dx;

    if (fdf->n != n || fdf->p != p)
        return ORC_EINVAL;
    w->fdf = fdf;
    fdf->nevalf = fdf->nevaldfu = fdf->nevaldf2 = fdf->nevalfvv = 0;
    memcpy(w->x, x0, p * sizeof(double));
    w->niter = 0;
    if (wts) {
        /* src/fdf.c:60-64: sqrt_wts_i = sqrt(w_i) */
        for (size_t i = 0; i < n; ++i)
            w->sqrt_wts_buf[i] = sqrt(wts[i]);
        w->sqrt_wts = w->sqrt_wts_buf;
    } else {
        w->sqrt_wts = NULL;
    }

    /* trust_init */
    status = eval_f(w, w->x, w->f);
    if (status)
        return status;
    status = eval_df(w, ORC_TRANS, w->x, w->f, w->g, w->JTJ);
    if (status)
        return status;
    scale_init(w, w->JTJ, w->diag);
    Dx = scaled_enorm(p, w->diag, w->x);
    w->delta = 0.3 * MAXV(1.0, Dx);
    nielsen_init(w, &w->mu, &w->nu);
    w->avratio = 0.0;
    for (size_t i = 0; i < p; ++i)
        w->acc[i] = 0.0;
    return ORC_SUCCESS;
}

static double trust_calc_rho(orc_workspace *w, const double *f_trial, const double *dx)
{
    const double normf = v_nrm2(w->n, w->f);
    const double normf_trial = v_nrm2(w->n, f_trial);
    double rho, actual_reduction, pred_reduction, u;
    int status;

    if (normf_trial >= normf)
        return -1.0;
    u = normf_trial / normf;
    actual_reduction = 1.0 - u * u;
    status = trs_preduction(w, dx, &pred_reduction);
    if (status)
        return -1.0;
    if (pred_reduction > 0.0)
        rho = actual_reduction / pred_reduction;
    else
        rho = -1.0;
    return rho;
}

static int trust_eval_step(orc_workspace *w, const double *f_trial, const double *dx, double *rho)
{
    int status = ORC_SUCCESS;
    if (w->params.trs == ORC_TRS_LMACCEL) {
        if (w->avratio > w->params.avmax)
            status = ORC_FAILURE;
    }
    *rho = trust_calc_rho(w, f_trial, dx);
    if (*rho <= 0.0)
        status = ORC_FAILURE;
    return status;
}

int orc_iterate(orc_workspace *w)
{
    const size_t p = w->p;
    int status;
    double rho;
    int foundstep = 0, bad_steps = 0;

    status = trs_preloop(w);
    if (status) {
        ++w->niter; /* gsl_multilarge_nlinear_iterate increments after the call regardless */
        return status;
    }

    while (!foundstep) {
        status = trs_step(w, w->delta, w->dx);
        if (status == ORC_SUCCESS) {
            for (size_t i = 0; i < p; ++i)
                w->x_trial[i] = w->x[i] + w->dx[i];
            status = eval_f(w, w->x_trial, w->f_trial);
            if (status) {
                ++w->niter;
                return status;
            }
            status = trust_eval_step(w, w->f_trial, w->dx, &rho);
            if (status == ORC_SUCCESS)
                foundstep = 1;
        } else {
            rho = -1.0;
        }

        if (rho > 0.75)
            w->delta *= w->params.factor_up;
        else if (rho < 0.25)
            w->delta /= w->params.factor_down;

        if (foundstep) {
            memcpy(w->x, w->x_trial, p * sizeof(double));
            memcpy(w->f, w->f_trial, w->n * sizeof(double));
            status = eval_df(w, ORC_TRANS, w->x, w->f, w->g, w->JTJ);
            if (status) {
                ++w->niter;
                return status;
            }
            scale_update(w, w->JTJ, w->diag);
            nielsen_accept(rho, &w->mu, &w->nu);
            bad_steps = 0;
        } else {
            nielsen_reject(&w->mu, &w->nu);
            if (++bad_steps > 15) {
                ++w->niter;
                return ORC_ENOPROG;
            }
        }
    }
    ++w->niter;
    return ORC_SUCCESS;
}

/* GSL multilarge_nlinear/convergence.c */
int orc_test(double xtol, double gtol, double ftol, int *info, const orc_workspace *w)
{
    const size_t p = w->p;
    double gnorm = 0.0, fnorm, phi;
    int ok = 1;
    (void)ftol;
    *info = 0;
    {
        const double epsabs = xtol * xtol, epsrel = xtol;
        for (size_t i = 0; i < p; ++i) {
            const double tolerance = epsabs + epsrel * fabs(w->x[i]);
            if (fabs(w->dx[i]) < tolerance)
                ok = 1;
            else {
                ok = 0;
                break;
            }
        }
        if (ok) {
            *info = 1;
            return ORC_SUCCESS;
        }
    }
    for (size_t i = 0; i < p; ++i) {
        const double xi = MAXV(w->x[i], 1.0);
        const double tmp = fabs(xi * w->g[i]);
        if (tmp > gnorm)
            gnorm = tmp;
    }
    fnorm = v_nrm2(w->n, w->f);
    phi = 0.5 * fnorm * fnorm;
    if (gnorm <= gtol * MAXV(phi, 1.0)) {
        *info = 2;
        return ORC_SUCCESS;
    }
    return ORC_CONTINUE;
}

int orc_covar(double *covar, orc_workspace *w)
{
    const size_t p = w->p;
    int status;
    for (size_t i = 0; i < p; ++i)
        for (size_t j = 0; j < p; ++j)
            covar[i * p + j] = (j <= i) ? w->JTJ[i * p + j] : 0.0;
    status = cholesky_decomp1(p, covar);
    if (status)
        return status;
    cholesky_invert(p, covar);
    return ORC_SUCCESS;
}

/* gsl_multilarge_nlinear_rcond (call site src/nls_large.c:735) -> the Cholesky solver's rcond (GSL
 * multilarge_nlinear/cholesky.c): re-factor J^T J, gsl_linalg_cholesky_rcond, square root.  Third-party,
 * restated from GSL 2.x linalg/cholesky.c + linalg/condest.c:
 *   ||A||_1      cholesky_norm1: diagonal rebuilt from the factor (dot of row j of L with itself), off-diagonal
 *                moduli from the original matrix that decomp1 keeps in the upper triangle;
 *   ||A^-1||_1   gsl_linalg_invnorm1: Hager's estimator with Higham's refinements (Algorithm 4.1 of
 *                "FORTRAN codes for estimating the one-norm of a real or complex matrix", ACM TOMS 14, 1988):
 *                start x = 1/N, at most five sign-vector sweeps, then the alternating-sign safeguard vector. */
static int sign_of(double v) { return v >= 0.0 ? 1 : -1; }

int orc_rcond(double *rcond, orc_workspace *w)
{
    const size_t N = w->p;
    double *L = (double *)malloc(N * N * sizeof(double));
    double *work = (double *)calloc(4 * N, sizeof(double));
    double *x = work, *v = work + N, *xi = work + 2 * N, *col = work + 3 * N;
    double Anorm = 0.0, gamma, gamma_old, temp;
    int status;
    *rcond = 0.0;
    for (size_t i = 0; i < N; ++i)
        for (size_t j = 0; j < N; ++j)
            L[i * N + j] = (j <= i) ? w->JTJ[i * N + j] : 0.0;
    status = cholesky_decomp1(N, L);
    if (status) {
        free(L);
        free(work);
        return status;
    }
    /* cholesky_norm1 */
    for (size_t j = 0; j < N; ++j) {
        double sum = 0.0, Ajj = 0.0;
        for (size_t k = 0; k <= j; ++k)
            Ajj += L[j * N + k] * L[j * N + k];
        for (size_t i = 0; i < j; ++i) {
            const double absAij = fabs(w->JTJ[j * N + i]);
            sum += absAij;
            col[i] += absAij;
        }
        col[j] = sum + fabs(Ajj);
    }
    for (size_t i = 0; i < N; ++i)
        Anorm = MAXV(Anorm, col[i]);
    if (Anorm == 0.0) {
        free(L);
        free(work);
        return ORC_SUCCESS;
    }
    /* gsl_linalg_invnorm1 with Ainvx = two triangular solves on L */
    for (size_t i = 0; i < N; ++i)
        x[i] = 1.0 / (double)N;
    cholesky_solve(N, L, x, v);
    gamma = 0.0;
    for (size_t i = 0; i < N; ++i) {
        gamma += fabs(v[i]);
        xi[i] = (double)sign_of(v[i]);
    }
    cholesky_solve(N, L, xi, x);
    for (size_t k = 0; k < 5; ++k) {
        size_t jmax = 0;
        double amax = 0.0;
        int same = 1;
        for (size_t i = 0; i < N; ++i) /* idamax: first index of the largest modulus */
            if (fabs(x[i]) > amax) {
                amax = fabs(x[i]);
                jmax = i;
            }
        for (size_t i = 0; i < N; ++i)
            v[i] = (i == jmax) ? 1.0 : 0.0;
        cholesky_solve(N, L, v, v);
        gamma_old = gamma;
        gamma = 0.0;
        for (size_t i = 0; i < N; ++i) {
            gamma += fabs(v[i]);
            if ((double)sign_of(v[i]) != xi[i])
                same = 0;
        }
        if (same || gamma < gamma_old)
            break;
        for (size_t i = 0; i < N; ++i)
            xi[i] = (double)sign_of(v[i]);
        cholesky_solve(N, L, xi, x);
    }
    temp = 1.0;
    for (size_t i = 0; i < N; ++i) {
        x[i] = temp * (1.0 + (double)i / ((double)N - 1.0));
        temp = -temp;
    }
    cholesky_solve(N, L, x, x);
    temp = 0.0;
    for (size_t i = 0; i < N; ++i)
        temp += fabs(x[i]);
    temp = 2.0 * temp / (3.0 * (double)N);
    if (temp > gamma)
        gamma = temp;
    if (gamma != 0.0)
        *rcond = sqrt((1.0 / Anorm) / gamma);
    free(L);
    free(work);
    return ORC_SUCCESS;
}

/* src/nls_fit.c:153-224, statement for statement */
int orc_driver2(size_t maxiter, double xtol, double gtol, double ftol, orc_callback cb, void *cbparams,
                int *info, double *chisq0, double *chisq1, orc_workspace *w)
{
    int status = ORC_CONTINUE;
    size_t iter = 0;
    do {
        chisq0[0] = chisq1[0];
        status = orc_iterate(w);
        chisq1[0] = v_dot(w->n, w->f, w->f);
        if (status == ORC_EBADFUNC || (status == ORC_ENOPROG && iter == 0)) {
            *info = status;
            return status;
        }
        ++iter;
        if (cb)
            cb(iter, cbparams, w);
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
