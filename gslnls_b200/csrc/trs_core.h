// trs_core.h -- the p-sized half of one gsl_nls_large() iteration, as a packet-driven state machine.
//
// The fused pass kernel (nls_pass_kernel.cuh) turns the n observations into the normal-equation
// packet [J^T J | J^T f | f^T f] at a requested parameter vector.  Everything the reference does
// between two such evaluations lives here and runs on the device in one warp:
//
//   trust_init             GSL multilarge trust.c;     in-tree spec src/trust.c:311-372
//   trust_iterate          accept/reject loop          src/trust.c:408-549
//   nielsen_*              mu / nu updates             src/trust.c:149-199
//   scaling more/levenberg/marquardt (on diag J^T J)   SURVEY A.2
//   lm / lmaccel step + More' predicted reduction      src/trust.c:223-292, SURVEY A.4
//   dogleg / double dogleg / 2D subspace / Steihaug CG SURVEY A.5-A.7
//   driver2 + convergence test                         src/nls_fit.c:153-224, SURVEY A.9
//   covariance (J^T J)^-1, cond(J) for the trace       src/nls_large.c:251-256, 715-739
//
// Execution model ("SPMD-redundant, cooperative Cholesky"): every lane of the warp runs the same
// control flow on private copies of all p-vectors and scalars; the two p x p matrices live in
// memory shared by the lanes and only the O(p^3) factorisation is split across lanes.  With the
// SingleLane policy the same code is one thread per problem (batched multi-start) and also
// compiles for the host, where tests/ drives it against the oracle without a GPU.  The host
// build is test infrastructure; libgslnls_b200.so contains the device instantiations only.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TRS_HD __host__ __device__
#else
#define TRS_HD
#endif

namespace trs {

enum { PH_INIT = 0, PH_TRIAL = 1, PH_ACCEL = 2, PH_DONE = 3 };
enum { MODE_IDLE = 0, MODE_FJ = 1, MODE_FVV = 2, MODE_JVP = 3 };
enum { TRS_LM = 0, TRS_LMACCEL, TRS_DOGLEG, TRS_DDOGLEG, TRS_SUBSPACE2D, TRS_CGST };
enum { SCALE_MORE = 0, SCALE_LEVENBERG, SCALE_MARQUARDT };
enum { E_SUCCESS = 0, E_FAILURE = -1, E_CONTINUE = -2, E_EDOM = 1, E_EBADFUNC = 9, E_EMAXITER = 11, E_ENOPROG = 27 };

// scalar slots at the head of a state record (all stored as doubles; integers are exact)
enum {
    S_PHASE = 0, S_STATUS, S_INFO, S_NITER, S_ITER, S_BAD, S_NU, S_NEVAL_F, S_NEVAL_DFU, S_NEVAL_DF2,
    S_NEVAL_FVV, S_MU, S_DELTA, S_AVRATIO, S_CHISQ0, S_CHISQ1, S_F2, S_CHISQ_INIT, S_NPASS, S_RHO,
    S_LOGDET0, S_NBAD, S_LOGDET1, S_COUNT = 24
};

struct Params {
    int p, maxiter, trs, scale, trace, batch_iters; // batch_iters > 0: stop after that many iterations
    long long cg_maxit;
    double factor_up, factor_down, avmax, h_df, h_fvv, xtol, ftol, gtol, cg_tol;
};

// record layout (doubles): [S_COUNT scalars | x dx g diag xtrial vel (6p) | JTJ p*p | covar p*p]
TRS_HD inline int state_doubles(int p) { return S_COUNT + 6 * p + 2 * p * p; }
TRS_HD inline int packet_doubles(int p) { return p * (p + 1) / 2 + p + 1; }
TRS_HD inline int request_doubles(int p) { return 2 * p + 2; }

struct SingleLane {
    TRS_HD int lane() const { return 0; }
    TRS_HD int nlanes() const { return 1; }
    TRS_HD void sync() const {}
    TRS_HD double bcast(double v, int) const { return v; }
    TRS_HD bool all(bool b) const { return b; }
};

#if defined(__CUDACC__)
struct WarpLanes {
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ int nlanes() const { return 32; }
    __device__ void sync() const { __syncwarp(); }
    __device__ double bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ bool all(bool b) const { return __all_sync(0xffffffffu, b) != 0; }
};
#endif

TRS_HD inline bool finite_d(double v) { return (v - v) == 0.0; }

// ||f|| as the reference computes it: gslcblas' dnrm2 (scaled sum of squares) returns +Inf for a
// vector with exactly one Inf entry but NaN as soon as there are two (Inf/Inf in the update), and
// every comparison in trust_calc_rho (src/trust.c:84-115) is then false -- the step is *accepted*.
// Reproduced on purpose: results must match the reference on the same inputs.
TRS_HD inline double norm_of(double sumsq, double nbad) { return nbad >= 2.0 ? NAN : sqrt(sumsq); }

// number of parameters: a run-time value bounded by PMAX, or -- FIXED -- the compile-time constant
// PMAX itself, which lets the compiler unroll the p-loops and keep the private vectors in registers
// (the resident server instantiates the small sizes that way)
template <int PMAX, bool FIXED>
struct Dim {
    int v;
    TRS_HD explicit Dim(int x) : v(x) {}
    TRS_HD operator int() const { return v; }
};
template <int PMAX>
struct Dim<PMAX, true> {
    TRS_HD explicit Dim(int) {}
    TRS_HD constexpr operator int() const { return PMAX; }
};

template <int PMAX, class Lanes, bool FIXED = false>
struct Solver {
    const Params &P;
    Lanes L;
    const Dim<PMAX, FIXED> p;
    double *JTJ; // p*p row-major, lower triangle valid; shared by the lanes
    double *A;   // p*p work / factor; shared by the lanes
    // private vectors
    double x[PMAX], dx[PMAX], g[PMAX], diag[PMAX], xt[PMAX], vel[PMAX], acc[PMAX];
    double w1[PMAX], w2[PMAX], w3[PMAX], gn[PMAX], sd[PMAX], W0[PMAX], W1[PMAX];
    // scalars
    int phase, status, info, niter, iter, bad;
    double nu;
    double nf, ndfu, ndf2, nfvv, npass;
    double mu, delta, avratio, chisq0, chisq1, f2, chisq_init, rho, logdet0, logdet1;
    double nbad; // number of non-finite residuals in the current f
    // per-iteration subproblem products (recomputed from g, JTJ, diag on every call)
    double norm_Dgn, norm_Dsd, norm_Dinvg, norm_JDinv2g;
    double trB, detB, normg, term0, term1, tau[2], subg[2], B00, B10, B11;
    int rank, gn_status;
    bool factor_valid; // A currently holds chol(JTJ + mu_f D^2)
    bool jtj_dirty;    // JTJ changed since load(): must be written back
    // resident use (trs_server): the solver object outlives the step, so the state record is read from
    // memory once per fit and written back only when the fit ends (finish) or the server is told to leave
    // (flush); a step then costs no global-memory round trips beyond the packet and the request
    bool keep_state = false, state_loaded = false;

    TRS_HD Solver(const Params &prm, Lanes lanes, double *jtj, double *work)
        : P(prm), L(lanes), p(prm.p), JTJ(jtj), A(work) {}

    // ---------------------------------------------------------------- small private helpers
    TRS_HD double dot(const double *a, const double *b) const
    {
        double r = 0.0;
        for (int i = 0; i < p; ++i)
            r += a[i] * b[i];
        return r;
    }
    TRS_HD double nrm2(const double *a, int n) const
    {
        double scale = 0.0, ssq = 1.0;
        if (n == 1)
            return fabs(a[0]);
        for (int i = 0; i < n; ++i) {
            if (a[i] != 0.0) {
                const double ax = fabs(a[i]);
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
    TRS_HD double scaled_enorm(const double *d, const double *a) const
    {
        double e2 = 0.0;
        for (int i = 0; i < p; ++i) {
            const double u = d[i] * a[i];
            e2 += u * u;
        }
        return sqrt(e2);
    }
    // p above which the lanes share the O(p^2) loops of symv() and solve_neg() instead of each running
    // them privately; below it the private loops (vectors in registers) are faster
    static constexpr int kCoopMinP = 9;
    static constexpr int kSlots = (PMAX + 31) / 32;

    // y = JTJ v using the stored lower triangle.  Small p: every lane computes all of y privately.
    // Large p with a warp: lane l computes the elements j = l, l + 32, ... and the results are exchanged
    // by shuffles.  Each element sees exactly the operations of the private loop in the same order
    // (row part in k order, then the diagonal term, then the column part in i order), so both forms --
    // and the one-lane host build the tests compare against -- give bitwise the same y.
    TRS_HD void symv(const double *v, double *y) const
    {
        const int nl = L.nlanes();
        if (nl > 1 && p >= kCoopMinP) {
            const int ln = L.lane();
            for (int q = 0; q * nl < p; ++q) {
                const int j = q * nl + ln;
                double yj = 0.0;
                if (j < p) {
                    double t2 = 0.0;
                    for (int k = 0; k < j; ++k)
                        t2 += JTJ[j * p + k] * v[k];
                    yj += v[j] * JTJ[j * p + j] + t2;
                    for (int i = j + 1; i < p; ++i)
                        yj += v[i] * JTJ[i * p + j];
                }
                for (int l = 0; l < nl; ++l) {
                    const double b = L.bcast(yj, l);
                    if (q * nl + l < p)
                        y[q * nl + l] = b;
                }
            }
            return;
        }
        for (int i = 0; i < p; ++i)
            y[i] = 0.0;
        for (int i = 0; i < p; ++i) {
            const double t1 = v[i];
            double t2 = 0.0;
            for (int j = 0; j < i; ++j) {
                const double a = JTJ[i * p + j];
                y[j] += t1 * a;
                t2 += a * v[j];
            }
            y[i] += t1 * JTJ[i * p + i] + t2;
        }
    }

    // ---------------------------------------------------------------- cooperative factorisation
    // A <- chol(JTJ + mu diag^2), lower, in place (left-looking, reciprocal scaling: the loop order
    // of gsl_linalg_cholesky_decomp1).  Returns E_EDOM when a pivot is not positive.
    TRS_HD int factor(double muval)
    {
        const int ln = L.lane(), nl = L.nlanes();
        L.sync();
        for (int i = ln; i < p; i += nl)
            for (int j = 0; j <= i; ++j)
                A[i * p + j] = JTJ[i * p + j] + (i == j ? muval * diag[i] * diag[i] : 0.0);
        L.sync();
        int st = E_SUCCESS;
        for (int j = 0; j < p; ++j) {
            for (int i = j + ln; i < p; i += nl) {
                double s = A[i * p + j];
                for (int k = 0; k < j; ++k)
                    s -= A[i * p + k] * A[j * p + k];
                A[i * p + j] = s;
            }
            L.sync();
            const double ajj = A[j * p + j];
            if (!(ajj > 0.0)) {
                st = E_EDOM;
                break; // uniform across lanes: every lane reads the same pivot
            }
            const double inv = 1.0 / sqrt(ajj);
            L.sync();
            for (int i = j + ln; i < p; i += nl)
                A[i * p + j] *= inv;
            L.sync();
        }
        factor_valid = (st == E_SUCCESS);
        return st;
    }
    // out = -(L L^T)^-1 b on the shared factor.  Small p: private substitution.  Large p with a warp: the
    // substitutions run column by column, lane l owning the elements l, l + 32, ...; the finished
    // component is broadcast by a shuffle and every lane updates its own elements.  Per element the
    // subtractions happen in the order of the private loops (j ascending forward, i descending
    // backward), so the result is bitwise the same.
    TRS_HD void solve_neg(const double *b, double *out) const
    {
        const int nl = L.nlanes();
        if (nl > 1 && p >= kCoopMinP) {
            const int ln = L.lane();
            double t[kSlots];
            for (int q = 0; q < kSlots; ++q)
                t[q] = (q * nl + ln < p) ? b[q * nl + ln] : 0.0;
            for (int j = 0; j < p; ++j) { // forward: L z = b
                const int oq = j / nl;
                double own = 0.0;
                for (int q = 0; q < kSlots; ++q)
                    if (q == oq)
                        own = t[q];
                const double zj = L.bcast(own / A[j * p + j], j % nl);
                out[j] = zj;
                for (int q = 0; q < kSlots; ++q) {
                    const int i = q * nl + ln;
                    if (i > j && i < p)
                        t[q] -= A[i * p + j] * zj;
                }
            }
            for (int q = 0; q < kSlots; ++q)
                t[q] = (q * nl + ln < p) ? out[q * nl + ln] : 0.0;
            for (int i = p - 1; i >= 0; --i) { // backward: L^T x = z
                const int oq = i / nl;
                double own = 0.0;
                for (int q = 0; q < kSlots; ++q)
                    if (q == oq)
                        own = t[q];
                const double xi = L.bcast(own / A[i * p + i], i % nl);
                out[i] = xi;
                for (int q = 0; q < kSlots; ++q) {
                    const int j = q * nl + ln;
                    if (j < i)
                        t[q] -= A[i * p + j] * xi;
                }
            }
            for (int i = 0; i < p; ++i)
                out[i] = -out[i];
            return;
        }
        solve_neg_private(b, out);
    }
    // the private substitution: no lane talks to another, so lanes may run it on different right-hand sides
    TRS_HD void solve_neg_private(const double *b, double *out) const
    {
        for (int i = 0; i < p; ++i)
            out[i] = b[i];
        for (int i = 0; i < p; ++i) {
            double t = out[i];
            for (int j = 0; j < i; ++j)
                t -= A[i * p + j] * out[j];
            out[i] = t / A[i * p + i];
        }
        for (int i = p - 1; i >= 0; --i) {
            out[i] = out[i] / A[i * p + i];
            const double xi = out[i];
            for (int j = 0; j < i; ++j)
                out[j] -= A[i * p + j] * xi;
        }
        for (int i = 0; i < p; ++i)
            out[i] = -out[i];
    }

    // ---------------------------------------------------------------- scaling + Nielsen
    TRS_HD void scale_apply(bool init)
    {
        for (int j = 0; j < p; ++j) {
            const double Jjj = JTJ[j * p + j];
            const double norm = (Jjj <= 0.0) ? 1.0 : sqrt(Jjj);
            if (P.scale == SCALE_LEVENBERG) {
                if (init)
                    diag[j] = 1.0;
            } else if (P.scale == SCALE_MARQUARDT) {
                diag[j] = norm;
            } else {
                const double prev = init ? 0.0 : diag[j];
                diag[j] = prev > norm ? prev : norm;
            }
        }
    }
    TRS_HD void nielsen_init()
    {
        double mx = -1.0;
        for (int j = 0; j < p; ++j) {
            const double val = JTJ[j * p + j] / (diag[j] * diag[j]);
            mx = mx > val ? mx : val;
        }
        nu = 2.0;
        mu = 1.0e-3 * mx;
    }
    TRS_HD void nielsen_accept(double r)
    {
        double b = 2.0 * r - 1.0;
        b = 1.0 - b * b * b;
        nu = 2.0;
        mu *= (0.333333333333333 > b ? 0.333333333333333 : b);
    }
    TRS_HD void nielsen_reject()
    {
        mu *= nu;
        nu *= 2.0;
    }

    // ---------------------------------------------------------------- packet -> (JTJ, g, f2)
    // packet: [JTJ lower packed row-major | JTf | fTf].  Returns false if the Jacobian was non-finite: the
    // reference scans the n x p matrix J itself (src/nls_large.c:515-522); here a NaN or Inf entry of J
    // shows as a NaN or Inf in its column's J^T J entries.  J^T f is NOT part of the scan: a +Inf residual
    // next to a finite Jacobian makes g infinite or NaN and the reference iterates on with it.
    // the lanes share the scan (a private short-circuit loop over the 1176 entries of a p = 48 packet is
    // 1176 dependent global loads: 180 us of a 350 us step, measured)
    TRS_HD bool packet_finite(const double *pk) const
    {
        const int npk = p * (p + 1) / 2;
        bool ok = true;
        for (int e = L.lane(); e < npk; e += L.nlanes())
            ok = ok & finite_d(pk[e]);
        return L.all(ok);
    }
    TRS_HD void take_packet(const double *pk)
    {
        const int ln = L.lane(), nl = L.nlanes();
        const int npk = p * (p + 1) / 2;
        L.sync();
        for (int i = ln; i < p; i += nl)
            for (int j = 0; j <= i; ++j)
                JTJ[i * p + j] = pk[i * (i + 1) / 2 + j];
        L.sync();
        for (int j = 0; j < p; ++j)
            g[j] = pk[npk + j];
        f2 = pk[npk + p];
        nbad = pk[npk + p + 1];
        factor_valid = false;
        jtj_dirty = true;
    }

    // ---------------------------------------------------------------- predicted reductions
    TRS_HD double pred_quadratic(const double *step)
    {
        const double normf = norm_of(f2, nbad);
        double pr = -2.0 * dot(g, step) / (normf * normf);
        symv(step, w3);
        pr -= dot(w3, step) / (normf * normf);
        return pr;
    }
    TRS_HD double pred_lm()
    {
        const double normf = norm_of(f2, nbad);
        const double norm_Dp = scaled_enorm(diag, vel);
        symv(vel, w3);
        const double norm_Jp = sqrt(dot(w3, vel));
        const double u = norm_Jp / normf, v = norm_Dp / normf;
        return u * u + 2.0 * mu * v * v;
    }

    // ---------------------------------------------------------------- dogleg family
    TRS_HD void sd_step()
    {
        for (int i = 0; i < p; ++i)
            w1[i] = g[i] / diag[i];
        norm_Dinvg = nrm2(w1, p);
        for (int i = 0; i < p; ++i)
            w1[i] /= diag[i];
        symv(w1, w2);
        norm_JDinv2g = sqrt(dot(w1, w2));
        const double u = norm_Dinvg / norm_JDinv2g;
        const double alpha = u * u;
        for (int i = 0; i < p; ++i)
            sd[i] = -alpha * w1[i];
        norm_Dsd = scaled_enorm(diag, sd);
    }
    TRS_HD int gn_step()
    {
        const int st = factor(0.0);
        if (st)
            return st;
        solve_neg(g, gn);
        norm_Dgn = scaled_enorm(diag, gn);
        return E_SUCCESS;
    }
    TRS_HD double dogleg_beta(double t, double dl)
    {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < p; ++i) {
            const double d = t * gn[i] + -1.0 * sd[i];
            const double u = diag[i] * d;
            a += u * u;
            w1[i] = d;
        }
        a = sqrt(a);
        a *= a;
        for (int i = 0; i < p; ++i)
            b += sd[i] * (w1[i] * diag[i] * diag[i]);
        b *= 2.0;
        const double c = (norm_Dsd + dl) * (norm_Dsd - dl);
        if (b > 0.0)
            return (-2.0 * c) / (b + sqrt(b * b - 4.0 * a * c));
        return (-b + sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
    }
    TRS_HD int dogleg_step(bool dbl)
    {
        if (norm_Dsd >= delta) {
            for (int i = 0; i < p; ++i)
                dx[i] = sd[i] * (delta / norm_Dsd);
            return E_SUCCESS;
        }
        if (norm_Dgn < 0.0) {
            if (gn_status == E_CONTINUE)
                gn_status = gn_step();
            if (gn_status)
                return gn_status;
        }
        if (norm_Dgn <= delta) {
            for (int i = 0; i < p; ++i)
                dx[i] = gn[i];
            return E_SUCCESS;
        }
        double t = 1.0;
        if (dbl) {
            double v = norm_Dinvg / norm_JDinv2g;
            const double u = v * v;
            v = dot(g, gn);
            const double c = u * (norm_Dinvg / fabs(v)) * norm_Dinvg;
            t = 1.0 - 0.8 * (1.0 - c);
            if (t * norm_Dgn <= delta) {
                for (int i = 0; i < p; ++i)
                    dx[i] = gn[i] * (delta / norm_Dgn);
                return E_SUCCESS;
            }
        }
        const double beta = dogleg_beta(t, delta);
        for (int i = 0; i < p; ++i) {
            const double d = t * gn[i] + -1.0 * sd[i];
            dx[i] = beta * d + 1.0 * sd[i];
        }
        return E_SUCCESS;
    }

    // ---------------------------------------------------------------- 2D subspace
    // Householder pieces on the private p x 2 matrix stored as two columns W0, W1.
    TRS_HD double house(double *v, int n) const
    {
        if (n <= 1)
            return 0.0;
        const double xnorm = nrm2(v + 1, n - 1);
        if (xnorm == 0.0)
            return 0.0;
        const double alpha = v[0];
        const double beta = -(alpha >= 0.0 ? 1.0 : -1.0) * hypot(alpha, xnorm);
        const double tauv = (beta - alpha) / beta;
        const double s = alpha - beta;
        if (fabs(s) > 2.2250738585072014e-308) {
            for (int i = 1; i < n; ++i)
                v[i] *= 1.0 / s;
        } else {
            for (int i = 1; i < n; ++i)
                v[i] *= 2.2204460492503131e-16 / s;
            for (int i = 1; i < n; ++i)
                v[i] *= 1.0 / 2.2204460492503131e-16;
        }
        v[0] = beta;
        return tauv;
    }
    TRS_HD void house_apply(double tauv, const double *v, double *wv, int n) const
    {
        if (tauv == 0.0)
            return;
        double d = wv[0];
        for (int i = 1; i < n; ++i)
            d += v[i] * wv[i];
        wv[0] -= tauv * d;
        for (int i = 1; i < n; ++i)
            wv[i] -= tauv * d * v[i];
    }
    TRS_HD void qt_vec(double *v) const
    {
        house_apply(tau[0], W0, v, p);
        if (p > 1)
            house_apply(tau[1], W1 + 1, v + 1, p - 1);
    }
    TRS_HD void q_vec(double *v) const
    {
        if (p > 1)
            house_apply(tau[1], W1 + 1, v + 1, p - 1);
        house_apply(tau[0], W0, v, p);
    }
    TRS_HD int subspace_preloop()
    {
        gn_status = gn_step();
        if (gn_status)
            return gn_status;
        sd_step();
        for (int i = 0; i < p; ++i) {
            double a = sd[i] * diag[i], b = gn[i] * diag[i];
            if (norm_Dsd != 0.0)
                a *= 1.0 / norm_Dsd;
            if (norm_Dgn != 0.0)
                b *= 1.0 / norm_Dgn;
            W0[i] = a;
            W1[i] = b;
        }
        // column-pivoted QR of [W0 W1]
        double n0 = 0.0, n1 = 0.0;
        for (int i = 0; i < p; ++i) {
            n0 += W0[i] * W0[i];
            n1 += W1[i] * W1[i];
        }
        n0 = sqrt(n0);
        n1 = sqrt(n1);
        if (n1 > n0)
            for (int i = 0; i < p; ++i) {
                const double t = W0[i];
                W0[i] = W1[i];
                W1[i] = t;
            }
        tau[0] = house(W0, p);
        tau[1] = 0.0;
        if (tau[0] != 0.0) {
            double wj = W1[0];
            for (int r = 1; r < p; ++r)
                wj += W1[r] * W0[r];
            W1[0] -= tau[0] * wj;
            for (int r = 1; r < p; ++r)
                W1[r] -= tau[0] * W0[r] * wj;
        }
        if (p > 1)
            tau[1] = house(W1 + 1, p - 1);
        // numerical rank, tolerance of gsl_linalg_QRPT_rank(tol < 0)
        {
            const double d0 = W0[0], d1 = p > 1 ? W1[1] : W0[0];
            const double mn = d0 < d1 ? d0 : d1, mx = d0 > d1 ? d0 : d1;
            const double absmax = fabs(mn) > fabs(mx) ? fabs(mn) : fabs(mx);
            int ee = 0;
            (void)frexp(absmax, &ee);
            const double eps = 20.0 * (double)(p + 2) * ldexp(1.0, ee) * 2.2204460492503131e-16;
            rank = 0;
            if (fabs(d0) > eps)
                ++rank;
            if (p > 1 && fabs(d1) > eps)
                ++rank;
        }
        if (rank == 2) {
            for (int i = 0; i < p; ++i)
                w1[i] = g[i] / diag[i];
            qt_vec(w1);
            const double g0 = w1[0], g1 = w1[1];
            subg[0] = g0;
            subg[1] = g1;
            for (int i = 0; i < p; ++i) {
                w1[i] = 0.0;
                w2[i] = 0.0;
            }
            w1[0] = 1.0;
            w2[1] = 1.0;
            q_vec(w1);
            q_vec(w2);
            for (int i = 0; i < p; ++i) {
                w1[i] /= diag[i];
                w2[i] /= diag[i];
            }
            symv(w1, w3);
            B00 = dot(w1, w3);
            B10 = dot(w2, w3);
            symv(w2, w3);
            B11 = dot(w2, w3);
            trB = B00 + B11;
            detB = B00 * B11 - B10 * B10;
            {
                const double sg[2] = {g0, g1};
                normg = nrm2(sg, 2);
            }
            term0 = (B10 * B10 + B11 * B11) * g0 * g0 - 2.0 * B10 * (B00 + B11) * g0 * g1 +
                    (B10 * B10 + B00 * B00) * g1 * g1;
            term1 = 2.0 * (B11 * g0 * g0 - 2.0 * B10 * g0 * g1 + B00 * g1 * g1);
        }
        return E_SUCCESS;
    }
    // (B + lambda I + E) x = -subg with the 2x2 modified Cholesky of gsl_linalg_mcholesky_decomp
    TRS_HD void sub_solution(double lambda, double *xs) const
    {
        const double eps = 2.2204460492503131e-16;
        double A00 = B00 + lambda, A10 = B10, A11 = B11 + lambda;
        const double gamma = fabs(A00) > fabs(A11) ? fabs(A00) : fabs(A11);
        const double xi = fabs(A10);
        const double nuv = sqrt(3.0);
        double bt = gamma > xi / nuv ? gamma : xi / nuv;
        bt = sqrt(bt > eps ? bt : eps);
        bool sw = false;
        if (fabs(A11) > fabs(A00)) {
            const double t = A00;
            A00 = A11;
            A11 = t;
            sw = true;
        }
        const double u = fabs(A10) / bt;
        double d0 = fabs(A00) > eps ? fabs(A00) : eps;
        d0 = d0 > u * u ? d0 : u * u;
        const double l10 = A10 / d0;
        const double a11 = A11 - A10 * A10 / d0;
        const double d1 = fabs(a11) > eps ? fabs(a11) : eps;
        const double b0 = sw ? subg[1] : subg[0], b1 = sw ? subg[0] : subg[1];
        double y0 = b0, y1 = b1 - l10 * y0;
        y0 /= d0;
        y1 /= d1;
        const double z1 = y1, z0 = y0 - l10 * z1;
        if (sw) {
            xs[0] = -z1;
            xs[1] = -z0;
        } else {
            xs[0] = -z0;
            xs[1] = -z1;
        }
    }
    TRS_HD double sub_objective(const double *xs) const
    {
        const double y0 = subg[0] + 0.5 * (B00 * xs[0] + B10 * xs[1]);
        const double y1 = subg[1] + 0.5 * (B10 * xs[0] + B11 * xs[1]);
        return xs[0] * y0 + xs[1] * y1;
    }
    // real parts of the four roots of a monic quartic (Aberth-Ehrlich iteration in complex double;
    // stands in for gsl_poly_complex_solve's companion-matrix QR)
    TRS_HD void quartic_real_parts(const double *c, double *re) const
    {
        double zr[4], zi[4];
        double R = 0.0;
        for (int i = 0; i < 4; ++i) {
            const double t = pow(fabs(c[i]), 1.0 / (double)(4 - i));
            R = R > t ? R : t;
        }
        R = 2.0 * R;
        if (!(R > 0.0))
            R = 1.0;
        for (int k = 0; k < 4; ++k) {
            const double ang = 1.5707963267948966 * (double)k + 0.4;
            zr[k] = 0.5 * R * cos(ang);
            zi[k] = 0.5 * R * sin(ang);
        }
        for (int it = 0; it < 200; ++it) {
            double maxrel = 0.0;
            for (int k = 0; k < 4; ++k) {
                double pr = 1.0, pi = 0.0, dr = 0.0, di = 0.0;
                for (int i = 3; i >= 0; --i) {
                    const double ndr = dr * zr[k] - di * zi[k] + pr;
                    const double ndi = dr * zi[k] + di * zr[k] + pi;
                    const double npr = pr * zr[k] - pi * zi[k] + c[i];
                    const double npi = pr * zi[k] + pi * zr[k];
                    dr = ndr; di = ndi; pr = npr; pi = npi;
                }
                const double dden = dr * dr + di * di;
                if (dden == 0.0)
                    continue;
                const double wr = (pr * dr + pi * di) / dden, wi = (pi * dr - pr * di) / dden;
                double sr = 0.0, si = 0.0;
                for (int j = 0; j < 4; ++j) {
                    if (j == k)
                        continue;
                    const double er = zr[k] - zr[j], ei = zi[k] - zi[j];
                    const double den = er * er + ei * ei;
                    if (den == 0.0)
                        continue;
                    sr += er / den;
                    si += -ei / den;
                }
                const double qr = 1.0 - (wr * sr - wi * si), qi = -(wr * si + wi * sr);
                const double qden = qr * qr + qi * qi;
                double cr = wr, ci = wi;
                if (qden != 0.0) {
                    cr = (wr * qr + wi * qi) / qden;
                    ci = (wi * qr - wr * qi) / qden;
                }
                zr[k] -= cr;
                zi[k] -= ci;
                const double mag = sqrt(zr[k] * zr[k] + zi[k] * zi[k]);
                const double cm = sqrt(cr * cr + ci * ci);
                const double rel = cm / (mag > 1e-300 ? mag : 1e-300);
                maxrel = maxrel > rel ? maxrel : rel;
            }
            if (maxrel < 1.0e-15)
                break;
        }
        for (int k = 0; k < 4; ++k)
            re[k] = zr[k];
    }
    TRS_HD int subspace_step()
    {
        if (norm_Dgn <= delta) {
            for (int i = 0; i < p; ++i)
                dx[i] = gn[i];
            return E_SUCCESS;
        }
        if (rank < 2) {
            for (int i = 0; i < p; ++i)
                dx[i] = sd[i] * (delta / norm_Dsd);
            return E_SUCCESS;
        }
        const double dsq = delta * delta, u = normg / delta;
        double c[4], re[4], xs[2];
        c[0] = detB * detB - term0 / dsq;
        c[1] = 2.0 * detB * trB - term1 / dsq;
        c[2] = trB * trB + 2.0 * detB - u * u;
        c[3] = 2.0 * trB;
        quartic_real_parts(c, re);
        int mini = -1;
        double minc = 0.0;
        for (int i = 0; i < 4; ++i) {
            sub_solution(re[i], xs);
            const double nx = nrm2(xs, 2);
            if (nx == 0.0)
                continue;
            xs[0] *= delta / nx;
            xs[1] *= delta / nx;
            const double cost = sub_objective(xs);
            if (mini < 0 || cost < minc) {
                mini = i;
                minc = cost;
            }
        }
        if (mini < 0)
            return E_FAILURE;
        sub_solution(re[mini], xs);
        for (int i = 0; i < p; ++i)
            dx[i] = 0.0;
        dx[0] = xs[0];
        dx[1] = xs[1];
        q_vec(dx);
        for (int i = 0; i < p; ++i)
            dx[i] /= diag[i];
        return E_SUCCESS;
    }

    // ---------------------------------------------------------------- Steihaug-Toint CG
    // J d and J^T (J d) are formed from the resident J^T J (||J d||^2 = d^T J^T J d), so the inner
    // CG iterations never touch the n observations; the logical df-callback counts are still kept.
    TRS_HD double cg_tau(const double *z, const double *d, double dl) const
    {
        const double norm_p = nrm2(z, p), norm_q = nrm2(d, p);
        const double u = dot(z, d);
        const double t1 = u / (norm_q * norm_q);
        const double t2 = t1 * u + (dl + norm_p) * (dl - norm_p);
        return -t1 + sqrt(t2) / norm_q;
    }
    TRS_HD int cgst_step()
    {
        double *z = gn, *r = sd, *d = W0; // reuse private vectors
        for (int i = 0; i < p; ++i) {
            z[i] = 0.0;
            r[i] = -g[i] / diag[i];
            d[i] = -g[i] / diag[i];
            w1[i] = g[i] / diag[i];
        }
        const double norm_g = nrm2(w1, p);
        for (long long it = 0; it < P.cg_maxit; ++it) {
            for (int i = 0; i < p; ++i)
                w1[i] = d[i] / diag[i];
            symv(w1, w2); // J^T J D^-1 d
            ndfu += 1.0;
            const double jd2 = dot(w1, w2);
            const double norm_Jd = sqrt(jd2 > 0.0 ? jd2 : 0.0);
            if (norm_Jd == 0.0) {
                const double tauv = cg_tau(z, d, delta);
                for (int i = 0; i < p; ++i)
                    dx[i] = (1.0 * z[i] + tauv * d[i]) / diag[i];
                return E_SUCCESS;
            }
            const double norm_r = nrm2(r, p);
            double u = norm_r / norm_Jd;
            const double alpha = u * u;
            for (int i = 0; i < p; ++i)
                w1[i] = 1.0 * z[i] + alpha * d[i];
            u = nrm2(w1, p);
            if (u >= delta) {
                const double tauv = cg_tau(z, d, delta);
                for (int i = 0; i < p; ++i)
                    dx[i] = (1.0 * z[i] + tauv * d[i]) / diag[i];
                return E_SUCCESS;
            }
            for (int i = 0; i < p; ++i)
                z[i] = w1[i];
            ndfu += 1.0;
            for (int i = 0; i < p; ++i)
                r[i] -= (w2[i] / diag[i]) * alpha;
            const double norm_rp1 = nrm2(r, p);
            if (norm_rp1 / norm_g < P.cg_tol) {
                for (int i = 0; i < p; ++i)
                    dx[i] = z[i] / diag[i];
                return E_SUCCESS;
            }
            u = norm_rp1 / norm_r;
            const double beta = u * u;
            for (int i = 0; i < p; ++i)
                d[i] = 1.0 * r[i] + beta * d[i];
        }
        for (int i = 0; i < p; ++i)
            dx[i] = z[i] / diag[i];
        return E_EMAXITER;
    }

    // ---------------------------------------------------------------- subproblem dispatch
    TRS_HD int preloop()
    {
        norm_Dgn = -1.0;
        gn_status = E_CONTINUE;
        switch (P.trs) {
        case TRS_DOGLEG:
        case TRS_DDOGLEG: sd_step(); return E_SUCCESS;
        case TRS_SUBSPACE2D: return subspace_preloop();
        default: return E_SUCCESS;
        }
    }
    // returns E_SUCCESS with dx set; for lmaccel returns E_CONTINUE after setting vel: the caller
    // must obtain J^T fvv(x, vel) with a MODE_FVV pass and call accel_finish().
    TRS_HD int step()
    {
        switch (P.trs) {
        case TRS_LM:
        case TRS_LMACCEL: {
            const int st = factor(mu);
            if (st)
                return st;
            solve_neg(g, vel);
            if (P.trs == TRS_LMACCEL)
                return E_CONTINUE;
            for (int i = 0; i < p; ++i) {
                acc[i] = 0.0;
                dx[i] = 1.0 * vel[i] + 0.5 * acc[i];
            }
            return E_SUCCESS;
        }
        case TRS_DOGLEG: return dogleg_step(false);
        case TRS_DDOGLEG: return dogleg_step(true);
        case TRS_SUBSPACE2D: return subspace_step();
        default: return cgst_step();
        }
    }
    TRS_HD double preduction()
    {
        if (P.trs == TRS_LM || P.trs == TRS_LMACCEL)
            return pred_lm();
        return pred_quadratic(dx);
    }

    // ---------------------------------------------------------------- convergence test (A.9)
    TRS_HD int test_convergence()
    {
        info = 0;
        bool ok = true;
        const double epsabs = P.xtol * P.xtol, epsrel = P.xtol;
        for (int i = 0; i < p; ++i) {
            const double tol = epsabs + epsrel * fabs(x[i]);
            if (!(fabs(dx[i]) < tol)) {
                ok = false;
                break;
            }
        }
        if (ok) {
            info = 1;
            return E_SUCCESS;
        }
        double gnorm = 0.0;
        for (int i = 0; i < p; ++i) {
            const double xi = x[i] > 1.0 ? x[i] : 1.0;
            const double tmp = fabs(xi * g[i]);
            if (tmp > gnorm)
                gnorm = tmp;
        }
        const double fnorm = norm_of(f2, nbad);
        const double phi = 0.5 * fnorm * fnorm;
        if (gnorm <= P.gtol * (phi > 1.0 ? phi : 1.0)) {
            info = 2;
            return E_SUCCESS;
        }
        return E_CONTINUE;
    }

    // cond(J) as callback_large prints it (src/nls_large.c:733-738): 1 / gsl_multilarge_nlinear_rcond, i.e.
    // 1 / sqrt(rcond_1(J^T J)) from the Cholesky solver's gsl_linalg_cholesky_rcond: ||J^T J||_1 rebuilt from
    // the factor (diagonal) and the saved off-diagonals, ||(J^T J)^-1||_1 by the Hager/Higham estimator
    // (gsl_linalg_invnorm1: at most 5 power-like sweeps + the alternating-sign safeguard vector) -- not the
    // exact norm: the trace shows the number the reference shows.  A failed factorisation prints +Inf.
    TRS_HD double cond_J()
    {
        if (factor(0.0))
            return HUGE_VAL;
        double anorm = 0.0;
        for (int j = 0; j < p; ++j) {
            double ajj = 0.0, sum = 0.0;
            for (int k = 0; k <= j; ++k)
                ajj += A[j * p + k] * A[j * p + k];
            for (int i = 0; i < j; ++i) {
                const double a = fabs(JTJ[j * p + i]);
                sum += a;
                w2[i] += a;
            }
            w2[j] = sum + fabs(ajj);
        }
        for (int i = 0; i < p; ++i)
            anorm = anorm > w2[i] ? anorm : w2[i];
        if (anorm == 0.0)
            return HUGE_VAL;
        double *xv = w1, *v = w2, *xi = w3, *t = W0;
        const int n = p;
        for (int i = 0; i < n; ++i)
            xv[i] = 1.0 / (double)n;
        solve_neg(xv, t);
        double gamma = 0.0;
        for (int i = 0; i < n; ++i) {
            v[i] = -t[i];
            gamma += fabs(v[i]);
            xi[i] = v[i] >= 0.0 ? 1.0 : -1.0;
        }
        solve_neg(xi, t);
        for (int i = 0; i < n; ++i)
            xv[i] = -t[i];
        for (int k = 0; k < 5; ++k) {
            int jmax = 0;
            double amax = 0.0;
            for (int i = 0; i < n; ++i)
                if (fabs(xv[i]) > amax) {
                    amax = fabs(xv[i]);
                    jmax = i;
                }
            for (int i = 0; i < n; ++i)
                xv[i] = (i == jmax) ? 1.0 : 0.0;
            solve_neg(xv, t);
            const double gamma_old = gamma;
            gamma = 0.0;
            bool same = true;
            for (int i = 0; i < n; ++i) {
                v[i] = -t[i];
                gamma += fabs(v[i]);
                same = same && ((v[i] >= 0.0 ? 1.0 : -1.0) == xi[i]);
            }
            if (same || gamma < gamma_old)
                break;
            for (int i = 0; i < n; ++i)
                xi[i] = v[i] >= 0.0 ? 1.0 : -1.0;
            solve_neg(xi, t);
            for (int i = 0; i < n; ++i)
                xv[i] = -t[i];
        }
        double sgn = 1.0;
        for (int i = 0; i < n; ++i) {
            xv[i] = sgn * (1.0 + (double)i / ((double)n - 1.0));
            sgn = -sgn;
        }
        solve_neg(xv, t);
        double alt = 0.0;
        for (int i = 0; i < n; ++i)
            alt += fabs(t[i]);
        alt = 2.0 * alt / (3.0 * (double)n);
        if (alt > gamma)
            gamma = alt;
        if (gamma == 0.0)
            return HUGE_VAL;
        return 1.0 / sqrt((1.0 / anorm) / gamma);
    }

    // ---------------------------------------------------------------- state record I/O
    TRS_HD void load(const double *S)
    {
        phase = (int)S[S_PHASE]; status = (int)S[S_STATUS]; info = (int)S[S_INFO];
        niter = (int)S[S_NITER]; iter = (int)S[S_ITER]; bad = (int)S[S_BAD]; nu = S[S_NU];
        nf = S[S_NEVAL_F]; ndfu = S[S_NEVAL_DFU]; ndf2 = S[S_NEVAL_DF2]; nfvv = S[S_NEVAL_FVV];
        mu = S[S_MU]; delta = S[S_DELTA]; avratio = S[S_AVRATIO]; chisq0 = S[S_CHISQ0];
        chisq1 = S[S_CHISQ1]; f2 = S[S_F2]; chisq_init = S[S_CHISQ_INIT]; npass = S[S_NPASS];
        rho = S[S_RHO]; logdet0 = S[S_LOGDET0]; nbad = S[S_NBAD]; logdet1 = S[S_LOGDET1];
        const double *v = S + S_COUNT;
        for (int i = 0; i < p; ++i) {
            x[i] = v[i]; dx[i] = v[p + i]; g[i] = v[2 * p + i]; diag[i] = v[3 * p + i];
            xt[i] = v[4 * p + i]; vel[i] = v[5 * p + i]; acc[i] = 0.0;
        }
        const int ln = L.lane(), nl = L.nlanes();
        const double *M = S + S_COUNT + 6 * p;
        L.sync();
        for (int e = ln; e < p * p; e += nl)
            JTJ[e] = M[e];
        L.sync();
        factor_valid = false;
        jtj_dirty = false;
        norm_Dgn = -1.0;
        gn_status = E_CONTINUE;
    }
    TRS_HD void store(double *S, bool with_matrix)
    {
        const int ln = L.lane(), nl = L.nlanes();
        if (ln == 0) {
            S[S_PHASE] = phase; S[S_STATUS] = status; S[S_INFO] = info; S[S_NITER] = niter;
            S[S_ITER] = iter; S[S_BAD] = bad; S[S_NU] = nu; S[S_NEVAL_F] = nf; S[S_NEVAL_DFU] = ndfu;
            S[S_NEVAL_DF2] = ndf2; S[S_NEVAL_FVV] = nfvv; S[S_MU] = mu; S[S_DELTA] = delta;
            S[S_AVRATIO] = avratio; S[S_CHISQ0] = chisq0; S[S_CHISQ1] = chisq1; S[S_F2] = f2;
            S[S_CHISQ_INIT] = chisq_init; S[S_NPASS] = npass; S[S_RHO] = rho; S[S_LOGDET0] = logdet0; S[S_NBAD] = nbad;
            S[S_LOGDET1] = logdet1;
            double *v = S + S_COUNT;
            for (int i = 0; i < p; ++i) {
                v[i] = x[i]; v[p + i] = dx[i]; v[2 * p + i] = g[i]; v[3 * p + i] = diag[i];
                v[4 * p + i] = xt[i]; v[5 * p + i] = vel[i];
            }
        }
        if (with_matrix || jtj_dirty) {
            double *M = S + S_COUNT + 6 * p;
            L.sync();
            for (int e = ln; e < p * p; e += nl)
                M[e] = JTJ[e];
        }
    }
    // covariance (J^T J)^-1 into the record (column-major == row-major, symmetric); NaN on failure
    TRS_HD void write_covar(double *S)
    {
        double *C = S + S_COUNT + 6 * p + p * p;
        const int ln = L.lane(), nl = L.nlanes();
        const int st = factor(0.0);
        if (st) {
            L.sync();
            for (int e = ln; e < p * p; e += nl)
                C[e] = NAN;
            return;
        }
        // one column of the inverse per lane, private substitution on the shared factor (the shared solve
        // would spend 2p serial broadcast steps per column; element for element the same arithmetic)
        for (int k = ln; k < p; k += nl) {
            for (int i = 0; i < p; ++i)
                w1[i] = (i == k) ? 1.0 : 0.0;
            solve_neg_private(w1, w3);
            for (int i = 0; i < p; ++i)
                C[i * p + k] = -w3[i];
        }
        L.sync();
    }

    // ---------------------------------------------------------------- the state machine
    // Consume the packet evaluated for the previous request, advance as far as possible without
    // new O(n) information, and write the next request.  trace_* may be null.
    TRS_HD void request(double *req, int mode, const double *theta, const double *v)
    {
        if (L.lane() == 0) {
            req[0] = (double)mode;
            for (int i = 0; i < p; ++i)
                req[1 + i] = theta[i];
            for (int i = 0; i < p; ++i)
                req[1 + p + i] = v ? v[i] : 0.0;
        }
    }
    TRS_HD void trace_row(double *partrace, double *ssrtrace, double *condtrace, int row, double ssr)
    {
        if (!P.trace || !ssrtrace)
            return;
        const double cj = condtrace ? cond_J() : 0.0; // cooperative: all lanes take part
        if (L.lane() == 0) {
            ssrtrace[row] = ssr;
            for (int k = 0; k < p; ++k)
                partrace[row + (P.maxiter + 1) * k] = x[k];
            if (condtrace)
                condtrace[row] = cj;
        }
    }

    TRS_HD void finish(int st, double *S, double *req)
    {
        // src/nls_fit.c:213-221
        if (iter >= P.maxiter && st != E_SUCCESS && st != E_EBADFUNC && !(st == E_ENOPROG && iter == 0))
            st = E_EMAXITER;
        status = st;
        phase = PH_DONE;
        if (L.lane() == 0)
            req[0] = (double)MODE_IDLE;
        if (P.batch_iters > 0) {
            // multi-start screen after the search: log det(J^T J) at the final point (det_cholesky_jtj,
            // src/nls_mstart.c:93, :251); -inf when the factorisation fails
            logdet1 = -HUGE_VAL;
            if (factor(0.0) == E_SUCCESS) {
                double s = 0.0;
                for (int i = 0; i < p; ++i)
                    s += 2.0 * log(A[i * p + i]);
                logdet1 = s;
            }
        }
        store(S, true);
        if (st == E_SUCCESS || st == E_EMAXITER)
            write_covar(S);
    }

    TRS_HD void flush(double *S) { store(S, true); }

    TRS_HD void advance(double *S, const double *pk, double *req, double *partrace, double *ssrtrace,
                        double *condtrace)
    {
        if (keep_state && state_loaded) {
            // what load() resets besides the record itself
            for (int i = 0; i < p; ++i)
                acc[i] = 0.0;
            factor_valid = false;
            norm_Dgn = -1.0;
            gn_status = E_CONTINUE;
        } else {
            load(S);
            state_loaded = true;
        }
        if (phase == PH_DONE)
            return;
        npass += 1.0;

        enum { GO_BEGIN, GO_STEP, GO_END, GO_OUT };
        int go = GO_OUT;
        int end_status = E_SUCCESS;

        if (phase == PH_INIT) {
            nf += 1.0; ndfu += 1.0; ndf2 += 1.0;
            const bool ok = packet_finite(pk);
            take_packet(pk);
            chisq_init = chisq0 = chisq1 = f2;
            if (!ok) {
                info = E_EBADFUNC;
                finish(E_EBADFUNC, S, req);
                return;
            }
            scale_apply(true);
            const double Dx = scaled_enorm(diag, x);
            delta = 0.3 * (Dx > 1.0 ? Dx : 1.0);
            nielsen_init();
            avratio = 0.0;
            for (int i = 0; i < p; ++i)
                dx[i] = 0.0;
            // log det(J^T J) at the start point: the det_cholesky_jtj screen of src/nls_utils.c:55
            logdet0 = -HUGE_VAL;
            if (P.batch_iters > 0 && factor(0.0) == E_SUCCESS) {
                double s = 0.0;
                for (int i = 0; i < p; ++i)
                    s += 2.0 * log(A[i * p + i]);
                logdet0 = s;
            }
            if (P.trace && ssrtrace && L.lane() == 0) {
                ssrtrace[0] = f2;
                for (int k = 0; k < p; ++k)
                    partrace[(P.maxiter + 1) * k] = x[k];
                if (condtrace)
                    condtrace[0] = 0.0;
            }
            go = GO_BEGIN;
        } else if (phase == PH_ACCEL) {
            // pk[0..p) = J^T fvv at (x, vel)
            nfvv += 1.0; ndfu += 1.0;
            bool ok = true;
            for (int i = 0; i < p; ++i)
                ok = ok && finite_d(pk[i]);
            bool step_ok = ok;
            if (ok) {
                // the factor of JTJ + mu D^2 is rebuilt (A is not persisted between launches)
                if (factor(mu) != E_SUCCESS)
                    step_ok = false;
            }
            if (step_ok) {
                for (int i = 0; i < p; ++i)
                    w1[i] = pk[i];
                solve_neg(w1, acc);
                const double anorm = nrm2(acc, p), vnorm = nrm2(vel, p);
                avratio = anorm / vnorm;
                for (int i = 0; i < p; ++i) {
                    dx[i] = 1.0 * vel[i] + 0.5 * acc[i];
                    xt[i] = x[i] + dx[i];
                }
                phase = PH_TRIAL;
                request(req, MODE_FJ, xt, nullptr);
                if (!keep_state) store(S, false);
                return;
            }
            // fvv evaluation failed: counts as a rejected step (src/trust.c:478-482)
            rho = -1.0;
            delta /= P.factor_down;
            nielsen_reject();
            if (++bad > 15) {
                end_status = E_ENOPROG;
                go = GO_END;
            } else {
                go = GO_STEP;
            }
        } else { // PH_TRIAL: packet evaluated at xt
            nf += 1.0;
            const double normf = norm_of(f2, nbad);
            const double f2t = pk[p * (p + 1) / 2 + p];
            const double normf_trial = norm_of(f2t, pk[p * (p + 1) / 2 + p + 1]);
            bool found = true;
            if (P.trs == TRS_LMACCEL && avratio > P.avmax)
                found = false;
            if (normf_trial >= normf) {
                rho = -1.0;
            } else {
                const double u = normf_trial / normf;
                const double actual = 1.0 - u * u;
                const double pred = preduction();
                rho = pred > 0.0 ? actual / pred : -1.0;
            }
            if (rho <= 0.0)
                found = false;
            if (rho > 0.75)
                delta *= P.factor_up;
            else if (rho < 0.25)
                delta /= P.factor_down;
            if (found) {
                const bool ok = packet_finite(pk);
                for (int i = 0; i < p; ++i)
                    x[i] = xt[i];
                take_packet(pk); // speculative J^T J / J^T f of the trial point become current
                ndfu += 1.0; ndf2 += 1.0;
                if (!ok) {
                    end_status = E_EBADFUNC;
                } else {
                    scale_apply(false);
                    nielsen_accept(rho);
                    bad = 0;
                    end_status = E_SUCCESS;
                }
                go = GO_END;
            } else {
                nielsen_reject();
                if (++bad > 15) {
                    end_status = E_ENOPROG;
                    go = GO_END;
                } else {
                    // same outer iteration: subproblem products are a pure function of (g, JTJ, D)
                    (void)preloop();
                    go = GO_STEP;
                }
            }
        }

        for (;;) {
            if (go == GO_BEGIN) {
                chisq0 = chisq1;
                bad = 0;
                const int st = preloop();
                if (st) {
                    end_status = st;
                    go = GO_END;
                } else {
                    go = GO_STEP;
                }
            } else if (go == GO_STEP) {
                const int st = step();
                if (st == E_SUCCESS) {
                    for (int i = 0; i < p; ++i)
                        xt[i] = x[i] + dx[i];
                    phase = PH_TRIAL;
                    request(req, MODE_FJ, xt, nullptr);
                    if (!keep_state) store(S, false);
                    return;
                }
                if (st == E_CONTINUE) { // lmaccel: need J^T fvv(x, vel)
                    phase = PH_ACCEL;
                    request(req, MODE_FVV, x, vel);
                    if (!keep_state) store(S, false);
                    return;
                }
                rho = -1.0;
                delta /= P.factor_down;
                nielsen_reject();
                if (++bad > 15) {
                    end_status = E_ENOPROG;
                    go = GO_END;
                }
            } else { // GO_END: one gsl_multilarge_nlinear_iterate() call has returned end_status
                ++niter;
                chisq1 = f2;
                if (end_status == E_EBADFUNC || (end_status == E_ENOPROG && iter == 0)) {
                    info = end_status;
                    finish(end_status, S, req);
                    return;
                }
                ++iter;
                trace_row(partrace, ssrtrace, condtrace, iter, chisq1);
                int st = test_convergence();
                if (P.batch_iters > 0 && iter >= P.batch_iters && st == E_CONTINUE)
                    st = E_EMAXITER;
                if (st == E_CONTINUE && iter < P.maxiter) {
                    go = GO_BEGIN;
                } else {
                    finish(st == E_CONTINUE ? E_EMAXITER : st, S, req);
                    return;
                }
            }
        }
    }
};

// record initialisation for a new fit
TRS_HD inline void state_reset(double *S, double *req, const double *start, int p)
{
    for (int i = 0; i < state_doubles(p); ++i)
        S[i] = 0.0;
    S[S_PHASE] = PH_INIT;
    S[S_STATUS] = E_CONTINUE;
    S[S_NU] = 2.0;
    for (int i = 0; i < p; ++i)
        S[S_COUNT + i] = start[i];
    req[0] = MODE_FJ;
    for (int i = 0; i < p; ++i) {
        req[1 + i] = start[i];
        req[1 + p + i] = 0.0;
    }
}

} // namespace trs
