// mstart.hpp -- control logic of the multi-start global search, on top of BATCHED local searches.
//
// What the reference does one candidate at a time (src/nls_mstart.c:24-349 gsl_multistart_driver, outer loop
// src/nls.c:274-399; method: Hickernell & Yuan 1997, Algorithm 2.1) is reorganised around two batched calls per
// major iteration:
//   concentrate   every sample point takes mstart_p inexpensive LM iterations      (src/nls_mstart.c:75-91)
//   polish        the points that survived mstart_s reductions take mstart_maxiter  (:241-251)
// Each call evaluates all its candidates side by side (on the GPU: candidates ride blockIdx.y of the pass
// kernel and one thread each of the batched trust-region step).  The candidates of a call are independent, so
// everything order-dependent in the reference -- the running best `mssropt`, the 0.99 improvement rule, the
// NSP / NWSP counters, the (1 + tol) gate of the polish stage -- is replayed afterwards on the host over the
// batch results in the reference's candidate order.  Quasi-random points come from restatements of GSL's
// gsl_qrng_sobol (p < 41) and gsl_qrng_halton generators (third-party libgsl, not in /root/reference; call
// sites src/nls.c:277-280, src/nls_mstart.c:47).
//
// Header-only and templated on the evaluator so that tests/host_harness can run the same control code against
// the host build of the trust-region core without a GPU.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

namespace gslnls {
namespace mstart {

// ---------------------------------------------------------------------------------------- quasi-random points
// gsl_qrng_sobol: Bratley & Fox (ACM TOMS 659) Gray-code generator, 30 bits, at most 40 dimensions; the first
// point returned is (0.5, ..., 0.5) -- the origin is never produced.
class Sobol {
public:
    static constexpr int kMaxDim = 40, kBits = 30;
    explicit Sobol(int dim) : dim_(dim), num_(dim, 0), v_(kBits, std::vector<int>(dim, 0)) { init(); }
    void reset()
    {
        count_ = 0;
        std::fill(num_.begin(), num_.end(), 0);
    }
    bool next(double *out)
    {
        int ell = 0;
        for (int c = count_;; c /= 2) { // position of the lowest zero bit of the counter
            ++ell;
            if ((c % 2) == 0)
                break;
        }
        if (ell > kBits)
            return false;
        for (int d = 0; d < dim_; ++d) {
            num_[d] ^= v_[ell - 1][d];
            out[d] = num_[d] * inv_;
        }
        ++count_;
        return true;
    }

private:
    void init()
    {
        static const int poly[kMaxDim] = {1,   3,   7,   11,  13,  19,  25,  37,  59,  47,  61,  55,  41,  67,
                                          97,  91,  109, 103, 115, 131, 193, 137, 145, 143, 241, 157, 185, 167,
                                          229, 171, 213, 191, 253, 203, 211, 239, 247, 285, 369, 299};
        static const int deg[kMaxDim] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 7,
                                         7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8};
        // leading direction numbers m_1..m_deg per dimension (Sobol' & Levitan 1976, as tabulated by Bratley & Fox)
        static const int m1[kMaxDim] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                        1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
        static const int m2[kMaxDim] = {0, 0, 1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3,
                                        1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3};
        static const int m3[kMaxDim] = {0, 0, 0, 7, 5, 1, 3, 3, 7, 5, 5, 7, 7, 1, 3, 3, 7, 5, 1, 1,
                                        5, 3, 3, 1, 7, 5, 1, 3, 3, 7, 5, 1, 1, 5, 7, 7, 5, 1, 3, 3};
        static const int m4[kMaxDim] = {0, 0, 0,  0, 0, 1, 7,  9, 13, 11, 1, 3,  7, 9,  5,  13, 13, 11, 3, 15,
                                        5, 3, 15, 7, 9, 13, 9, 1, 11, 7,  5, 15, 1, 15, 11, 5,  3,  1,  7, 9};
        static const int m5[kMaxDim] = {0,  0,  0, 0,  0,  0,  0,  9,  3,  27, 15, 29, 21, 23, 19, 11, 25, 7,  13, 17,
                                        1,  25, 29, 3, 31, 11, 5,  23, 27, 19, 21, 5,  1,  17, 13, 7,  15, 9,  31, 9};
        static const int m6[kMaxDim] = {0,  0,  0,  0, 0,  0,  0,  0,  0,  0,  0,  0,  0,  37, 33, 7,  5,  11, 39, 63,
                                        27, 17, 15, 23, 29, 3, 21, 13, 31, 25, 9,  49, 33, 19, 29, 11, 19, 27, 15, 25};
        static const int m7[kMaxDim] = {0,  0,  0,   0,  0,  0,  0,  0,   0,  0,  0,   0, 0,  0,  0,  0, 0,   0,  0,  13,
                                        33, 115, 41, 79, 17, 29, 119, 75, 73, 105, 7, 59, 65, 21, 3, 113, 61, 89, 45, 107};
        static const int m8[kMaxDim] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 7, 23, 39};
        static const int *const minit[8] = {m1, m2, m3, m4, m5, m6, m7, m8};
        for (int k = 0; k < kBits; ++k)
            v_[k][0] = 1; // first dimension: van der Corput
        for (int d = 1; d < dim_; ++d) {
            const int dg = deg[d];
            bool inc[8];
            int pbits = poly[d];
            for (int k = dg - 1; k >= 0; --k) { // coefficients of the primitive polynomial, highest first
                inc[k] = (pbits % 2) == 1;
                pbits /= 2;
            }
            for (int j = 0; j < dg; ++j)
                v_[j][d] = minit[j][d];
            for (int j = dg; j < kBits; ++j) { // recurrence of Bratley & Fox, section 2
                int nv = v_[j - dg][d], ell = 1;
                for (int k = 0; k < dg; ++k) {
                    ell *= 2;
                    if (inc[k])
                        nv ^= ell * v_[j - k - 1][d];
                }
                v_[j][d] = nv;
            }
        }
        int ell = 1;
        for (int j = kBits - 2; j >= 0; --j) { // scale column j by 2^(kBits-1-j): common denominator 2^kBits
            ell *= 2;
            for (int d = 0; d < dim_; ++d)
                v_[j][d] *= ell;
        }
        inv_ = 1.0 / (2.0 * ell);
        reset();
    }
    int dim_, count_ = 0;
    double inv_ = 0.0;
    std::vector<int> num_;
    std::vector<std::vector<int>> v_;
};

// gsl_qrng_halton: radical inverse in the first `dim` primes, counter starting at 1
class Halton {
public:
    explicit Halton(int dim) : dim_(dim)
    {
        for (int c = 2; (int)primes_.size() < dim; ++c) {
            bool pr = true;
            for (int q : primes_)
                if (q * q > c)
                    break;
                else if (c % q == 0) {
                    pr = false;
                    break;
                }
            if (pr)
                primes_.push_back(c);
        }
    }
    void reset() { count_ = 0; }
    bool next(double *out)
    {
        ++count_;
        for (int d = 0; d < dim_; ++d) {
            const int b = primes_[d];
            double f = 1.0, r = 0.0;
            for (unsigned k = count_; k > 0; k /= (unsigned)b) {
                f /= (double)b;
                r += f * (double)(k % (unsigned)b);
            }
            out[d] = r;
        }
        return true;
    }

private:
    int dim_;
    unsigned count_ = 0;
    std::vector<int> primes_;
};

// ---------------------------------------------------------------------------------------- batch interface
struct BatchResult { // one candidate after a batched local search
    std::vector<double> par;  // final position
    std::vector<double> diag; // trust-region scaling D at the final position
    double ssr = 0.0;         // chisq1
    double ssr_prev = 0.0;    // chisq0 of the last iteration (ssrconv = ssr_prev - ssr)
    double ssr_start = 0.0;   // f.f at the start point
    double logdet_start = 0.0, logdet_end = 0.0; // log det(J^T J); -inf when the Cholesky factorisation fails
    int status = 0;           // GSL status of the search
};

struct Control {
    int n = 30, p = 5, q = 3, s = 2, niter = 10, max = 250, minsp = 1; // control_int[6..12], src/nls.c:300-306
    double r = 4.0, tol = 0.25;                                        // control_dbl[8], [9]
};

struct Outcome {
    std::vector<double> par; // the start the final fit is launched from (src/nls.c:518-531)
    double ssr = std::numeric_limits<double>::infinity(), ssrconv = 1.0;
    int nsp = 0, nwsp = 0, mstarts = 0;
    int status = 0; // 0: stopping rule met, 11: reached mstart_maxstart major iterations
    std::vector<double> range; // final sampling ranges, 2 per parameter
    long long searches = 0;    // local searches run (concentrate + polish)
};

// Evaluator: void operator()(const std::vector<double> &starts /* S x p */, int S, int iters,
//                            std::vector<BatchResult> &out)   -- `iters` LM-type iterations per candidate with
// the caller's xtol / ftol and gtol = 1e-3 (src/nls_mstart.c:90, :250)
template <class Evaluator>
class Driver {
public:
    Driver(int npar, const Control &c, const double *range /* 2p */, const int *has_range /* 2p */, double xtol,
           double ftol, Evaluator &ev)
        : p_(npar), c_(c), xtol_(xtol), ftol_(ftol), ev_(ev), range_(range, range + 2 * npar),
          range0_(range, range + 2 * npar), maxlims_(range, range + 2 * npar), has_(has_range, has_range + 2 * npar),
          ntix_(c.n, 0), luchange_(npar, 0), x_((size_t)c.n * npar, 0.0), mssr_(c.n, kNA), power_(npar, 1.0),
          best_(npar, 0.0), backup_(npar, 0.0), sobol_(std::min(npar, Sobol::kMaxDim)), halton_(npar)
    {
        // sampling is first concentrated around the centre of fully specified ranges (src/nls.c:347-361)
        for (int k = 0; k < p_; ++k) {
            if (!has_[2 * k] || !has_[2 * k + 1]) {
                power_[k] = 1.0;
                all_ranges_ = false;
            } else {
                power_[k] = 0.75;
                if (range_[2 * k] + xtol_ > range_[2 * k + 1])
                    rejectscl_ = -1.0; // a point, not an interval
            }
        }
    }

    Outcome run()
    {
        int stop = -2; // GSL_CONTINUE
        do { // src/nls.c:364-393
            major_iteration();
            ++mstarts_;
            if (mstarts_ > c_.max)
                stop = 11;
            if (nsp_ >= c_.minsp && (double)nwsp_ > c_.r + std::sqrt(c_.r) * (double)nsp_)
                stop = 0;
            if (mstarts_ % 10 == 0 && !(opt_[0] < kInf)) {
                dtol_ = gmax(0.5 * dtol_, 2.2204460492503131e-16); // nothing found yet: relax the det screen
                if (mstarts_ % 100 == 0)
                    range_ = range0_;
            }
        } while (stop == -2);
        Outcome o;
        o.status = stop;
        o.nsp = nsp_;
        o.nwsp = nwsp_;
        o.mstarts = mstarts_;
        o.range = range_;
        o.searches = searches_;
        const bool use_backup = opt_[1] < opt_[0]; // src/nls.c:518-523
        o.par = use_backup ? backup_ : best_;
        o.ssr = use_backup ? opt_[1] : opt_[0];
        o.ssrconv = use_backup ? conv_[1] : conv_[0];
        if (o.ssr < ftol_ || o.ssrconv < ftol_)
            o.par[0] += 1.0e-4; // :525-531: jitter so that the final fit does not start on the optimum
        return o;
    }

private:
    // GSL_MAX / GSL_MIN are plain comparisons: with a NaN first argument they return the second one
    static double gmax(double a, double b) { return a > b ? a : b; }
    static double gmin(double a, double b) { return a < b ? a : b; }
    static constexpr double kInf = std::numeric_limits<double>::infinity();
    static constexpr double kNA = std::numeric_limits<double>::quiet_NaN();
    static bool is_na(double v) { return v != v; }

    // unit cube -> sampling range, with the power transform that concentrates points (src/nls_mstart.c:48-71)
    void draw(int slot)
    {
        std::vector<double> u(p_);
        if (p_ < 41)
            sobol_.next(u.data());
        else
            halton_.next(u.data());
        for (int k = 0; k < p_; ++k) {
            const double l0 = range_[2 * k], l1 = range_[2 * k + 1];
            double v = l0;
            if (l1 > l0) {
                const double kd = power_[k], t = l0 + (l1 - l0) * u[k];
                if (l0 > 0.0)
                    v = (std::pow(t - l0 + 1.0, kd) - 1.0) / kd + l0;
                else if (l1 < 0.0)
                    v = -(std::pow(-t + l1 + 1.0, kd) - 1.0) / kd + l1;
                else if (t > 0.0)
                    v = (std::pow(t + 1.0, kd) - 1.0) / kd;
                else
                    v = -(std::pow(-t + 1.0, kd) - 1.0) / kd;
            }
            x_[(size_t)slot * p_ + k] = v;
        }
    }

    double current_min() const { return gmin(opt_[0], opt_[1]); }

    void major_iteration()
    {
        const int n = c_.n;
        const double log_dtol = std::log(dtol_);
        // ---- sample: free slots get a fresh quasi-random point; every slot is concentrated ----
        for (int i = 0; i < n; ++i) {
            mssr_[i] = kNA;
            if (ntix_[i] == 0)
                draw(i);
        }
        std::vector<BatchResult> res;
        ev_(x_, n, c_.p, res);
        searches_ += n;
        for (int i = 0; i < n; ++i) { // replay in candidate order (src/nls_mstart.c:74-128)
            const BatchResult &b = res[i];
            if (b.logdet_start > log_dtol) {
                last_prev_ = b.ssr_prev; // the reference's mchisq0 is whatever the last search left in it
                if (b.ssr < kInf) {
                    if (b.logdet_end > log_dtol) {
                        std::copy(b.par.begin(), b.par.end(), x_.begin() + (size_t)i * p_);
                        mssr_[i] = b.ssr;
                        if (b.ssr < 0.99 * current_min()) {
                            opt_[0] = b.ssr;
                            conv_[0] = b.ssr_prev - b.ssr;
                            best_ = b.par;
                        }
                    } else if (b.ssr < 0.99 * current_min()) {
                        opt_[1] = b.ssr;
                        conv_[1] = b.ssr_prev - b.ssr;
                        backup_ = b.par;
                    }
                }
            } else if (!(opt_[0] < kInf) && b.logdet_start > std::log(2.2204460492503131e-16)) {
                // nothing stationary yet: remember the best raw sample as a fall-back
                if (b.ssr_start < 0.99 * opt_[1]) {
                    opt_[1] = b.ssr_start;
                    conv_[1] = last_prev_ - b.ssr_start; // no search ran: mchisq0 is stale (src/nls_mstart.c:122)
                    backup_.assign(x_.begin() + (size_t)i * p_, x_.begin() + (size_t)(i + 1) * p_);
                }
            }
        }
        // ---- reduce: the q best keep their slot and age by one (src/nls_mstart.c:130-138) ----
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { // R_orderVector1, nalast = TRUE
            const bool na = is_na(mssr_[a]), nb = is_na(mssr_[b]);
            if (na != nb)
                return nb;
            return !na && mssr_[a] < mssr_[b];
        });
        for (int r = 0; r < n; ++r) {
            const int i = order[r];
            if (r < c_.q && !is_na(mssr_[i]))
                ntix_[i] += 1;
            else
                ntix_[i] = 0;
        }
        if (!all_ranges_)
            adapt_ranges(order);
        polish_stage();
    }

    // ---- dynamic sampling ranges for parameters without user limits (src/nls_mstart.c:140-235) ----
    void adapt_ranges(const std::vector<int> &order)
    {
        const int n = c_.n;
        double spread = mssr_[order[0]];
        if (!is_na(spread))
            for (int r = n - 1; r > 0; --r)
                if (!is_na(mssr_[order[r]])) {
                    spread -= mssr_[order[r]];
                    break;
                }
        if (is_na(spread) || std::fabs(spread) < 1e-5)
            for (int k = 0; k < p_; ++k)
                luchange_[k] += 1;
        const std::vector<double> &ref = (opt_[1] < opt_[0]) ? backup_ : best_;
        double pmin = 0.0, pmax = 1.0; // NOT reset per parameter while nothing stationary is known (:144)
        for (int k = 0; k < p_; ++k) {
            if (opt_[0] < kInf)
                pmin = pmax = ref[k];
            for (int r = 0; r < std::min(c_.q, n); ++r) {
                const int i = order[r];
                if (ntix_[i] > 0 && mssr_[i] < 1.25 * opt_[0]) {
                    const double v = x_[(size_t)i * p_ + k];
                    pmin = v < pmin ? v : pmin;
                    pmax = v > pmax ? v : pmax;
                }
            }
            const double l0 = range_[2 * k], l1 = range_[2 * k + 1];
            int vote = 0;
            if (!has_[2 * k]) {
                if (pmin < 0.9 * l0 || luchange_[k] > 4) { // the good points crowd the lower edge: widen
                    range_[2 * k] = l0 < 0.0 ? gmax(l0 / std::pow(-1e-5 * (l0 - 1.0), 0.1) - 1.0, -1.0e5) : -0.1;
                    maxlims_[2 * k] = gmin(range_[2 * k], maxlims_[2 * k]);
                    vote = -1;
                } else if (pmin > 0.2 * l0) { // they sit far inside: narrow
                    range_[2 * k] = gmin(l0 / std::pow(-0.05 * (l0 - 1.0), 0.05), -0.01);
                    vote = (opt_[0] < kInf) ? -1 : 1;
                } else {
                    vote = 1;
                }
            }
            if (!has_[2 * k + 1]) {
                if (pmax > 0.9 * l1 || luchange_[k] > 4) {
                    range_[2 * k + 1] = gmin(l1 / std::pow(1e-5 * (l1 + 1.0), 0.1) + 1.0, 1.0e5);
                    maxlims_[2 * k + 1] = gmax(range_[2 * k + 1], maxlims_[2 * k + 1]);
                    vote = -1;
                } else if (pmax < 0.2 * l1) {
                    range_[2 * k + 1] = gmax(l1 / std::pow(0.05 * (l1 + 1.0), 0.05), 0.1);
                    vote = (opt_[0] < kInf) ? -1 : 1;
                } else {
                    vote = 1;
                }
            }
            if (vote)
                luchange_[k] = vote > 0 ? luchange_[k] + 1 : 0;
        }
    }

    // ---- polish: survivors of s reductions get the long local search (src/nls_mstart.c:237-349) ----
    void polish_stage()
    {
        const int n = c_.n;
        std::vector<int> who;
        for (int i = 0; i < n; ++i)
            if (ntix_[i] >= c_.s)
                who.push_back(i);
        if (who.empty())
            return;
        // all survivors are searched in one batch; the (1 + tol) gate below decides afterwards, in candidate
        // order, which of those searches the reference would have run at all
        std::vector<double> starts(who.size() * (size_t)p_);
        for (size_t j = 0; j < who.size(); ++j)
            std::copy(x_.begin() + (size_t)who[j] * p_, x_.begin() + (size_t)(who[j] + 1) * p_, starts.begin() + j * p_);
        std::vector<BatchResult> res;
        ev_(starts, (int)who.size(), c_.niter, res);
        for (size_t j = 0; j < who.size(); ++j) {
            const int i = who[j];
            ntix_[i] = 0;
            nwsp_ += 1;
            if (!(nsp_ == 0 || mssr_[i] < (1.0 + c_.tol) * opt_[0]))
                continue;
            ++searches_;
            const BatchResult &b = res[j];
            const bool nonsingular = b.logdet_end > std::log(dtol_);
            if (b.ssr < kInf && (nsp_ == 0 || b.ssr < 0.99 * opt_[0]) && (nonsingular || b.ssr < 2.0 * ftol_)) {
                bool reject = false;
                if (rejectscl_ > 0.0) { // a stationary point far outside every range seen so far is not trusted
                    for (int k = 0; k < p_ && !reject; ++k) {
                        const double xk = b.par[k], lo = maxlims_[2 * k], hi = maxlims_[2 * k + 1];
                        if (all_ranges_)
                            reject = xk > gmax(hi, 1.0) || xk < gmin(lo, -1.0);
                        else
                            reject = xk > gmax(std::pow(hi, rejectscl_), 1.0) ||
                                     xk < gmin(-std::pow(-lo, rejectscl_), -1.0);
                    }
                    if (!all_ranges_)
                        rejectscl_ += 0.05;
                }
                if (!reject) {
                    opt_[0] = b.ssr;
                    conv_[0] = b.ssr_prev - b.ssr;
                    best_ = b.par;
                    nsp_ += 1;
                    nwsp_ = 0;
                    if (rejectscl_ > 0.0)
                        rejectscl_ = 1.25;
                    if (all_ranges_) { // re-focus the sampling transform with the scaling found at the optimum
                        const double dmin = *std::min_element(b.diag.begin(), b.diag.end());
                        for (int k = 0; k < p_; ++k)
                            power_[k] = std::pow(dmin / b.diag[k], 0.25);
                    }
                }
            } else if (b.ssr < 0.99 * current_min()) {
                opt_[1] = b.ssr;
                conv_[1] = b.ssr_prev - b.ssr;
                backup_ = b.par;
            }
        }
    }

    int p_;
    Control c_;
    double xtol_, ftol_;
    Evaluator &ev_;
    std::vector<double> range_, range0_, maxlims_;
    std::vector<int> has_, ntix_, luchange_;
    std::vector<double> x_, mssr_, power_, best_, backup_;
    Sobol sobol_;
    Halton halton_;
    bool all_ranges_ = true;
    double rejectscl_ = 1.25, dtol_ = 1.0e-6, last_prev_ = kInf;
    double opt_[2] = {kInf, kInf}, conv_[2] = {1.0, 1.0};
    int nsp_ = 0, nwsp_ = 0, mstarts_ = 0;
    long long searches_ = 0;
};

} // namespace mstart
} // namespace gslnls
