"""oracle/irls.py -- CPU restatement of the reference's IRLS driver for robust losses.  TEST INFRASTRUCTURE ONLY
(see oracle/oracle.h): the product never imports this.

Follows gsl_multifit_nlinear_rho_driver (src/nls_irls.c:412-546) statement by statement, with
oracle.nls_large() as the weighted least-squares solver of each iteration; psi functions restate
src/nls_irls.c:10-330 (credited there to robustbase's lmrob.c); default tuning constants R/nls_rho.R:106-116.
Parity status: the reference's tests for this path (inst/unit_tests/unit_tests_gslnls.R:180-215) check final
coefficients against nls() fits with tolerance 1e-4 only; IRLS weights / sigma are recorded nowhere -- parity
unpinned beyond robustbase's published psi formulas, which tests/test_irls_cpu.py checks by their defining
properties (psi odd, psi(x) ~ x at 0, redescending where the family redescends, psi' = d psi / dx numerically).
"""
import math

import numpy as np

from . import oracle as O

LOSSES = {"huber": 1, "barron": 2, "bisquare": 3, "welsh": 4, "optimal": 5, "hampel": 6, "ggw": 7, "lqq": 8}
CC_DEFAULT = {"huber": [1.345], "barron": [1.0, 1.345], "bisquare": [4.685061], "welsh": [2.11],
              "optimal": [1.060158], "hampel": [0.9016085], "ggw": [1.387, 1.5, 1.063], "lqq": [1.473, 0.982, 1.5]}
EPS = 2.2204460492503131e-16


def psi(x, c, which):
    """(psi(x), psi'(x)) for scalar x: src/nls_irls.c:10-330"""
    ax = abs(x)
    if which == 1:
        return (-c[0] if x <= -c[0] else (x if x < c[0] else c[0])), (0.0 if ax >= c[0] else 1.0)
    if which == 2:
        alpha, c2, x2 = c[0], c[1] * c[1], x * x
        if abs(alpha - 2.0) < O.SQRT_EPS:
            return x / c2, 1.0 / c2
        if abs(alpha) < O.SQRT_EPS:
            return 2.0 * x / (x2 + 2 * c2), -2.0 * (x2 - 2.0 * c2) / ((2.0 * c2 + x2) ** 2)
        if alpha > -1e8:
            den = x2 - (alpha - 2.0) * c2
            d = (alpha - 2.0) * ((alpha - 2.0) * c2 - (alpha - 1.0) * x2) * \
                math.pow(1.0 - x2 / ((alpha - 2.0) * c2), 0.5 * alpha) / (den * den)
            return x / c2 * math.pow((x2 / c2) / abs(alpha - 2.0) + 1, 0.5 * alpha - 1.0), d
        return x / c2 * math.exp(-0.5 * x2 / c2), math.exp(-x2 / (2.0 * c2)) * (c2 - x2) / (c2 * c2)
    if which == 3:
        if ax > c[0]:
            return 0.0, 0.0
        a = x / c[0]
        u = 1.0 - a * a
        return x * u * u, (1.0 - a * a) * (1 - 5 * a * a)
    if which == 4:
        a = x / c[0]
        if abs(a) > 37.7:
            return 0.0, 0.0
        e = math.exp(-(a * a) / 2)
        return x * e, e * (1.0 - a * a)
    if which == 5:
        R1, R2, R3, R4 = -1.944, 1.728, -0.312, 0.016
        ac = x / c[0]
        aa = abs(ac)
        if aa > 3.0:
            return 0.0, 0.0
        if aa > 2.0:
            a2 = ac * ac
            v = c[0] * ((((R4 * a2 + R3) * a2 + R2) * a2 + R1) * ac)
            return (max(0.0, v) if ac > 0 else -abs(v)), R1 + a2 * (3 * R2 + a2 * (5 * R3 + a2 * 7 * R4))
        return x, 1.0
    if which == 6:
        a, b, r = 1.5 * c[0], 3.5 * c[0], 8.0 * c[0]
        sx = -1.0 if x < 0 else 1.0
        if ax <= a:
            return x, 1.0
        if ax <= b:
            return sx * a, 0.0
        if ax <= r:
            return sx * a * (r - ax) / (r - b), a / (b - r)
        return 0.0, 0.0
    if which == 7:
        if ax < c[2]:
            return x, 1.0
        ea = -math.pow(ax - c[2], c[1]) / 2 / c[0]
        if ea < -708.4:
            return 0.0, 0.0
        return x * math.exp(ea), math.exp(ea) * (1 - c[1] / (2 * c[0]) * ax * math.pow(ax - c[2], c[1] - 1))
    if which == 8:
        if ax <= c[1]:
            return x, 1.0
        k01 = c[0] + c[1]
        sg = 1.0 if x > 0 else (-1.0 if x < 0 else 0.0)
        if ax <= k01:
            return sg * (ax - c[2] * (ax - c[1]) ** 2 / c[0] / 2.0), 1.0 - c[2] / c[0] * (ax - c[1])
        s5, s6 = c[2] - 1.0, -2 * k01 + c[0] * c[2]
        aa = (c[0] * c[2] - 2 * k01) / (1.0 - c[2])
        d = -(1.0 - c[2]) * ((ax - k01) / aa - 1.0) if ax < k01 + aa else 0.0
        if ax < k01 - s6 / s5:
            return (1.0 if x > 0 else -1.0) * (-s6 / 2.0 - s5 ** 2 / s6 * ((ax - k01) ** 2 / 2.0 + s6 / s5 * (ax - k01))), d
        return 0.0, d
    raise ValueError(which)


def median(a):
    """gsl_median, src/nls_utils.c:162-189"""
    s = np.sort(np.asarray(a))
    n = s.size
    lo, hi = (n - 1) // 2, n // 2
    return float(s[lo]) if lo == hi else float((s[lo] + s[hi]) / 2.0)


def irls(model, y, start, loss="huber", cc=None, x=None, weights=None, rows=None, algorithm="lm", irls_maxiter=50,
         irls_xtol=EPS ** 0.25, **control):
    """returns (fit dict of the last weighted fit, info dict); `rows` evaluates the model values fn(theta)"""
    which = LOSSES[loss]
    c = list(CC_DEFAULT[loss] if cc is None else cc) + [0.0, 0.0]
    y = np.asarray(y, dtype=float)
    n, p = y.size, len(start)
    userw = np.ones(n) if weights is None else np.asarray(weights, dtype=float)
    work = userw.copy()
    prev = np.array(start, dtype=float)
    niter, sigma, delta, status = 0, 1.0, 0.0, -2
    while True:
        niter += 1
        fit = O.nls_large(model, y, start, x=x, weights=work, algorithm=algorithm, **control)
        st = fit["conv"]
        if st == 9 or (st == 27 and niter == 1):
            return fit, {"sigma": sigma, "niter": niter, "status": -1, "delta": delta, "weights": work}
        cur = fit["x_final"]
        resid = rows(cur) - y                                   # unweighted residuals (:488-492)
        sigma = 1.482602218505602 * median(np.abs(resid))
        w = np.empty(n)
        for i in range(n):
            rs = resid[i] / sigma
            ps, _ = psi(rs, c, which)
            with np.errstate(all="ignore"):
                q = np.float64(ps) / np.float64(rs)
            w[i] = q if q > EPS else EPS
        w *= n / np.sum(w)
        work = w * userw
        d = np.abs(prev - cur)
        with np.errstate(all="ignore"):
            rel = d / np.abs(cur)
        delta = float(np.max(d))
        conv = all((r if r < a else a) < irls_xtol for r, a in zip(rel, d))
        if conv:
            return fit, {"sigma": sigma, "niter": niter, "status": 0, "delta": delta, "weights": work}
        prev = cur.copy()
        if niter >= irls_maxiter:
            break
    fit["conv"] = 11
    return fit, {"sigma": sigma, "niter": niter, "status": 11, "delta": delta, "weights": work}
