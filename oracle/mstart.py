"""oracle/mstart.py -- CPU restatement of the reference's multi-start global search.  TEST INFRASTRUCTURE ONLY
(see oracle/oracle.h): the product never imports this.

Follows gsl_multistart_driver (src/nls_mstart.c:24-349) and its outer loop (src/nls.c:274-399, :518-531) one
candidate at a time, in the reference's order, with oracle.nls_large() as the local search (the reference runs
gsl_multifit_nlinear there; on this path the searches are the multilarge-type iterations of the large solver --
the control logic is what is restated).  Quasi-random points: gsl_qrng_sobol / gsl_qrng_halton of libgsl
(third-party, not in /root/reference), restated from Bratley & Fox, ACM TOMS 659, and the radical-inverse
definition; the published sequence heads pin them (tests/test_mstart_cpu.py).
Parity status: the reference's tests check the final coefficients only (inst/unit_tests/unit_tests_gslnls.R:
137-176); per-iteration multi-start state is recorded nowhere -- parity unpinned beyond those end points.
"""
import math

import numpy as np

from . import oracle as O

_POLY = [1, 3, 7, 11, 13, 19, 25, 37, 59, 47, 61, 55, 41, 67, 97, 91, 109, 103, 115, 131, 193, 137, 145, 143, 241,
         157, 185, 167, 229, 171, 213, 191, 253, 203, 211, 239, 247, 285, 369, 299]
_MINIT = {  # dimension (1-based) -> leading direction numbers, Bratley & Fox table
    2: [1], 3: [1, 1], 4: [1, 3, 7], 5: [1, 1, 5], 6: [1, 3, 1, 1], 7: [1, 1, 3, 7], 8: [1, 3, 3, 9, 9],
    9: [1, 3, 7, 13, 3], 10: [1, 1, 5, 11, 27], 11: [1, 3, 5, 1, 15], 12: [1, 1, 7, 3, 29], 13: [1, 3, 7, 7, 21],
    14: [1, 1, 1, 9, 23, 37], 15: [1, 3, 3, 5, 19, 33], 16: [1, 1, 3, 13, 11, 7], 17: [1, 1, 7, 13, 25, 5],
    18: [1, 3, 5, 11, 7, 11], 19: [1, 1, 1, 3, 13, 39], 20: [1, 3, 1, 15, 17, 63, 13],
}


class Sobol:
    """gsl_qrng_sobol for up to 20 dimensions (enough for the oracle's cases): 30-bit Gray-code generator"""
    BITS = 30

    def __init__(self, dim):
        assert 1 <= dim <= 20
        self.dim, self.count = dim, 0
        self.num = [0] * dim
        self.v = [[0] * dim for _ in range(self.BITS)]
        for k in range(self.BITS):
            self.v[k][0] = 1
        for d in range(1, dim):
            poly = _POLY[d]
            deg = poly.bit_length() - 1
            coef = [(poly >> (deg - 1 - k)) & 1 for k in range(deg)]   # a_1 .. a_deg (a_deg = 1)
            m = list(_MINIT[d + 1])
            for j in range(deg, self.BITS):
                nv = m[j - deg]
                for k in range(deg):
                    if coef[k]:
                        nv ^= (2 ** (k + 1)) * m[j - k - 1]
                m.append(nv)
            for j in range(self.BITS):
                self.v[j][d] = m[j]
        for j in range(self.BITS):
            for d in range(dim):
                self.v[j][d] <<= (self.BITS - 1 - j)
        self.inv = 1.0 / float(2 ** self.BITS)

    def next(self):
        c, ell = self.count, 1
        while c & 1:
            c >>= 1
            ell += 1
        for d in range(self.dim):
            self.num[d] ^= self.v[ell - 1][d]
        self.count += 1
        return [n * self.inv for n in self.num]


def _logdet(jtj_lower):
    """log of det_cholesky_jtj (src/nls_utils.c:55-74): (prod diag L)^2, -inf if not positive definite"""
    A = jtj_lower + np.tril(jtj_lower, -1).T
    if not np.all(np.isfinite(A)):
        return -math.inf
    try:
        Lc = np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return -math.inf
    return 2.0 * float(np.sum(np.log(np.diag(Lc))))


def _gmax(a, b):
    return a if a > b else b


def _gmin(a, b):
    return a if a < b else b


def multistart(model, y, start_range, has_range, x=None, weights=None, algorithm="lm", mstart_n=30, mstart_p=5,
               mstart_q=None, mstart_r=4.0, mstart_s=2, mstart_tol=0.25, mstart_maxiter=10, mstart_maxstart=250,
               mstart_minsp=1, **control):
    """returns dict(par, ssr, ssrconv, nsp, nwsp, mstarts, status, range)"""
    rng = [list(map(float, r)) for r in start_range]          # [[l0, l1], ...]
    rng0 = [r[:] for r in rng]
    has = [list(map(bool, h)) for h in has_range]
    p = len(rng)
    n = int(mstart_n)
    q = int(mstart_q) if mstart_q is not None else n // 10
    ctl = dict(control)
    xtol = ctl.get("xtol", O.SQRT_EPS)
    ftol = ctl.get("ftol", O.SQRT_EPS)
    qr = Sobol(p)
    ntix = [0] * n
    luchange = [0] * p
    maxlims = [r[:] for r in rng]
    mx = np.zeros((n, p))
    mssr = [math.nan] * n
    diag = [1.0] * p
    all_start, rejectscl, dtol = True, 1.25, 1.0e-6
    for k in range(p):                                           # src/nls.c:347-361
        if not (has[k][0] and has[k][1]):
            diag[k] = 1.0
            all_start = False
        else:
            diag[k] = 0.75
            if rng[k][0] + xtol > rng[k][1]:
                rejectscl = -1.0
    opt, conv = [math.inf, math.inf], [1.0, 1.0]
    mpopt, mpopt1 = np.zeros(p), np.zeros(p)
    nsp = nwsp = mstarts = 0
    mchisq0 = math.inf

    def search(start, iters):
        r = O.nls_large(model, y, start, x=x, weights=weights, algorithm=algorithm, maxiter=iters,
                        **dict(ctl, gtol=1.0e-3))
        return r

    def det_at(theta):
        pk = O.eval_packet(model, y, theta, x=x, weights=weights)
        J = np.zeros((p, p))
        J[np.tril_indices(p)] = pk[: p * (p + 1) // 2]
        return _logdet(J), float(pk[-1])

    stop = -2
    while stop == -2:
        # ---- gsl_multistart_driver -------------------------------------------------------------------
        for nn in range(n):                                      # :42-128
            mssr[nn] = math.nan
            if ntix[nn] == 0:
                u = qr.next()
                for k in range(p):
                    l0, l1 = rng[k]
                    if l1 > l0:
                        kd = diag[k]
                        t = l0 + (l1 - l0) * u[k]
                        if l0 > 0.0:
                            mx[nn, k] = (math.pow(t - l0 + 1.0, kd) - 1.0) / kd + l0
                        elif l1 < 0.0:
                            mx[nn, k] = -(math.pow(-t + l1 + 1.0, kd) - 1.0) / kd + l1
                        elif t > 0.0:
                            mx[nn, k] = (math.pow(t + 1.0, kd) - 1.0) / kd
                        else:
                            mx[nn, k] = -(math.pow(-t + 1.0, kd) - 1.0) / kd
                    else:
                        mx[nn, k] = l0
            ld0, ssr0 = det_at(mx[nn])
            if ld0 > math.log(dtol):
                r = search(mx[nn].copy(), mstart_p)
                mchisq0 = r["ssr"] + r["ssrtol"]
                mchisq1 = r["ssr"]
                ld1 = _logdet(r["jtj"])
                xfin = r["par"] if r["conv"] in (0, 11) else None
                if xfin is None:                                  # driver2 failed: w->x is still the last accepted point
                    xfin = r["par"]
                if mchisq1 < math.inf:
                    if ld1 > math.log(dtol):
                        mx[nn] = xfin
                        mssr[nn] = mchisq1
                        if mchisq1 < 0.99 * _gmin(opt[0], opt[1]):
                            opt[0], conv[0], mpopt = mchisq1, mchisq0 - mchisq1, np.array(xfin)
                    elif mchisq1 < 0.99 * _gmin(opt[0], opt[1]):
                        opt[1], conv[1], mpopt1 = mchisq1, mchisq0 - mchisq1, np.array(xfin)
            elif not (opt[0] < math.inf) and ld0 > math.log(2.2204460492503131e-16):
                if ssr0 < 0.99 * opt[1]:
                    opt[1], conv[1], mpopt1 = ssr0, mchisq0 - ssr0, mx[nn].copy()
        order = sorted(range(n), key=lambda i: (math.isnan(mssr[i]), mssr[i] if not math.isnan(mssr[i]) else 0.0))
        for r_, i in enumerate(order):                           # :130-138
            if r_ < q and not math.isnan(mssr[i]):
                ntix[i] += 1
            else:
                ntix[i] = 0
        if not all_start:                                        # :140-235
            diff = mssr[order[0]]
            if not math.isnan(diff):
                for r_ in range(n - 1, 0, -1):
                    if not math.isnan(mssr[order[r_]]):
                        diff -= mssr[order[r_]]
                        break
            if math.isnan(diff) or abs(diff) < 1e-5:
                for k in range(p):
                    luchange[k] += 1
            pmin, pmax = 0.0, 1.0
            for k in range(p):
                if opt[0] < math.inf:
                    ref = mpopt1 if opt[1] < opt[0] else mpopt
                    pmin = pmax = float(ref[k])
                for r_ in range(min(q, n)):
                    i = order[r_]
                    if ntix[i] > 0 and mssr[i] < 1.25 * opt[0]:
                        pk = float(mx[i, k])
                        pmin = pk if pk < pmin else pmin
                        pmax = pk if pk > pmax else pmax
                l0, l1 = rng[k]
                add = 0
                if not has[k][0]:
                    if pmin < 0.9 * l0 or luchange[k] > 4:
                        rng[k][0] = _gmax(l0 / math.pow(-1e-5 * (l0 - 1.0), 0.1) - 1.0, -1.0e5) if l0 < 0 else -0.1
                        maxlims[k][0] = _gmin(rng[k][0], maxlims[k][0])
                        add = -1
                    elif pmin > 0.2 * l0:
                        rng[k][0] = _gmin(l0 / math.pow(-0.05 * (l0 - 1.0), 0.05), -0.01)
                        add = -1 if opt[0] < math.inf else 1
                    else:
                        add = 1
                if not has[k][1]:
                    if pmax > 0.9 * l1 or luchange[k] > 4:
                        rng[k][1] = _gmin(l1 / math.pow(1e-5 * (l1 + 1.0), 0.1) + 1.0, 1.0e5)
                        maxlims[k][1] = _gmax(rng[k][1], maxlims[k][1])
                        add = -1
                    elif pmax < 0.2 * l1:
                        rng[k][1] = _gmax(l1 / math.pow(0.05 * (l1 + 1.0), 0.05), 0.1)
                        add = -1 if opt[0] < math.inf else 1
                    else:
                        add = 1
                if add:
                    luchange[k] = luchange[k] + 1 if add > 0 else 0
        for nn in range(n):                                      # :237-349
            if ntix[nn] >= mstart_s:
                ntix[nn] = 0
                nwsp += 1
                if nsp == 0 or mssr[nn] < (1 + mstart_tol) * opt[0]:
                    r = search(mx[nn].copy(), mstart_maxiter)
                    mchisq0 = r["ssr"] + r["ssrtol"]
                    mchisq1 = r["ssr"]
                    ld1 = _logdet(r["jtj"])
                    xfin = r["par"]
                    if mchisq1 < math.inf and (nsp == 0 or mchisq1 < 0.99 * opt[0]) and \
                            (ld1 > math.log(dtol) or mchisq1 < 2 * ftol):
                        reject = 0
                        if rejectscl > 0:
                            for k in range(p):
                                xk = float(xfin[k])
                                if all_start:
                                    reject += int(xk > _gmax(maxlims[k][1], 1.0) or xk < _gmin(maxlims[k][0], -1.0))
                                else:
                                    with np.errstate(all="ignore"):
                                        hi = float(np.power(np.float64(maxlims[k][1]), rejectscl))
                                        lo = -float(np.power(np.float64(-maxlims[k][0]), rejectscl))
                                    reject += int(xk > _gmax(hi, 1.0) or xk < _gmin(lo, -1.0))
                                if reject > 0:
                                    break
                            if not all_start:
                                rejectscl += 0.05
                        if not reject:
                            opt[0], conv[0], mpopt = mchisq1, mchisq0 - mchisq1, np.array(xfin)
                            nsp += 1
                            nwsp = 0
                            if rejectscl > 0:
                                rejectscl = 1.25
                            if all_start:
                                dmin = float(np.min(r["diag"]))
                                for k in range(p):
                                    diag[k] = math.pow(dmin / float(r["diag"][k]), 0.25)
                    elif mchisq1 < 0.99 * _gmin(opt[0], opt[1]):
                        opt[1], conv[1], mpopt1 = mchisq1, mchisq0 - mchisq1, np.array(xfin)
        # ---- src/nls.c:368-393 -------------------------------------------------------------------------
        mstarts += 1
        if mstarts > mstart_maxstart:
            stop = 11
        if nsp >= mstart_minsp and nwsp > (mstart_r + math.sqrt(mstart_r) * nsp):
            stop = 0
        if mstarts % 10 == 0 and not (opt[0] < math.inf):
            dtol = _gmax(0.5 * dtol, 2.2204460492503131e-16)
            if mstarts % 100 == 0:
                rng = [r[:] for r in rng0]
    if opt[1] < opt[0]:                                          # :518-523
        opt[0], conv[0], mpopt = opt[1], conv[1], mpopt1
    par = np.array(mpopt, dtype=float)
    if opt[0] < ftol or conv[0] < ftol:                          # :525-531
        par[0] += 1.0e-4
    return {"par": par, "ssr": opt[0], "ssrconv": conv[0], "nsp": nsp, "nwsp": nwsp, "mstarts": mstarts,
            "status": stop, "range": np.array(rng)}
