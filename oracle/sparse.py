"""oracle/sparse.py -- CPU restatement of gsl_nls_large() with a SPARSE Jacobian and trs = cgst.  TEST
INFRASTRUCTURE ONLY (see oracle/oracle.h): the product never imports this; tests/ and bench.py's CPU legs do.

Data flow of the reference for a dgT/dgC/dgRMatrix Jacobian (src/nls_large.c:528-648): the R closure's sparse
matrix is rebuilt as a triplet gsl_spmatrix on every callback (:575-622), `J u` / `J^T u` are gsl_spblas_dgemv
(:635-639), and J^T J -- whenever libgsl asks for it -- is a dense p x p dsyrk of the densified J (:641-648).  Here
the triplets go into a scipy.sparse CSR matrix and J^T J = J^T @ J stays sparse (same numbers, summed in another
order); everything above the callbacks restates oracle/multilarge.c statement by statement: orc_winit (:1104-1140),
orc_iterate (:1164-1232), cgst_step (:994-1060), quadratic_preduction (:416-426), scale_init / scale_update
(:321-347), orc_test (:1235-1275) and orc_driver2 (:1389-1416).

Parity status: pinned on oracle.nls_large() (dense J, same algorithm; itself pinned on README Example 4's printed
SSR 0.004778845 and on f* = 7.08765e-5 of More'-Garbow-Hillstrom problem 23 at n = 10) by
tests/test_sparse_cpu.py: identical iteration counts and evaluation counters, coefficients / SSR within 1e-10.

A model is a callable  model(theta, want_J) -> (f, (rows, cols, vals) or None)  with f the UNWEIGHTED model
values (y is subtracted here, like the formula path of R/nls_large.R:273).
"""
import math

import numpy as np

SQRT_EPS = float(np.sqrt(np.finfo(float).eps))
EBADFUNC, EMAXITER, ENOPROG = 9, 11, 27
SCALE = {"more": 0, "levenberg": 1, "marquardt": 2}


class _Work:
    pass


def _nrm2(v):
    return math.sqrt(float(v @ v))


def _eval_f(w, x):
    f, _ = w.model(x, False)
    w.nevalf += 1
    f = np.asarray(f, dtype=np.float64) - w.y
    f = np.where(np.isfinite(f), f, np.inf)  # src/nls_large.c:464-465
    return f * w.sw if w.sw is not None else f


def _eval_J(w, x):
    """the df callback's triplet rebuild (src/nls_large.c:575-622); sqrt(W) scales the rows (product default mode)"""
    import scipy.sparse as sps
    _, trip = w.model(x, True)
    r, c, v = trip
    v = np.asarray(v, dtype=np.float64)
    if not np.all(np.isfinite(v)):
        return None  # "Missing/infinite values not allowed when evaluating jac", src/nls_large.c:560-566
    J = sps.csr_matrix((v, (np.asarray(r), np.asarray(c))), shape=(w.n, w.p))
    if w.sw is not None:
        J = sps.diags(w.sw) @ J
    return J.tocsr()


def _df_trans_with_JTJ(w, x, f):
    """eval_df(CblasTrans, x, f, g, JTJ): g = J^T f and J^T J, counted as one dfu and one df2"""
    J = _eval_J(w, x)
    if J is None:
        return EBADFUNC
    w.J = J
    w.JT = J.T.tocsr()
    w.g = np.asarray(w.JT @ f).ravel()
    w.JTJ = (w.JT @ J).tocsr()
    w.nevaldfu += 1
    w.nevaldf2 += 1
    return 0


def _cgst_calc_tau(pv, q, delta):
    norm_p, norm_q, u = _nrm2(pv), _nrm2(q), float(pv @ q)
    t1 = u / (norm_q * norm_q)
    t2 = t1 * u + (delta + norm_p) * (delta - norm_p)
    return -t1 + math.sqrt(t2) / norm_q


def _cgst_step(w, delta):
    """oracle/multilarge.c:994-1060; returns (status, dx)"""
    cgmaxit = w.n
    g, D = w.g, w.diag
    z = np.zeros(w.p)
    r = -g / D
    d = r.copy()
    cg_norm_g = _nrm2(g / D)
    for _ in range(cgmaxit):
        w.cg_iters += 1
        workn = np.asarray(w.J @ (d / D)).ravel()
        w.nevaldfu += 1
        norm_Jd = _nrm2(workn)
        if norm_Jd == 0.0:
            return 0, (z + _cgst_calc_tau(z, d, delta) * d) / D
        norm_r = _nrm2(r)
        u = norm_r / norm_Jd
        alpha = u * u
        znew = z + alpha * d
        if _nrm2(znew) >= delta:
            return 0, (z + _cgst_calc_tau(z, d, delta) * d) / D
        z = znew
        workp = np.asarray(w.JT @ workn).ravel()
        w.nevaldfu += 1
        r = r - (workp / D) * alpha
        norm_rp1 = _nrm2(r)
        if norm_rp1 / cg_norm_g < 1.0e-6:  # cg_tol, GSL default
            return 0, z / D
        u = norm_rp1 / norm_r
        d = r + (u * u) * d
    return EMAXITER, z / D


def _calc_rho(w, f_trial, dx):
    normf, normf_trial = _nrm2(w.f), _nrm2(f_trial)
    if not (normf_trial < normf):
        return -1.0
    u = normf_trial / normf
    actual = 1.0 - u * u
    pred = -2.0 * float(w.g @ dx) / (normf * normf)
    pred -= float(np.asarray(w.JTJ @ dx).ravel() @ dx) / (normf * normf)
    return actual / pred if pred > 0.0 else -1.0


def _iterate(w):
    """oracle/multilarge.c:1164-1232 (orc_iterate) for trs = cgst (no preloop)"""
    bad_steps = 0
    while True:
        status, dx = _cgst_step(w, w.delta)
        found = False
        if status == 0:
            w.dx = dx
            x_trial = w.x + dx
            f_trial = _eval_f(w, x_trial)
            rho = _calc_rho(w, f_trial, dx)
            found = rho > 0.0
        else:
            w.dx = dx
            rho = -1.0
        if rho > 0.75:
            w.delta *= w.factor_up
        elif rho < 0.25:
            w.delta /= w.factor_down
        if found:
            w.x, w.f = x_trial, f_trial
            s = _df_trans_with_JTJ(w, w.x, w.f)
            if s:
                w.niter += 1
                return s
            norm = np.where(w.JTJ.diagonal() <= 0.0, 1.0, np.sqrt(np.maximum(w.JTJ.diagonal(), 0.0)))
            if w.scale == 2:
                w.diag = norm
            elif w.scale == 0:
                w.diag = np.maximum(w.diag, norm)
            w.niter += 1
            return 0
        bad_steps += 1
        if bad_steps > 15:
            w.niter += 1
            return ENOPROG


def _test(w):
    """gsl_multilarge_nlinear_test, oracle/multilarge.c:1235-1275"""
    tol = w.xtol * w.xtol + w.xtol * np.abs(w.x)
    if np.all(np.abs(w.dx) < tol):
        return 1
    gnorm = float(np.max(np.abs(np.maximum(w.x, 1.0) * w.g)))
    fnorm = _nrm2(w.f)
    phi = 0.5 * fnorm * fnorm
    if gnorm <= w.gtol * max(phi, 1.0):
        return 2
    return 0


def nls_large_sparse(model, y, start, weights=None, maxiter=100, scale="more", factor_up=2.0, factor_down=3.0,
                     xtol=SQRT_EPS, gtol=SQRT_EPS, trace=False):
    """C_nls_large with a sparse Jacobian and algorithm = "cgst"; result fields as oracle.nls_large()"""
    w = _Work()
    w.model, w.y = model, np.asarray(y, dtype=np.float64)
    w.x = np.array(start, dtype=np.float64)
    w.n, w.p = w.y.size, w.x.size
    w.sw = None if weights is None else np.sqrt(np.asarray(weights, dtype=np.float64))
    w.scale, w.factor_up, w.factor_down, w.xtol, w.gtol = SCALE[scale], factor_up, factor_down, xtol, gtol
    w.nevalf = w.nevaldfu = w.nevaldf2 = w.cg_iters = w.niter = 0
    w.dx = np.zeros(w.p)
    # orc_winit / trust_init
    w.f = _eval_f(w, w.x)
    chisq_init = float(w.f @ w.f)
    out = {"chisq_init": chisq_init, "par": w.x.copy(), "niter": 0}
    if _df_trans_with_JTJ(w, w.x, w.f):
        out.update({"conv": EBADFUNC, "info": EBADFUNC, "ssr": chisq_init})
        return out
    dj = w.JTJ.diagonal()
    norm = np.where(dj <= 0.0, 1.0, np.sqrt(np.maximum(dj, 0.0)))
    w.diag = np.ones(w.p) if w.scale == 1 else norm.copy()
    w.delta = 0.3 * max(1.0, _nrm2(w.diag * w.x))
    # orc_driver2 (src/nls_fit.c:153-224)
    chisq0 = chisq1 = chisq_init
    ssrtrace = [chisq_init]
    it, status, info = 0, -2, 0
    while True:
        chisq0 = chisq1
        status = _iterate(w)
        chisq1 = float(w.f @ w.f)
        if status == EBADFUNC or (status == ENOPROG and it == 0):
            info = status
            break
        it += 1
        ssrtrace.append(chisq1)
        info = _test(w)
        status = 0 if info else -2
        if not (status == -2 and it < maxiter):
            break
    if it >= maxiter and status != 0:
        status = EMAXITER
    ok = status in (0, EMAXITER)
    out.update({"par": w.x.copy() if ok else np.array(start, dtype=np.float64), "conv": status, "info": info,
                "niter": w.niter, "ssr": chisq1, "ssrtol": chisq0 - chisq1, "x_final": w.x.copy(),
                "grad_vec": w.g.copy(), "diag": w.diag.copy(), "cg_iters": w.cg_iters,
                "neval": {"f": w.nevalf, "dfu": w.nevaldfu, "df2": w.nevaldf2, "fvv": 0}})
    if trace:
        out["ssrtrace"] = np.array(ssrtrace)
    return out


# ------------------------------------------------------------------------------------------ fixtures as models
def penalty_model(p, alpha=1e-5):
    """Penalty function I (README Example 4, inst/unit_tests/unit_tests_gslnls.R:316-346): returns (model, y)"""
    sa = math.sqrt(alpha)
    rows = np.concatenate([np.arange(p), np.full(p, p)])
    cols = np.concatenate([np.arange(p), np.arange(p)])

    def model(th, want_J):
        f = np.concatenate([sa * (th - 1), [np.sum(th ** 2)]])
        return f, ((rows, cols, np.concatenate([np.full(p, sa), 2 * th])) if want_J else None)
    y = np.zeros(p + 1)
    y[p] = 0.25
    return model, y


def grouped_exp_model(g, x, ngroups):
    """y = A[g] * exp(-lam * x) + b[g]; theta = (A[0..G), b[0..G), lam)"""
    n = x.size
    ar = np.arange(n)
    rows = np.concatenate([ar, ar, ar])
    cols = np.concatenate([g, ngroups + g, np.full(n, 2 * ngroups)])

    def model(th, want_J):
        e = np.exp(-th[2 * ngroups] * x)
        f = th[g] * e + th[ngroups + g]
        return f, ((rows, cols, np.concatenate([e, np.ones(n), -th[g] * x * e])) if want_J else None)
    return model
