"""ctypes front-end of the CPU oracle (oracle/multilarge.c, oracle/dense_model.c).

TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h.  The product package (gslnls_b200) never
imports this module; tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.

Models are given either as one of the built-in C row evaluators ("exp3", "gauss",
"gaussmix", "expmix2") or as a Python callable ``rows(theta, v, want_f, want_J, want_fvv)``
returning numpy arrays; `sympy_rows` builds such a callable from an R-style formula with
sympy (an independent differentiator from the product's C++ one).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRS = {"lm": 0, "lmaccel": 1, "dogleg": 2, "ddogleg": 3, "subspace2D": 4, "cgst": 5}
SCALE = {"more": 0, "levenberg": 1, "marquardt": 2}
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))

_ROWS_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t, C.c_size_t,
                       C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


class _XData(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_double))]


class _Opts(C.Structure):
    _fields_ = [("longdouble", C.c_int), ("threads", C.c_int), ("fd_jac", C.c_int), ("fd_fvv", C.c_int),
                ("weights_gsl", C.c_int)]


class _FitResult(C.Structure):
    _fields_ = [("par", C.POINTER(C.c_double)), ("covar", C.POINTER(C.c_double)), ("ssr", C.c_double),
                ("ssrtol", C.c_double), ("niter", C.c_int), ("conv", C.c_int), ("info", C.c_int),
                ("neval", C.c_size_t * 4), ("partrace", C.POINTER(C.c_double)),
                ("ssrtrace", C.POINTER(C.c_double)), ("chisq_init", C.c_double),
                ("condtrace", C.POINTER(C.c_double)), ("diag", C.POINTER(C.c_double)),
                ("jtj", C.POINTER(C.c_double)), ("xfinal", C.POINTER(C.c_double))]


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("multilarge.c", "dense_model.c", "oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_nls_large.restype = C.c_int
        L.orc_nls_large.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.c_size_t, C.POINTER(C.c_double), C.c_size_t, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(_Opts),
                                    C.POINTER(_FitResult), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_eval_packet.restype = C.c_int
        L.orc_eval_packet.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.POINTER(_Opts),
                                      C.POINTER(C.c_double)]
        L.orc_fit_result_free.argtypes = [C.POINTER(_FitResult)]
        L.orc_strerror.restype = C.c_char_p
        L.orc_strerror.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def control_vectors(algorithm="lm", maxiter=100, trace=False, scale="more", fdtype="forward",
                    factor_up=2.0, factor_down=3.0, avmax=0.75, h_df=SQRT_EPS, h_fvv=0.02,
                    xtol=SQRT_EPS, ftol=SQRT_EPS, gtol=SQRT_EPS):
    """.ctrl_int / .ctrl_dbl exactly as packed by R/nls_large.R:383-407 (defaults R/nls.R:1186-1189)."""
    ci = np.array([int(maxiter), int(bool(trace)), TRS[algorithm], SCALE[scale],
                   {"forward": 0, "center": 1}[fdtype], -2, 0], dtype=np.int32)
    cd = np.array([factor_up, factor_down, avmax, h_df, h_fvv, xtol, ftol, gtol], dtype=np.float64)
    return ci, cd


class _Model:
    """keeps the ctypes callback and its data alive"""

    def __init__(self, model, x=None, p=None):
        L = lib()
        self.keep = []
        if isinstance(model, str):
            fn = getattr(L, "orc_rows_" + model)
            self.fnptr = C.cast(fn, C.c_void_p)
            xa = np.ascontiguousarray(x, dtype=np.float64)
            xd = _XData(_dptr(xa))
            self.keep += [xa, xd]
            self.data = C.cast(C.pointer(xd), C.c_void_p)
        else:
            def cb(theta, v, n, pp, data, fval, J, fvv):
                th = np.ctypeslib.as_array(theta, shape=(pp,)).copy()
                vv = np.ctypeslib.as_array(v, shape=(pp,)).copy() if v else None
                try:
                    f_, J_, h_ = model(th, vv, bool(fval), bool(J), bool(fvv))
                except Exception:  # noqa: BLE001 - surfaces as GSL_EBADFUNC like a failing R closure
                    return 9
                if fval:
                    np.ctypeslib.as_array(fval, shape=(n,))[:] = f_
                if J:
                    np.ctypeslib.as_array(J, shape=(n, pp))[:, :] = J_
                if fvv:
                    np.ctypeslib.as_array(fvv, shape=(n,))[:] = h_
                return 0
            self.cb = _ROWS_FN(cb)
            self.fnptr = C.cast(self.cb, C.c_void_p)
            self.data = None


def nls_large(model, y, start, x=None, weights=None, algorithm="lm", have_fvv=None, longdouble=False,
              threads=0, fd_jac=0, fd_fvv=0, want_resid_grad=False, weights_gsl=False, **control):
    """C_nls_large restated; returns a dict with the fields of the reference's result list.
    weights_gsl=True applies raw weights exactly as the reference does (sqrt(w) on f and fvv inside libgsl,
    J^T J from the unweighted Jacobian); False scales the rows of J too (the product's default mode)."""
    L = lib()
    y = np.ascontiguousarray(y, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.float64)
    n, p = y.size, start.size
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    ci, cd = control_vectors(algorithm=algorithm, **control)
    m = _Model(model, x=x, p=p)
    if have_fvv is None:
        have_fvv = (algorithm == "lmaccel") and not fd_fvv
    opts = _Opts(int(longdouble), int(threads), int(fd_jac), int(fd_fvv), int(bool(weights_gsl)))
    res = _FitResult()
    resid = np.empty(n) if want_resid_grad else None
    grad = np.empty(n * p) if want_resid_grad else None
    status = L.orc_nls_large(m.fnptr, m.data, _dptr(y), _dptr(w), n, _dptr(start), p, int(bool(have_fvv)),
                             ci.ctypes.data_as(C.POINTER(C.c_int)), _dptr(cd), C.byref(opts), C.byref(res),
                             _dptr(resid), _dptr(grad))
    maxiter = int(ci[0])
    out = {
        "par": np.ctypeslib.as_array(res.par, shape=(p,)).copy(),
        "covar": np.ctypeslib.as_array(res.covar, shape=(p, p)).copy().T,
        "niter": res.niter, "conv": res.conv, "info": res.info, "status": L.orc_strerror(status).decode(),
        "ssr": res.ssr, "ssrtol": res.ssrtol, "chisq_init": res.chisq_init,
        "neval": {"f": res.neval[0], "dfu": res.neval[1], "df2": res.neval[2], "fvv": res.neval[3]},
        "x_final": np.ctypeslib.as_array(res.xfinal, shape=(p,)).copy(),
        "diag": np.ctypeslib.as_array(res.diag, shape=(p,)).copy(),
        "jtj": np.tril(np.ctypeslib.as_array(res.jtj, shape=(p, p)).copy()),
    }
    if ci[1]:
        out["partrace"] = np.ctypeslib.as_array(res.partrace, shape=(p, maxiter + 1)).copy().T[: res.niter + 1]
        out["ssrtrace"] = np.ctypeslib.as_array(res.ssrtrace, shape=(maxiter + 1,)).copy()[: res.niter + 1]
        out["condtrace"] = np.ctypeslib.as_array(res.condtrace, shape=(maxiter + 1,)).copy()[: res.niter + 1]
    if want_resid_grad:
        out["resid"] = resid
        out["grad"] = grad.reshape(p, n).T.copy()
    L.orc_fit_result_free(C.byref(res))
    return out


def eval_packet(model, y, theta, x=None, weights=None, longdouble=False, threads=0, fd_jac=0, weights_gsl=False):
    """[J^T J lower packed (row-major (0,0),(1,0),(1,1),..) | J^T f | f^T f] in reference order."""
    L = lib()
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, p = y.size, theta.size
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    m = _Model(model, x=x, p=p)
    opts = _Opts(int(longdouble), int(threads), int(fd_jac), 0, int(bool(weights_gsl)))
    pk = np.zeros(p * (p + 1) // 2 + p + 1)
    s = L.orc_eval_packet(m.fnptr, m.data, _dptr(y), _dptr(w), n, p, _dptr(theta), C.byref(opts), _dptr(pk))
    if s:
        raise RuntimeError("oracle eval_packet failed: %s" % L.orc_strerror(s).decode())
    return pk


# ------------------------------------------------------------------------------------------
# sympy-built row evaluators for arbitrary R-style formulas
# ------------------------------------------------------------------------------------------
def sympy_rows(rhs, param_names, data):
    """rows(theta, v, want_f, want_J, want_fvv) for the R expression `rhs` (e.g. "A*exp(-lam*x)+b")."""
    import sympy as sp

    names = list(param_names)
    dnames = [k for k in data]
    syms = {k: sp.Symbol(k, real=True) for k in names + dnames}
    loc = dict(syms)
    loc["pi"] = sp.pi
    expr = sp.sympify(rhs.replace("^", "**"), locals=loc)
    ps = [syms[k] for k in names]
    ds = [syms[k] for k in dnames]
    darr = [np.asarray(data[k], dtype=np.float64) for k in dnames]
    n = darr[0].size if darr else 1
    f_l = sp.lambdify(ps + ds, expr, "numpy")
    grads = [sp.diff(expr, s) for s in ps]
    g_l = [sp.lambdify(ps + ds, g, "numpy") for g in grads]
    vs = [sp.Symbol("v__%d" % i, real=True) for i in range(len(ps))]
    hexpr = sum(vs[i] * vs[j] * sp.diff(expr, ps[i], ps[j]) for i in range(len(ps)) for j in range(len(ps)))
    h_l = sp.lambdify(ps + vs + ds, hexpr, "numpy")

    def bc(a):
        return np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))

    def rows(theta, v, want_f, want_J, want_fvv):
        args = list(theta) + darr
        with np.errstate(all="ignore"):
            f_ = bc(f_l(*args)) if want_f else None
            J_ = np.stack([bc(g(*args)) for g in g_l], axis=1) if want_J else None
            h_ = bc(h_l(*(list(theta) + list(v) + darr))) if want_fvv else None
        return f_, J_, h_

    return rows


def split_formula(formula):
    lhs, rhs = formula.split("~", 1)
    return lhs.strip(), rhs.strip()
