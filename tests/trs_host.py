"""ctypes driver of tests/host_harness (host build of the device trust-region state machine)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

S_NAMES = ["phase", "status", "info", "niter", "iter", "bad", "nu", "neval_f", "neval_dfu", "neval_df2",
           "neval_fvv", "mu", "delta", "avratio", "chisq0", "chisq1", "f2", "chisq_init", "npass", "rho",
           "logdet0", "nbad"]
S_COUNT = 24


class HostParams(C.Structure):
    _fields_ = [("p", C.c_int), ("maxiter", C.c_int), ("trs", C.c_int), ("scale", C.c_int), ("trace", C.c_int),
                ("batch_iters", C.c_int), ("cg_maxit", C.c_longlong), ("factor_up", C.c_double),
                ("factor_down", C.c_double), ("avmax", C.c_double), ("h_df", C.c_double), ("h_fvv", C.c_double),
                ("xtol", C.c_double), ("ftol", C.c_double), ("gtol", C.c_double), ("cg_tol", C.c_double)]


_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


def lib():
    global _LIB
    if _LIB is None:
        d = os.path.join(_HERE, "host_harness")
        subprocess.check_call(["make", "-C", d, "-s"])
        _LIB = C.CDLL(os.path.join(d, "_build", "libtrs_host.so"))
        _LIB.trs_host_fit.restype = C.c_long
    return _LIB


def packet_from_rows(rows, y, weights=None, weights_gsl=False):
    """packet provider implementing the pass kernel's contract with numpy; weights_gsl=True leaves the rows of
    J unweighted (GSLNLS_WEIGHTS_GSL: only f and fvv carry sqrt(w))"""
    sw = None if weights is None else np.sqrt(np.asarray(weights, dtype=float))
    swJ = None if weights_gsl else sw

    def provider(mode, theta, v):
        p = theta.size
        if mode == 1:
            f, J, _ = rows(theta, None, True, True, False)
            r = np.where(np.isfinite(f), f - y, np.inf)
            if sw is not None:
                r = r * sw
            if swJ is not None:
                J = J * swJ[:, None]
            with np.errstate(all="ignore"):
                JTJ = J.T @ J
                pk = np.concatenate([JTJ[np.tril_indices(p)], J.T @ r, [r @ r], [np.sum(~np.isfinite(r))]])
            return pk
        if mode == 2:
            _, J, h = rows(theta, v, False, True, True)
            if sw is not None:
                h = h * sw
            if swJ is not None:
                J = J * swJ[:, None]
            with np.errstate(all="ignore"):
                return np.concatenate([J.T @ h, [h @ h]])
        raise ValueError(mode)
    return provider


def fit(provider, start, algorithm="lm", maxiter=100, scale="more", trace=True, factor_up=2.0, factor_down=3.0,
        avmax=0.75, h_df=None, h_fvv=0.02, xtol=None, ftol=None, gtol=None, n=1000, batch_iters=0, lanes=1):
    """lanes=1: the SingleLane host build; lanes=32: the warp emulated with one host thread per lane (the
    cooperative loops of trs_core.h run exactly as on the device, with barriers for __syncwarp)"""
    from oracle.oracle import SCALE, SQRT_EPS, TRS
    L = lib()
    start = np.ascontiguousarray(start, dtype=float)
    p = start.size
    hp = HostParams(p, maxiter, TRS[algorithm], SCALE[scale], int(trace), batch_iters, n, factor_up, factor_down,
                    avmax, h_df or SQRT_EPS, h_fvv, xtol or SQRT_EPS, ftol or SQRT_EPS, gtol or SQRT_EPS, 1e-6)
    state = np.zeros(S_COUNT + 6 * p + 2 * p * p)
    partrace = np.zeros((maxiter + 1) * p)
    ssrtrace = np.zeros(maxiter + 1)
    condtrace = np.zeros(maxiter + 1)
    npk = p * (p + 1) // 2 + p + 2

    def cb(ctx, mode, th, v, out):
        theta = np.ctypeslib.as_array(th, shape=(p,)).copy()
        vel = np.ctypeslib.as_array(v, shape=(p,)).copy()
        pk = provider(mode, theta, vel)
        np.ctypeslib.as_array(out, shape=(npk,))[:] = 0.0
        np.ctypeslib.as_array(out, shape=(npk,))[: pk.size] = pk
        return 0
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    if lanes == 1:
        npass = L.trs_host_fit(C.byref(hp), dp(start), _CB(cb), None, dp(state), dp(partrace), dp(ssrtrace),
                               dp(condtrace), 100000)
    else:
        L.trs_host_fit_lanes.restype = C.c_long
        npass = L.trs_host_fit_lanes(C.byref(hp), dp(start), _CB(cb), None, dp(state), dp(partrace), dp(ssrtrace),
                                     dp(condtrace), 100000, int(lanes))
        assert npass >= 0
    out = {k: state[i] for i, k in enumerate(S_NAMES)}
    v = state[S_COUNT:]
    out.update(par=v[:p].copy(), dx=v[p:2 * p].copy(), g=v[2 * p:3 * p].copy(), diag=v[3 * p:4 * p].copy(),
               jtj=v[6 * p:6 * p + p * p].reshape(p, p).copy(),
               covar=v[6 * p + p * p:6 * p + 2 * p * p].reshape(p, p).copy(), npackets=npass, state=state.copy())
    nit = int(out["niter"])
    out["partrace"] = partrace.reshape(p, maxiter + 1).T[: nit + 1].copy()
    out["ssrtrace"] = ssrtrace[: nit + 1].copy()
    out["condtrace"] = condtrace[: nit + 1].copy()
    return out


def multistart(provider, start_range, has_range, algorithm="lm", mstart_n=30, mstart_p=5, mstart_q=None, mstart_r=4.0,
               mstart_s=2, mstart_tol=0.25, mstart_maxiter=10, mstart_maxstart=250, mstart_minsp=1, n=1000):
    """the product's multi-start control logic (gslnls_b200/csrc/mstart.hpp) over the host build of the
    trust-region core; packets for every candidate come from `provider`"""
    from oracle.oracle import SCALE, SQRT_EPS, TRS
    L = lib()
    rng = np.ascontiguousarray(start_range, dtype=float).reshape(-1)
    has = np.ascontiguousarray(np.asarray(has_range, dtype=np.int32).reshape(-1))
    p = rng.size // 2
    hp = HostParams(p, 100, TRS[algorithm], SCALE["more"], 0, 0, n, 2.0, 3.0, 0.75, SQRT_EPS, 0.02, SQRT_EPS,
                    SQRT_EPS, SQRT_EPS, 1e-6)
    npk = p * (p + 1) // 2 + p + 2

    def cb(ctx, mode, th, v, out):
        theta = np.ctypeslib.as_array(th, shape=(p,)).copy()
        vel = np.ctypeslib.as_array(v, shape=(p,)).copy()
        pk = provider(mode, theta, vel)
        np.ctypeslib.as_array(out, shape=(npk,))[:] = 0.0
        np.ctypeslib.as_array(out, shape=(npk,))[: pk.size] = pk
        return 0
    mi = np.array([mstart_n, mstart_p, mstart_q if mstart_q is not None else mstart_n // 10, mstart_s, mstart_maxiter,
                   mstart_maxstart, mstart_minsp], dtype=np.int32)
    md = np.array([mstart_r, mstart_tol], dtype=float)
    out = np.zeros(3 * p + 7)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
    L.trs_host_multistart(C.byref(hp), dp(rng), ip(has), ip(mi), dp(md), _CB(cb), None, dp(out))
    s = out[3 * p:]
    return {"par": out[:p].copy(), "range": out[p:3 * p].reshape(p, 2).copy(), "ssr": s[0], "ssrconv": s[1],
            "nsp": int(s[2]), "nwsp": int(s[3]), "mstarts": int(s[4]), "status": int(s[5]), "searches": int(s[6])}


def qrng(dim, count):
    out = np.zeros((count, dim))
    lib().trs_host_qrng(dim, count, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out
