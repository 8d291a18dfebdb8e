"""gsl_nls_large() -- host-side mirror of the reference's R front-end for the large-NLS path.

Reference: R/nls_large.R:124-445 (formula method) and :459-627 (function method).  Same argument
names, same algorithm names, same validation messages, same control vectors; the closures
(.fn, .jac, .fvv) are replaced by a compiled model (see include/gslnls_b200.h) because R (or
Python) callbacks cannot run on the GPU.  All O(n) work happens in libgslnls_b200.so.
"""
import ctypes as C
import math
import re

import numpy as np

from . import _lib
from .control import ALGORITHMS, gsl_nls_control, pack_control

JAC_MODES = {True: 0, "symbolic": 0, "forward": 1, "center": 2}
# robust losses of gsl_nls_loss() (R/nls_rho.R:101-144): codes of src/nls_irls.c:300-326, default tuning constants
LOSSES = {"huber": 1, "barron": 2, "bisquare": 3, "welsh": 4, "optimal": 5, "hampel": 6, "ggw": 7, "lqq": 8}
LOSS_CC = {"huber": [1.345], "barron": [1.0, 1.345], "bisquare": [4.685061], "welsh": [2.11], "optimal": [1.060158],
           "hampel": [0.9016085], "ggw": [1.387, 1.5, 1.063], "lqq": [1.473, 0.982, 1.5]}


def gsl_nls_loss(rho="default", cc=None):
    """tuning constants of a robust loss, as gsl_nls_loss() (R/nls_rho.R:101-144) returns them"""
    if rho == "default":
        return {"rho": "default", "cc": []}
    if rho not in LOSSES:
        raise ValueError("'arg' should be one of \"default\", %s" % ", ".join('"%s"' % k for k in LOSSES))
    base = LOSS_CC[rho]
    if cc is None:
        cc = list(base)
    elif len(cc) != len(base):
        raise ValueError("'cc' must be of length %d for function '%s'" % (len(base), rho))
    cc = [float(v) for v in cc]
    if rho == "barron" and cc[0] > 2:
        cc[0] = 2.0
    return {"rho": rho, "cc": cc}
# raw weights in the normal equations: "consistent" scales the rows of f and J by sqrt(w) (default);
# "gsl" reproduces the reference -- libgsl scales f / fvv, gsl_df_large's J^T J stays unweighted
WEIGHTS_MODES = {"consistent": 0, "gsl": 1}
FVV_MODES = {None: 0, False: 0, True: 1, "symbolic": 1, "fd": 2}


def _dptr(a):
    return a.ctypes.data_as(_lib.c_double_p) if a is not None else None


class Model:
    """compiled model (gslnls_model): formula RHS -> symbolic J / fvv -> NVRTC kernels"""

    def __init__(self, rhs, param_names, var_names, jac="symbolic", fvv=None):
        L = _lib.lib()
        self.rhs, self.param_names, self.var_names = rhs, list(param_names), list(var_names)
        self.jac_mode, self.fvv_mode = JAC_MODES[jac], FVV_MODES[fvv]
        pn = (C.c_char_p * len(self.param_names))(*[s.encode() for s in self.param_names])
        vn = (C.c_char_p * max(1, len(self.var_names)))(*[s.encode() for s in self.var_names])
        h = C.c_void_p()
        err = C.create_string_buffer(8192)
        rc = L.gslnls_model_compile(rhs.encode(), pn, len(self.param_names), vn, len(self.var_names),
                                    self.jac_mode, self.fvv_mode, C.byref(h), err, len(err))
        if rc:
            raise _lib.GslnlsError(rc, err.value.decode(errors="replace"))
        self.handle = h

    @property
    def source(self):
        return _lib.lib().gslnls_model_source(self.handle).decode()

    @property
    def p(self):
        return len(self.param_names)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().gslnls_model_free(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def _result_to_dict(res, p, n_local, maxiter, trace, want_rg):
    out = {
        "par": np.ctypeslib.as_array(res.par, shape=(p,)).copy(),
        "covar": np.ctypeslib.as_array(res.covar, shape=(p, p)).copy().T,
        "jtj": np.ctypeslib.as_array(res.jtj, shape=(p, p)).copy().T,
        "grad_vec": np.ctypeslib.as_array(res.grad_vec, shape=(p,)).copy(),
        "niter": res.niter, "conv": res.conv, "info": res.info, "status": res.status.decode(),
        "algorithm": res.algorithm.decode(), "ssr": res.ssr, "ssrtol": res.ssrtol, "chisq_init": res.chisq_init,
        "neval": {"f": res.neval[0], "dfu": res.neval[1], "df2": res.neval[2], "fvv": res.neval[3]},
        "npass": res.npass, "n": res.n,
        "x_final": np.ctypeslib.as_array(res.x_final, shape=(p,)).copy(),
    }
    if trace and res.ntrace:
        nt = res.ntrace
        out["partrace"] = np.ctypeslib.as_array(res.partrace, shape=(p, nt)).copy().T
        out["ssrtrace"] = np.ctypeslib.as_array(res.ssrtrace, shape=(nt,)).copy()
        out["condtrace"] = np.ctypeslib.as_array(res.condtrace, shape=(nt,)).copy()
    if want_rg and res.resid:
        out["resid"] = np.ctypeslib.as_array(res.resid, shape=(n_local,)).copy()
        out["grad"] = np.ctypeslib.as_array(res.grad, shape=(p, n_local)).copy().T
    return out


class Problem:
    """device-resident data + solver workspace (gslnls_problem)"""

    def __init__(self, model, n, has_weights=False, device=0):
        self.model, self.n, self.has_weights, self.device = model, int(n), bool(has_weights), device
        h = C.c_void_p()
        _lib.check(_lib.lib().gslnls_problem_create(model.handle, self.n, int(self.has_weights), device, C.byref(h)))
        self.handle = h
        self._keep = []

    def upload(self, vars_, y, weights=None):
        cols = [np.ascontiguousarray(v, dtype=np.float64) for v in vars_]
        y = np.ascontiguousarray(y, dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        arr = (_lib.c_double_p * max(1, len(cols)))(*[_dptr(c) for c in cols])
        _lib.check(_lib.lib().gslnls_problem_upload(self.handle, arr, _dptr(y), _dptr(w)))
        return self

    def bind_device(self, var_ptrs, y_ptr, w_ptr=None, keepalive=()):
        """use device buffers owned by the caller (integers, e.g. torch.Tensor.data_ptr())"""
        arr = (C.c_void_p * max(1, len(var_ptrs)))(*[C.c_void_p(int(v)) for v in var_ptrs])
        _lib.check(_lib.lib().gslnls_problem_bind_device(self.handle, arr, C.c_void_p(int(y_ptr)),
                                                         C.c_void_p(int(w_ptr)) if w_ptr else None))
        self._keep = list(keepalive)
        return self

    def set_weights_mode(self, mode):
        _lib.check(_lib.lib().gslnls_problem_set_weights_mode(self.handle, WEIGHTS_MODES[mode]))
        return self

    def set_comm(self, comm):
        self.comm = comm
        _lib.check(_lib.lib().gslnls_problem_set_comm(self.handle, comm.handle if comm else None))
        return self

    def eval_packet(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        p = th.size
        pk = np.zeros(p * (p + 1) // 2 + p + 1)
        _lib.check(_lib.lib().gslnls_problem_eval_packet(self.handle, _dptr(th), _dptr(pk)))
        return pk

    def eval_jtfvv(self, theta, v):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        vv = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(th.size)
        _lib.check(_lib.lib().gslnls_problem_eval_jtfvv(self.handle, _dptr(th), _dptr(vv), _dptr(out)))
        return out

    def time_passes(self, theta, npass):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        ms = C.c_float()
        _lib.check(_lib.lib().gslnls_problem_time_passes(self.handle, _dptr(th), int(npass), C.byref(ms)))
        return ms.value

    def residuals(self, theta, want_grad=False):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        r = np.empty(self.n)
        g = np.empty(self.n * th.size) if want_grad else None
        _lib.check(_lib.lib().gslnls_problem_residuals(self.handle, _dptr(th), _dptr(r), _dptr(g)))
        return (r, g.reshape(th.size, self.n).T.copy()) if want_grad else r

    def fit(self, start, algorithm="lm", control=None, trace=False, want_resid_grad=False):
        ctrl = gsl_nls_control() if control is None else control
        ci, cd = pack_control(ctrl, algorithm, trace)
        st = np.ascontiguousarray(start, dtype=np.float64)
        res = _lib.Result()
        rc = _lib.lib().gslnls_problem_fit(self.handle, _dptr(st), ci.ctypes.data_as(_lib.c_int_p), _dptr(cd),
                                           int(want_resid_grad), C.byref(res))
        _lib.check(rc)
        out = _result_to_dict(res, st.size, self.n, int(ci[0]), trace, want_resid_grad)
        _lib.lib().gslnls_result_free(C.byref(res))
        return out

    # --- stepwise interface used by bench.py -------------------------------------------------
    def fit_begin(self, start, algorithm="lm", control=None, trace=False, packed=None):
        """packed=(ci, cd, start_array) re-uses control vectors packed once (benchmark loops)"""
        if packed is None:
            ctrl = gsl_nls_control() if control is None else control
            ci, cd = pack_control(ctrl, algorithm, trace)
            st = np.ascontiguousarray(start, dtype=np.float64)
        else:
            ci, cd, st = packed
        self._fit = (st.size, int(ci[0]), trace)
        _lib.check(_lib.lib().gslnls_problem_fit_begin(self.handle, _dptr(st), ci.ctypes.data_as(_lib.c_int_p),
                                                       _dptr(cd)))

    def fit_run(self, max_passes=0, want_ms=False):
        """run up to max_passes trial steps; want_ms=True also returns the device time of the call, which
        costs a stream synchronisation that a finished fit otherwise does not need"""
        done, run, ms = C.c_int(), C.c_int64(), C.c_float()
        _lib.check(_lib.lib().gslnls_problem_fit_run(self.handle, int(max_passes), C.byref(done), C.byref(run),
                                                     C.byref(ms) if want_ms else None))
        return bool(done.value), run.value, (ms.value if want_ms else None)

    def fit_end(self, want_resid_grad=False, light=False):
        """light=True returns only the scalars (niter, npass, conv, ssr) and par: the full record costs ~40 us
        of numpy conversions, which a benchmark loop over back-to-back fits should not charge to the device"""
        p, maxiter, trace = self._fit
        res = _lib.Result()
        rc = _lib.lib().gslnls_problem_fit_end(self.handle, int(want_resid_grad), C.byref(res))
        _lib.check(rc)
        if light:
            out = {"niter": res.niter, "npass": res.npass, "conv": res.conv, "ssr": res.ssr,
                   "par": [res.par[i] for i in range(p)], "status": res.status.decode()}
            _lib.lib().gslnls_result_free(C.byref(res))
            return out
        out = _result_to_dict(res, p, self.n, maxiter, trace, want_resid_grad)
        _lib.lib().gslnls_result_free(C.byref(res))
        return out

    def fit_batch(self, starts, iters=5, algorithm="lm", control=None):
        """batched multi-start inner loops: det(J^T J) screen + `iters` iterations per start point"""
        ctrl = dict(gsl_nls_control() if control is None else control)
        ctrl["maxiter"] = int(iters)
        ci, cd = pack_control(ctrl, algorithm, False)
        st = np.ascontiguousarray(starts, dtype=np.float64)
        S, p = st.shape
        par, ssr, ld = np.empty((S, p)), np.empty(S), np.empty(S)
        conv, nit = np.empty(S, dtype=np.int32), np.empty(S, dtype=np.int32)
        _lib.check(_lib.lib().gslnls_problem_fit_batch(self.handle, _dptr(st), S, ci.ctypes.data_as(_lib.c_int_p),
                                                       _dptr(cd), _dptr(par), _dptr(ssr), _dptr(ld),
                                                       conv.ctypes.data_as(_lib.c_int_p),
                                                       nit.ctypes.data_as(_lib.c_int_p)))
        return {"par": par, "ssr": ssr, "logdet": ld, "conv": conv, "niter": nit}

    def fit_irls(self, start, loss="huber", cc=None, algorithm="lm", control=None):
        """robust fit by iteratively reweighted least squares (src/nls_irls.c:412-546); the problem must have a
        weights column (has_weights=True; upload ones when the user has none).  Returns (fit, irls info)."""
        ctrl = gsl_nls_control() if control is None else gsl_nls_control(**dict(control))
        ci, cd = pack_control(ctrl, algorithm, False)
        st = np.ascontiguousarray(start, dtype=np.float64)
        lc = gsl_nls_loss(loss, cc)
        ccv = np.array(list(lc["cc"]) + [0.0] * (3 - len(lc["cc"])), dtype=np.float64)
        res, info = _lib.Result(), _lib.IrlsInfo()
        rc = _lib.lib().gslnls_problem_fit_irls(self.handle, _dptr(st), ci.ctypes.data_as(_lib.c_int_p), _dptr(cd),
                                                LOSSES[lc["rho"]], _dptr(ccv), int(ctrl["irls_maxiter"]),
                                                float(ctrl["irls_xtol"]), C.byref(res), C.byref(info))
        _lib.check(rc)
        out = _result_to_dict(res, st.size, self.n, int(ci[0]), False, False)
        _lib.lib().gslnls_result_free(C.byref(res))
        return out, {"sigma": info.sigma, "delta": info.delta, "niter": info.niter, "status": info.status}

    def weights(self):
        w = np.empty(self.n)
        _lib.check(_lib.lib().gslnls_problem_get_weights(self.handle, _dptr(w)))
        return w

    def median_abs_resid(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        m = C.c_double()
        _lib.check(_lib.lib().gslnls_problem_median_abs_resid(self.handle, _dptr(th), C.byref(m)))
        return m.value

    def multistart(self, start_range, has_range=None, algorithm="lm", control=None):
        """multi-start global search (gsl_multistart_driver, src/nls_mstart.c + src/nls.c:274-399) over batched
        local searches: returns the start vector for the final fit plus the search statistics.
        start_range  p pairs (lower, upper); has_range  p pairs of flags, False = side adapts dynamically"""
        ctrl = gsl_nls_control() if control is None else gsl_nls_control(**dict(control))
        ci, cd = pack_control(ctrl, algorithm, False)
        rng = np.ascontiguousarray(start_range, dtype=np.float64).reshape(-1)
        p = rng.size // 2
        has = np.ones(2 * p, dtype=np.int32) if has_range is None else \
            np.ascontiguousarray(np.asarray(has_range, dtype=np.int32).reshape(-1))
        r = ctrl["mstart_r"] * (10 if not has.all() else 1)     # R/nls.R:711-713
        mi = np.array([ctrl[k] for k in ("mstart_n", "mstart_p", "mstart_q", "mstart_s", "mstart_maxiter",
                                         "mstart_maxstart", "mstart_minsp")], dtype=np.int32)
        md = np.array([r, ctrl["mstart_tol"]], dtype=np.float64)
        res = _lib.MstartResult()
        _lib.check(_lib.lib().gslnls_problem_multistart(self.handle, _dptr(rng), has.ctypes.data_as(_lib.c_int_p),
                                                        ci.ctypes.data_as(_lib.c_int_p), _dptr(cd),
                                                        mi.ctypes.data_as(_lib.c_int_p), _dptr(md), C.byref(res)))
        out = {"par": np.ctypeslib.as_array(res.par, shape=(p,)).copy(),
               "range": np.ctypeslib.as_array(res.range, shape=(p, 2)).copy(), "ssr": res.ssr,
               "ssrconv": res.ssrconv, "nsp": res.nsp, "nwsp": res.nwsp, "mstarts": res.mstarts,
               "status": res.status, "searches": res.searches}
        _lib.lib().gslnls_mstart_result_free(C.byref(res))
        return out

    def timer_start(self):
        _lib.check(_lib.lib().gslnls_problem_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        _lib.check(_lib.lib().gslnls_problem_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def set_profile(self, max_passes):
        _lib.check(_lib.lib().gslnls_problem_set_profile(self.handle, int(max_passes)))

    def profile(self):
        ms, cnt = C.c_float(), C.c_int64()
        _lib.check(_lib.lib().gslnls_problem_profile(self.handle, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def channel_stats(self, reset=False):
        """device clocks of the resident-server mode: (avg pass µs request->packet, avg step µs packet->request, passes)"""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(_lib.lib().gslnls_problem_channel_stats(self.handle, int(reset), C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    @property
    def launch_count(self):
        return _lib.lib().gslnls_problem_launch_count(self.handle)

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().gslnls_problem_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Session:
    """gslnls_session: the data of a fit resident on one or several GPUs of this process.  What the fitted
    object keeps so that residuals / gradient (src/nls_large.c:339-385) are produced only when asked for."""

    def __init__(self, model, n, has_weights=False, devices=(0,)):
        self.model, self.n, self.has_weights = model, int(n), bool(has_weights)
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        _lib.check(_lib.lib().gslnls_session_create(model.handle, self.n, int(self.has_weights), dev.size,
                                                    dev.ctypes.data_as(_lib.c_int_p), C.byref(h)))
        self.handle = h
        self.device = int(dev[0])

    @property
    def ngpu(self):
        return _lib.lib().gslnls_session_ngpu(self.handle)

    def set_weights_mode(self, mode):
        _lib.check(_lib.lib().gslnls_session_set_weights_mode(self.handle, WEIGHTS_MODES[mode]))
        return self

    def upload(self, vars_, y, weights=None):
        cols = [np.ascontiguousarray(v, dtype=np.float64) for v in vars_]
        y = np.ascontiguousarray(y, dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        arr = (_lib.c_double_p * max(1, len(cols)))(*[_dptr(c) for c in cols])
        _lib.check(_lib.lib().gslnls_session_upload(self.handle, arr, _dptr(y), _dptr(w)))
        return self

    def fit(self, start, algorithm="lm", control=None, trace=False, want_resid_grad=False):
        ctrl = gsl_nls_control() if control is None else control
        ci, cd = pack_control(ctrl, algorithm, trace)
        st = np.ascontiguousarray(start, dtype=np.float64)
        res = _lib.Result()
        rc = _lib.lib().gslnls_session_fit(self.handle, _dptr(st), ci.ctypes.data_as(_lib.c_int_p), _dptr(cd),
                                           int(want_resid_grad), C.byref(res))
        _lib.check(rc)
        out = _result_to_dict(res, st.size, self.n, int(ci[0]), trace, want_resid_grad)
        _lib.lib().gslnls_result_free(C.byref(res))
        return out

    def residuals(self, theta, want_grad=False):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        r = np.empty(self.n)
        g = np.empty(self.n * th.size) if want_grad else None
        _lib.check(_lib.lib().gslnls_session_residuals(self.handle, _dptr(th), _dptr(r), _dptr(g)))
        return (r, g.reshape(th.size, self.n).T.copy()) if want_grad else r

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().gslnls_session_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ------------------------------------------------------------------------------------------------
# formula handling (R/nls_large.R:141-283)
# ------------------------------------------------------------------------------------------------
_NAME = re.compile(r"[A-Za-z.][A-Za-z0-9._]*")
_KNOWN = {"exp", "log", "log2", "log10", "log1p", "expm1", "sqrt", "sin", "cos", "tan", "asin", "acos", "atan",
          "sinh", "cosh", "tanh", "abs", "sign", "pnorm", "dnorm", "sinpi", "cospi", "pi", "I", "SSasymp",
          "SSasympOff", "SSasympOrig", "SSbiexp", "SSfpl", "SSgompertz", "SSlogis", "SSmicmen", "SSweibull"}


def all_vars(expr):
    """names referenced by an R expression, in order of first appearance (all.vars)"""
    seen = []
    for m in _NAME.finditer(expr):
        nm = m.group(0)
        j = m.end()
        while j < len(expr) and expr[j] == " ":
            j += 1
        if (j < len(expr) and expr[j] == "(") or nm in _KNOWN:
            continue
        if re.fullmatch(r"[0-9.]+([eE][-+]?[0-9]+)?L?", nm):
            continue
        if nm not in seen:
            seen.append(nm)
    return seen


def _eval_lhs(lhs, data):
    lhs = lhs.strip()
    if lhs in data:
        return np.asarray(data[lhs], dtype=np.float64)
    env = {k: np.asarray(v, dtype=np.float64) for k, v in data.items()}
    env.update({f: getattr(np, f) for f in ("exp", "log", "log2", "log10", "log1p", "expm1", "sqrt", "sin", "cos",
                                             "tan", "abs")})
    return np.asarray(eval(lhs.replace("^", "**"), {"__builtins__": {}}, env), dtype=np.float64)  # noqa: S307


class GslNls:
    """The fitted-model object: what gsl_nls_large() returns as class c("gsl_nls", "nls")
    (R/nls_large.R:416-443), with the S3 methods of R/nls_methods.R as Python methods."""

    def __init__(self, formula, param_names, cfit, problem, model, control, algorithm, weights, lhs, trace):
        self.formula, self.param_names = formula, list(param_names)
        self.cfit, self._problem, self._model = cfit, problem, model
        self.control, self.algorithm, self.weights, self._lhs = control, algorithm, weights, lhs
        self.convInfo = {
            "isConv": not cfit["conv"], "finIter": cfit["niter"], "finTol": cfit["ssrtol"], "nEval": cfit["neval"],
            "trsName": "multilarge/" + cfit["algorithm"], "stopCode": cfit["conv"], "stopMessage": cfit["status"],
        }
        if trace and "partrace" in cfit:
            k = cfit["niter"] + 1
            self.partrace = cfit["partrace"][:k]
            self.devtrace = cfit["ssrtrace"][:k]
        self._resid = None

    # coef / deviance / nobs / df.residual / sigma / vcov ------------------------------------------
    def coef(self):
        return dict(zip(self.param_names, self.cfit["par"]))

    def deviance(self):
        return self.cfit["ssr"]

    def nobs(self):
        return int(self.cfit["n"])

    def df_residual(self):
        return self.nobs() - len(self.param_names)

    def sigma(self):
        return math.sqrt(self.deviance() / self.df_residual())

    def vcov(self):
        """sigma^2 (J^T J)^-1; the reference gets R^-1 from a QR of the n x p gradient (R/nls.R:1295)"""
        return self.sigma() ** 2 * self.cfit["covar"]

    def Rmat(self):
        """upper-triangular R with R^T R = J^T J (m$Rmat()), from the p x p device result"""
        return np.linalg.cholesky(self.cfit["jtj"]).T

    def residuals(self):
        """response residuals y - fitted, weighted like m$resid() = -cFit$resid (R/nls.R:1255); computed from
        the device-resident data on first use, never by the fit itself"""
        if self._resid is None:
            self._resid = -self._problem.residuals(self.cfit["par"])
        return self._resid

    def gradient(self):
        """m$gradient(): the n x p (weighted) Jacobian at the estimates, on demand"""
        return self._problem.residuals(self.cfit["par"], want_grad=True)[1]

    def fitted(self):
        r = self.residuals()
        sw = 1.0 if self.weights is None else np.sqrt(self.weights)
        return self._lhs - r / sw

    def logLik(self):
        n = self.nobs()
        w = np.ones(n) if self.weights is None else np.asarray(self.weights)
        return -n / 2.0 * (math.log(2 * math.pi) + 1 - math.log(n) - np.sum(np.log(w)) / n + math.log(self.deviance()))

    def summary(self):
        from scipy import stats
        est = self.cfit["par"]
        se = np.sqrt(np.diag(self.vcov()))
        t = est / se
        pv = 2 * stats.t.sf(np.abs(t), self.df_residual())
        return {"coefficients": {n: dict(estimate=e, std_error=s, t_value=tv, p_value=p) for n, e, s, tv, p in
                                 zip(self.param_names, est, se, t, pv)},
                "sigma": self.sigma(), "df": (len(est), self.df_residual()), "convInfo": self.convInfo}

    def confint(self, level=0.95):
        from scipy import stats
        est = self.cfit["par"]
        se = np.sqrt(np.diag(self.vcov()))
        q = stats.t.ppf(0.5 + level / 2, self.df_residual())
        return {n: (e - q * s, e + q * s) for n, e, s in zip(self.param_names, est, se)}

    def predict(self, newdata):
        """evaluate the fitted curve on new predictor values (on the device, like everything O(n))"""
        cols = [np.ascontiguousarray(newdata[v], dtype=np.float64) for v in self._model.var_names]
        n = cols[0].size if cols else 1
        pb = Problem(self._model, n, False, self._problem.device)
        pb.upload(cols, np.zeros(n))
        out = pb.residuals(self.cfit["par"])
        pb.close()
        return out

    def __repr__(self):
        co = ", ".join("%s=%.6g" % kv for kv in self.coef().items())
        return ("Nonlinear regression model\n  model: %s\n  %s\n residual sum-of-squares: %.4g\n\n"
                "Algorithm: %s, (scaling: %s, solver: cholesky)\n\nNumber of iterations%s: %d\n"
                "Achieved convergence tolerance: %.4g" % (
                    self.formula, co, self.deviance(), self.convInfo["trsName"], self.control["scale"],
                    " to convergence" if self.convInfo["isConv"] else " till stop", self.convInfo["finIter"],
                    self.convInfo["finTol"]))


def fit_large_multi(model, cols, y, weights, start, algorithm="lm", control=None, trace=False, devices=(0,),
                    want_resid_grad=False, weights_mode="consistent"):
    """gslnls_fit_large_multi(): one call, host arrays in, the rows split over `devices` inside the library
    (one host thread and one PCIe link per GPU, packets over NVLink peer memory)"""
    ctrl = gsl_nls_control() if control is None else control
    ci, cd = pack_control(ctrl, algorithm, trace)
    st = np.ascontiguousarray(start, dtype=np.float64)
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    arr = (_lib.c_double_p * max(len(cols), 1))(*[_dptr(c) for c in cols])
    dev = np.ascontiguousarray(list(devices), dtype=np.int32)
    res = _lib.Result()
    _lib.check(_lib.lib().gslnls_set_weights_mode(WEIGHTS_MODES[weights_mode]))  # one-shot calls use the default
    try:
        rc = _lib.lib().gslnls_fit_large_multi(model.handle, arr, _dptr(y), _dptr(w) if w is not None else None,
                                               y.size, _dptr(st), ci.ctypes.data_as(_lib.c_int_p), _dptr(cd),
                                               dev.size, dev.ctypes.data_as(_lib.c_int_p), int(want_resid_grad),
                                               C.byref(res))
    finally:
        _lib.lib().gslnls_set_weights_mode(0)
    _lib.check(rc)
    out = _result_to_dict(res, st.size, y.size, int(ci[0]), trace, want_resid_grad)
    _lib.lib().gslnls_result_free(C.byref(res))
    return out


def gsl_nls_large(fn, data=None, start=None, algorithm="lm", control=None, jac=None, fvv=None, trace=False,
                  weights=None, y=None, device=0, comm=None, model=None, devices=None, weights_mode="consistent",
                  loss="default", **kwargs):
    """Fit a nonlinear least-squares model with the large-problem trust-region path on a B200.

    fn        two-sided formula text "y ~ A * exp(-lam * x) + b" (formula method, R/nls_large.R:135),
              or a right-hand-side expression with the responses in `y` (function method, :459)
    data      mapping name -> array with the predictors (and the response for a formula)
    start     mapping / sequence of (name, value): starting values, in parameter order
    algorithm "lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"
    jac       True: symbolic Jacobian as deriv() would give; "forward"/"center": finite differences
              with the step rule of src/fdjac.c.  Required, as in the reference (R/nls_large.R:319-321)
    fvv       for "lmaccel": True symbolic, "fd" finite difference (src/fdfvv.c); required (:354-356)
    loss      "default" (least squares) or a robust loss -- name or gsl_nls_loss(...) -- fitted by IRLS as
              gsl_nls(loss = ) does (src/nls_irls.c:412-546): huber, barron, bisquare, welsh, optimal, hampel, ggw, lqq
    weights_mode  "consistent" (default): g = J^T W f, J^T W J; "gsl": exactly the reference's numbers for
              non-unit weights (sqrt(w) on f only, unweighted J^T J -- what R/nls_large.R:587-600 + libgsl compute)
    """
    if weights_mode not in WEIGHTS_MODES:
        raise ValueError("'weights_mode' should be one of \"consistent\", \"gsl\"")
    if algorithm not in ALGORITHMS:
        raise ValueError("'arg' should be one of %s" % ", ".join('"%s"' % a for a in ALGORITHMS))
    if data is None:
        data = {}
    if not isinstance(data, dict):
        raise TypeError("'data' must be a list or an environment")  # R/nls_large.R:144-145
    if start is None:
        raise ValueError("starting values need to be provided")       # :468-470 (no selfStart on the device)
    start = dict(start)
    pnames = list(start)
    # start values given as ranges (lower, upper) or missing (None / NaN): multi-start, as gsl_nls() does for
    # `start` lists with length-2 elements / NA (R/nls.R:398-424); missing sides default to (-0.1, 0.75)
    ranges, has_range = None, None
    if any(v is None or isinstance(v, (tuple, list)) or (isinstance(v, float) and v != v) for v in start.values()):
        ranges, has_range = [], []
        for v in start.values():
            lo, hi = (v if isinstance(v, (tuple, list)) else (v, v))
            miss = [b is None or b != b or math.isinf(b) for b in (lo, hi)]
            ranges.append([-0.1 if miss[0] else float(lo), 0.75 if miss[1] else float(hi)])
            has_range.append([not miss[0], not miss[1]])
        start = {k: r[0] for k, r in zip(pnames, ranges)}
    two_sided = "~" in fn
    lhs_txt, rhs = (s.strip() for s in fn.split("~", 1)) if two_sided else (None, fn.strip())
    if two_sided and not lhs_txt:
        lhs_txt = None
    names = all_vars(rhs)
    var_names = [v for v in names if v not in pnames]
    missing = [v for v in var_names if v not in data]
    if missing:
        raise ValueError("parameters without starting value in 'data': %s" % ", ".join(missing))  # :210-211
    if lhs_txt is not None:
        lhs = _eval_lhs(lhs_txt, data)
    elif y is not None:
        lhs = np.asarray(y, dtype=np.float64)
        if lhs.ndim != 1:
            raise ValueError("'y' should be a numeric response vector")  # :473-474
    else:
        lhs = None
    cols = [np.ascontiguousarray(data[v], dtype=np.float64) for v in var_names]
    n = cols[0].size if cols else (lhs.size if lhs is not None else 0)
    if lhs is None:
        lhs = np.zeros(n)  # one-sided formula: response 0 (R/nls_large.R:150-153)
    if any(c.size != lhs.size for c in cols):
        raise ValueError("variable lengths differ")
    if not cols and not var_names:
        if lhs.size == 0:
            raise ValueError("no parameters to fit and/or no data variables present")  # :214-216
    if comm is None and lhs.size < len(pnames):
        # a rank of a sharded fit may hold fewer rows than parameters: the library checks the global count
        raise ValueError("negative residual degrees of freedom, cannot fit a model with less observations "
                         "than parameters")  # :286-288
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        if weights.size != lhs.size:
            raise ValueError("'weights' should be numeric equal in length to 'y'")  # :589-590
        if np.any(~(weights > 0)):
            raise ValueError("missing or non-positive weights not allowed")  # :258-259
    if jac is None or jac is False:
        raise ValueError("analytic Jacobian function 'jac' is required, but none is available")  # :319-321
    if jac not in JAC_MODES:
        raise ValueError("'jac' must be True, \"forward\" or \"center\"")
    if algorithm == "lmaccel":
        if fvv is None or fvv is False:
            raise ValueError("analytic second derivative function 'fvv' is required, but none is available")
        if fvv not in FVV_MODES:
            raise ValueError("'fvv' must be True or \"fd\"")
    else:
        fvv = None
    # control (R/nls_large.R:361-382): defaults, minus mstart*, merged with the user's list
    ctrl = gsl_nls_control()
    if control is not None:
        ctrl.update(dict(control))
    ctrl = gsl_nls_control(**ctrl)
    ctrl["solver"] = "cholesky"  # fixed (:370)
    st = np.array([float(start[k]) for k in pnames], dtype=np.float64)

    mdl = model if model is not None else Model(rhs, pnames, var_names, jac=jac, fvv=fvv)
    if devices is not None and len(devices) > 1:
        # several GPUs from this one process: the library splits the rows and runs one thread per GPU; the
        # session stays behind the fitted object, residuals / gradient are lazy (no O(n) arrays from the fit)
        ses = Session(mdl, lhs.size, weights is not None, devices).set_weights_mode(weights_mode)
        ses.upload(cols, lhs, weights)
        cfit = ses.fit(st, algorithm=algorithm, control=ctrl, trace=bool(trace))
        return GslNls(fn, pnames, cfit, ses, mdl, ctrl, algorithm, weights, lhs, bool(trace))
    lossd = gsl_nls_loss(loss) if isinstance(loss, str) else dict(loss)
    robust = lossd["rho"] != "default"
    pb = Problem(mdl, lhs.size, weights is not None or robust, device)
    pb.set_weights_mode(weights_mode)
    pb.upload(cols, lhs, weights if (weights is not None or not robust) else np.ones(lhs.size))
    if comm is not None:
        pb.set_comm(comm)
    ms = None
    if ranges is not None:
        ms = pb.multistart(ranges, has_range, algorithm=algorithm, control=ctrl)
        st = ms["par"]                                          # src/nls.c:534-541: the fit restarts from mpopt
    irls = None
    if robust:
        cfit, irls = pb.fit_irls(st, loss=lossd["rho"], cc=lossd["cc"], algorithm=algorithm, control=ctrl)
    else:
        cfit = pb.fit(st, algorithm=algorithm, control=ctrl, trace=bool(trace))
    obj = GslNls(fn, pnames, cfit, pb, mdl, ctrl, algorithm, weights, lhs, bool(trace))
    obj.mstart = ms
    obj.irls = irls
    return obj
