"""gsl_nls_control() -- the 23 tunables of the reference (R/nls.R:1186-1229), same names and defaults."""
import math
import sys

SQRT_EPS = math.sqrt(sys.float_info.epsilon)

ALGORITHMS = ("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst")
SCALES = ("more", "levenberg", "marquardt")
SOLVERS = ("qr", "cholesky", "svd")
FDTYPES = ("forward", "center")


def _match(value, choices, what):
    if value not in choices:
        raise ValueError("'%s' should be one of %s" % (what, ", ".join('"%s"' % c for c in choices)))
    return value


def gsl_nls_control(maxiter=100, scale="more", solver="qr", fdtype="forward", factor_up=2, factor_down=3,
                    avmax=0.75, h_df=SQRT_EPS, h_fvv=0.02, xtol=SQRT_EPS, ftol=SQRT_EPS, gtol=SQRT_EPS,
                    mstart_n=30, mstart_p=5, mstart_q=None, mstart_r=4, mstart_s=2, mstart_tol=0.25,
                    mstart_maxiter=10, mstart_maxstart=250, mstart_minsp=1, irls_maxiter=50,
                    irls_xtol=sys.float_info.epsilon ** 0.25, **_ignored):
    """Tunable parameters; validation mirrors the stopifnot() block at R/nls.R:1198-1219."""
    scale = _match(scale, SCALES, "scale")
    solver = _match(solver, SOLVERS, "solver")
    fdtype = _match(fdtype, FDTYPES, "fdtype")
    if mstart_q is None:
        mstart_q = mstart_n // 10
    num = dict(maxiter=maxiter, factor_up=factor_up, factor_down=factor_down, avmax=avmax, h_df=h_df, h_fvv=h_fvv,
               xtol=xtol, ftol=ftol, gtol=gtol, mstart_n=mstart_n, mstart_p=mstart_p, mstart_q=mstart_q,
               mstart_r=mstart_r, mstart_s=mstart_s, mstart_tol=mstart_tol, mstart_maxiter=mstart_maxiter,
               mstart_maxstart=mstart_maxstart, mstart_minsp=mstart_minsp, irls_maxiter=irls_maxiter,
               irls_xtol=irls_xtol)
    for k, v in num.items():
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise ValueError("'%s' should be a numeric scalar" % k)
    ge1 = ("maxiter", "mstart_n", "mstart_p", "mstart_q", "mstart_s", "mstart_maxiter", "mstart_maxstart",
           "mstart_minsp", "irls_maxiter")
    for k in ge1:
        if not num[k] >= 1:
            raise ValueError("%s >= 1 is not TRUE" % k)
    for k in ("factor_up", "factor_down", "avmax", "h_df", "h_fvv", "xtol", "ftol", "gtol", "mstart_tol", "irls_xtol"):
        if not num[k] > 0:
            raise ValueError("%s > 0 is not TRUE" % k)
    if not mstart_r > 1:
        raise ValueError("mstart_r > 1 is not TRUE")
    return dict(maxiter=int(maxiter), scale=scale, solver=solver, fdtype=fdtype, factor_up=factor_up,
                factor_down=factor_down, avmax=avmax, h_df=h_df, h_fvv=h_fvv, xtol=xtol, ftol=ftol, gtol=gtol,
                mstart_n=int(mstart_n), mstart_p=int(mstart_p), mstart_q=int(mstart_q), mstart_r=mstart_r,
                mstart_s=int(mstart_s), mstart_tol=mstart_tol, mstart_maxiter=int(mstart_maxiter),
                mstart_maxstart=int(mstart_maxstart), mstart_minsp=int(mstart_minsp),
                irls_maxiter=int(irls_maxiter), irls_xtol=irls_xtol)


def pack_control(ctrl, algorithm, trace):
    """.ctrl_int (7) and .ctrl_dbl (8) exactly as R/nls_large.R:383-407 packs them."""
    import numpy as np
    ci = np.array([int(ctrl["maxiter"]), int(bool(trace)), ALGORITHMS.index(algorithm), SCALES.index(ctrl["scale"]),
                   FDTYPES.index(ctrl["fdtype"]), -2, 0], dtype=np.int32)
    cd = np.array([ctrl[k] for k in ("factor_up", "factor_down", "avmax", "h_df", "h_fvv", "xtol", "ftol", "gtol")],
                  dtype=np.float64)
    return ci, cd
