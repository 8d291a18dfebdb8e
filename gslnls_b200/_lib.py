"""ctypes binding of gslnls_b200/csrc/libgslnls_b200.so (C ABI: include/gslnls_b200.h).

There is no Python or CPU fallback: if the shared library is missing this module raises at
import of the symbols, and every compute entry point fails when no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgslnls_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Result(C.Structure):
    """struct gslnls_result"""
    _fields_ = [
        ("n", C.c_int64), ("p", C.c_int), ("par", c_double_p), ("covar", c_double_p), ("ssr", C.c_double),
        ("ssrtol", C.c_double), ("chisq_init", C.c_double), ("niter", C.c_int), ("conv", C.c_int),
        ("info", C.c_int), ("status", C.c_char_p), ("algorithm", C.c_char_p), ("neval", C.c_int64 * 4),
        ("npass", C.c_int64), ("ntrace", C.c_int), ("partrace", c_double_p), ("ssrtrace", c_double_p),
        ("condtrace", c_double_p), ("resid", c_double_p), ("grad", c_double_p), ("n_local", C.c_int64),
        ("jtj", c_double_p), ("grad_vec", c_double_p), ("x_final", c_double_p),
    ]


class IrlsInfo(C.Structure):
    """struct gslnls_irls_info"""
    _fields_ = [("sigma", C.c_double), ("delta", C.c_double), ("niter", C.c_int), ("status", C.c_int)]


class MstartResult(C.Structure):
    """struct gslnls_mstart_result"""
    _fields_ = [("p", C.c_int), ("par", c_double_p), ("range", c_double_p), ("ssr", C.c_double),
                ("ssrconv", C.c_double), ("nsp", C.c_int), ("nwsp", C.c_int), ("mstarts", C.c_int),
                ("status", C.c_int), ("searches", C.c_int64)]


class SparseResult(C.Structure):
    """struct gslnls_sparse_result"""
    _fields_ = [("p", C.c_int), ("nrows", C.c_int64), ("nterms", C.c_int64), ("nnz", C.c_int64), ("par", c_double_p),
                ("ssr", C.c_double), ("ssrtol", C.c_double), ("chisq_init", C.c_double), ("niter", C.c_int),
                ("conv", C.c_int), ("info", C.c_int), ("status", C.c_char_p), ("neval", C.c_int64 * 4),
                ("cg_iters", C.c_int64), ("launches", C.c_int64), ("eval_ms", C.c_double), ("solver_ms", C.c_double),
                ("ntrace", C.c_int), ("ssrtrace", c_double_p),
                ("grad_vec", c_double_p), ("jtj", c_double_p), ("resid", c_double_p)]


# every symbol include/gslnls_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gslnls_model_compile": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                       C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]),
    "gslnls_model_free": (None, [C.c_void_p]),
    "gslnls_model_p": (C.c_int, [C.c_void_p]),
    "gslnls_model_nvar": (C.c_int, [C.c_void_p]),
    "gslnls_model_source": (C.c_char_p, [C.c_void_p]),
    "gslnls_fit_large": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_double_p, c_double_p, C.c_int64, c_double_p,
                                   c_int_p, c_double_p, C.c_int, C.c_int, C.POINTER(Result)]),
    "gslnls_fit_large_sharded": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_double_p, c_double_p, C.c_int64,
                                           c_double_p, c_int_p, c_double_p, C.c_int, C.c_void_p, C.c_int,
                                           C.POINTER(Result)]),
    "gslnls_fit_large_multi": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_double_p, c_double_p, C.c_int64,
                                         c_double_p, c_int_p, c_double_p, C.c_int, c_int_p, C.c_int,
                                         C.POINTER(Result)]),
    "gslnls_result_free": (None, [C.POINTER(Result)]),
    "gslnls_session_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_int_p, C.POINTER(C.c_void_p)]),
    "gslnls_session_free": (None, [C.c_void_p]),
    "gslnls_session_ngpu": (C.c_int, [C.c_void_p]),
    "gslnls_session_set_weights_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "gslnls_session_upload": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_double_p, c_double_p]),
    "gslnls_session_fit": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, C.c_int, C.POINTER(Result)]),
    "gslnls_session_residuals": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "gslnls_cache_clear": (None, []),
    "gslnls_problem_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "gslnls_problem_free": (None, [C.c_void_p]),
    "gslnls_problem_upload": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_double_p, c_double_p]),
    "gslnls_problem_bind_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "gslnls_problem_set_weights_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "gslnls_set_weights_mode": (C.c_int, [C.c_int]),
    "gslnls_problem_set_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gslnls_problem_fit": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, C.c_int, C.POINTER(Result)]),
    "gslnls_problem_eval_packet": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "gslnls_problem_eval_jtfvv": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "gslnls_problem_time_passes": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.POINTER(C.c_float)]),
    "gslnls_problem_residuals": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "gslnls_problem_fit_begin": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p]),
    "gslnls_problem_fit_run": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "gslnls_problem_fit_end": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Result)]),
    "gslnls_problem_launch_count": (C.c_int64, [C.c_void_p]),
    "gslnls_problem_trace": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_int, c_int_p]),
    "gslnls_problem_timer_start": (C.c_int, [C.c_void_p]),
    "gslnls_problem_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "gslnls_problem_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "gslnls_problem_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "gslnls_problem_channel_stats": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.POINTER(C.c_int64)]),
    "gslnls_problem_fit_batch": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_int_p, c_double_p, c_double_p,
                                           c_double_p, c_double_p, c_int_p, c_int_p]),
    "gslnls_problem_fit_irls": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, C.c_int, c_double_p, C.c_int,
                                          C.c_double, C.POINTER(Result), C.POINTER(IrlsInfo)]),
    "gslnls_problem_get_weights": (C.c_int, [C.c_void_p, c_double_p]),
    "gslnls_problem_median_abs_resid": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "gslnls_problem_multistart": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_int_p, c_double_p, c_int_p, c_double_p,
                                            C.POINTER(MstartResult)]),
    "gslnls_mstart_result_free": (None, [C.POINTER(MstartResult)]),
    "gslnls_qrng_points": (C.c_int, [C.c_int, C.c_int, c_double_p]),
    "gslnls_sparse_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "gslnls_sparse_free": (None, [C.c_void_p]),
    "gslnls_sparse_add_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(c_double_p), c_int_p,
                                          C.POINTER(c_int_p), c_int_p, C.c_int64]),
    "gslnls_sparse_set_response": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "gslnls_sparse_finalize": (C.c_int, [C.c_void_p]),
    "gslnls_sparse_nnz": (C.c_int64, [C.c_void_p]),
    "gslnls_sparse_fit": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_double_p, C.c_int, C.c_int,
                                    C.POINTER(SparseResult)]),
    "gslnls_sparse_eval": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "gslnls_sparse_result_free": (None, [C.POINTER(SparseResult)]),
    "gslnls_comm_get_unique_id": (C.c_int, [C.c_void_p]),
    "gslnls_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "gslnls_comm_create_local": (C.c_int, [C.c_int, c_int_p, C.POINTER(C.c_void_p)]),
    "gslnls_comm_free": (None, [C.c_void_p]),
    "gslnls_comm_has_peer_memory": (C.c_int, [C.c_void_p]),
    "gslnls_comm_rank": (C.c_int, [C.c_void_p]),
    "gslnls_comm_size": (C.c_int, [C.c_void_p]),
    "gslnls_measure_fp64_peak": (C.c_int, [C.c_int, c_double_p, c_double_p]),
    "gslnls_measure_read_bandwidth": (C.c_int, [C.c_int, C.c_size_t, c_double_p]),
    "gslnls_strerror": (C.c_char_p, [C.c_int]),
    "gslnls_trs_name": (C.c_char_p, [C.c_int]),
    "gslnls_last_error": (C.c_char_p, []),
    "gslnls_device_count": (C.c_int, []),
    "gslnls_version": (C.c_char_p, []),
}

_LIB = None


class GslnlsError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().gslnls_strerror(code).decode()
        super().__init__("%s (code %d)%s" % (msg, code, ": " + detail if detail else ""))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gslnls_b200/csrc`). gslnls_b200 has no fallback path." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(code):
    """raise on library-level errors (>= 1000) and invalid arguments; GSL statuses pass through"""
    if code >= 1000 or code in (4,):
        raise GslnlsError(code, lib().gslnls_last_error().decode())
    return code
