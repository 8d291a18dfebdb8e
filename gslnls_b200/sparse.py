"""Sparse-row problems: the host-side mirror of gsl_nls_large() with a sparse Jacobian (R/nls_large.R:397-404,
src/nls_large.c:528-648) on top of the gslnls_sparse_* C ABI (include/gslnls_b200.h).

The reference receives the sparsity from an R closure that returns a dgC/dgR/dgTMatrix; here it is data.  A
problem has `p` parameters and `nrows` residual rows and is made of blocks:

    sp = SparseProblem(p=G + 2, nrows=n)
    sp.add_block("A * exp(-lam * x)", {"A": (0, g), "lam": G}, {"x": x})    # A = theta[0 + g[t]], lam = theta[G]
    sp.set_response(y)
    fit = sp.fit(start)                                                       # algorithm "cgst", matrix-free

A parameter binding is an int (fixed global index) or (base, index_column).  Rows default to one term per row in
the order blocks are added; `rows=` assigns terms to rows explicitly and a row is the SUM of its terms.
"""
import ctypes as C

import numpy as np

from . import _lib
from .control import gsl_nls_control, pack_control
from .nls_large import Model


class SparseProblem:
    def __init__(self, p, nrows, device=0):
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.gslnls_sparse_create(int(device), int(p), int(nrows), C.byref(h)))
        self.handle, self.p, self.nrows = h, int(p), int(nrows)
        self._models, self._next_row, self._final = [], 0, False

    def add_block(self, rhs, params, data=None, rows=None, nterms=None):
        """rhs: R-style expression; params: {name: index | (base, int column)} in the order of appearance in the
        dict; data: {name: float column}; rows: int column or None (consecutive rows after the previous block)"""
        L = _lib.lib()
        data = data or {}
        names = list(params)
        m = Model(rhs, names, list(data), jac="symbolic")
        cols = [np.ascontiguousarray(v, dtype=np.float64) for v in data.values()]
        base = np.zeros(len(names), dtype=np.int32)
        idx = []
        for s, k in enumerate(names):
            b = params[k]
            if isinstance(b, tuple):
                base[s] = int(b[0])
                idx.append(np.ascontiguousarray(b[1], dtype=np.int32))
            else:
                base[s] = int(b)
                idx.append(None)
        sizes = {c.size for c in cols} | {i.size for i in idx if i is not None}
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.int32)
            sizes.add(rows.size)
        if nterms is not None:
            sizes.add(int(nterms))
        if len(sizes) != 1:
            raise ValueError("variable lengths differ" if sizes else "the number of terms of the block is not known")
        nt = sizes.pop()
        vp = (_lib.c_double_p * max(1, len(cols)))(*[c.ctypes.data_as(_lib.c_double_p) for c in cols])
        ip = (_lib.c_int_p * len(names))(*[i.ctypes.data_as(_lib.c_int_p) if i is not None else None for i in idx])
        _lib.check(L.gslnls_sparse_add_block(self.handle, m.handle, nt, vp, base.ctypes.data_as(_lib.c_int_p), ip,
                                             rows.ctypes.data_as(_lib.c_int_p) if rows is not None else None,
                                             self._next_row))
        if rows is None:
            self._next_row += nt
        self._models.append(m)  # the problem uses the model's kernels: keep it alive
        return self

    def set_response(self, y=None, weights=None):
        L = _lib.lib()
        ya = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        wa = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        for a in (ya, wa):
            if a is not None and a.size != self.nrows:
                raise ValueError("'y' / 'weights' should have one entry per row")
        _lib.check(L.gslnls_sparse_set_response(self.handle,
                                                ya.ctypes.data_as(_lib.c_double_p) if ya is not None else None,
                                                wa.ctypes.data_as(_lib.c_double_p) if wa is not None else None))
        return self

    def finalize(self):
        if not self._final:
            _lib.check(_lib.lib().gslnls_sparse_finalize(self.handle))
            self._final = True
        return self

    @property
    def nnz(self):
        return int(_lib.lib().gslnls_sparse_nnz(self.handle))

    def eval(self, theta):
        """weighted residuals, J^T f, diag(J^T J), f^T f at theta"""
        self.finalize()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        r, g, d = np.empty(self.nrows), np.empty(self.p), np.empty(self.p)
        ssr = C.c_double()
        rc = _lib.lib().gslnls_sparse_eval(self.handle, th.ctypes.data_as(_lib.c_double_p),
                                           r.ctypes.data_as(_lib.c_double_p), g.ctypes.data_as(_lib.c_double_p),
                                           d.ctypes.data_as(_lib.c_double_p), C.byref(ssr))
        _lib.check(rc)
        return {"resid": r, "grad_vec": g, "jtj_diag": d, "ssr": ssr.value, "status": rc}

    def fit(self, start, algorithm="cgst", control=None, trace=False, want_jtj=False, want_resid=False):
        self.finalize()
        ctrl = gsl_nls_control(**(control or {}))
        ci, cd = pack_control(ctrl, algorithm, trace)
        st = np.ascontiguousarray(start, dtype=np.float64)
        if st.size != self.p:
            raise ValueError("'start' should have one value per parameter")
        res = _lib.SparseResult()
        rc = _lib.lib().gslnls_sparse_fit(self.handle, st.ctypes.data_as(_lib.c_double_p),
                                          ci.ctypes.data_as(_lib.c_int_p), cd.ctypes.data_as(_lib.c_double_p),
                                          int(want_jtj), int(want_resid), C.byref(res))
        _lib.check(rc)
        p = self.p
        out = {
            "par": np.ctypeslib.as_array(res.par, shape=(p,)).copy(),
            "grad_vec": np.ctypeslib.as_array(res.grad_vec, shape=(p,)).copy(),
            "ssr": res.ssr, "ssrtol": res.ssrtol, "chisq_init": res.chisq_init, "niter": res.niter,
            "conv": res.conv, "info": res.info, "status": res.status.decode(),
            "neval": {"f": res.neval[0], "dfu": res.neval[1], "df2": res.neval[2], "fvv": res.neval[3]},
            "cg_iters": res.cg_iters, "launches": res.launches, "eval_ms": res.eval_ms, "solver_ms": res.solver_ms, "nnz": res.nnz, "nterms": res.nterms,
            "algorithm": algorithm,
        }
        if trace:
            out["ssrtrace"] = np.ctypeslib.as_array(res.ssrtrace, shape=(res.ntrace,)).copy()[: res.niter + 1]
        if want_jtj:
            out["jtj"] = np.ctypeslib.as_array(res.jtj, shape=(p, p)).copy().T
        if want_resid:
            out["resid"] = np.ctypeslib.as_array(res.resid, shape=(self.nrows,)).copy()
        _lib.lib().gslnls_sparse_result_free(C.byref(res))
        return out

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().gslnls_sparse_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
