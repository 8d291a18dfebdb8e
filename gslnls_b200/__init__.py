"""gslnls_b200 -- B200-native gsl_nls_large() hot path (host-side mirror of the reference's R API).

Public names follow the reference package: gsl_nls_large(), gsl_nls_control().  Everything O(n)
runs in gslnls_b200/csrc/libgslnls_b200.so (hand-written sm_100a CUDA behind a C ABI, see
include/gslnls_b200.h); importing this package without that library raises.
"""
from .control import gsl_nls_control, pack_control  # noqa: F401
from .nls_large import (GslNls, Model, Problem, Session, fit_large_multi, gsl_nls_large,  # noqa: F401
                        gsl_nls_loss)
from .sparse import SparseProblem  # noqa: F401

__all__ = ["gsl_nls_large", "gsl_nls_control", "GslNls", "Model", "Problem", "pack_control", "fit_large_multi",
           "Session", "gsl_nls_loss", "SparseProblem"]
