"""Binding of pu_refmath_eval (include/pu/pu_capi.h): the libm restatements of csrc/ref_math.cuh evaluated on the
host (ctx=None) or on the device."""
import ctypes as C

import numpy as np

from . import capi

OPS = {"atan2f": 0, "sinf": 1, "cosf": 2, "hypotf": 3, "atanf": 4}


def evaluate(op, a, b=None, ctx=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros_like(a)
    capi.check(capi.lib().pu_refmath_eval(ctx._h if ctx is not None else None, OPS[op], capi._ptr(a), capi._ptr(bb),
                                          capi._ptr(out), C.c_size_t(a.size)))
    return out
