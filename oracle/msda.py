"""ORACLE (test infrastructure): numpy front-end of oracle/msda_ref.c.

Restates MSDA.ms_deform_attn_forward / _backward of the reference
(<proj>/models/model_utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, 87-159, 301-403).
Pinned by tests/golden/msda_*.npz, generated from the reference's own
ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:41-61).
"""
import ctypes

import numpy as np

from . import lib

_I64 = ctypes.c_int64


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = value.dtype
    assert dt in (np.float32, np.float64)
    value = np.ascontiguousarray(value)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2) and attn.shape == (N, Lq, M, L, P)
    return value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P)


def msda_forward(value, shapes, lsi, loc, attn):
    value, shapes, lsi, loc, attn, dims = _prep(value, shapes, lsi, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    fn = getattr(lib(), "oracle_msda_forward_" + ("f32" if value.dtype == np.float32 else "f64"))
    fn.restype = None
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(out), *[_I64(d) for d in dims])
    return out


def msda_backward(value, shapes, lsi, loc, attn, grad_out):
    value, shapes, lsi, loc, attn, dims = _prep(value, shapes, lsi, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = getattr(lib(), "oracle_msda_backward_" + ("f32" if value.dtype == np.float32 else "f64"))
    fn.restype = None
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(grad_out), _p(gv), _p(gl), _p(ga),
       *[_I64(d) for d in dims])
    return gv, gl, ga
