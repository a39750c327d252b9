"""ORACLE (test infrastructure): numpy front-end of oracle/spconv_ref.c — the reference's vendored
spconv v1 CPU algorithms (rulebook: include/spconv/geometry.h:24-297; conv fwd/bwd:
include/spconv/spconv_ops.h:260-456), plus the canonical relabelling used to compare orders."""
import ctypes

import numpy as np

from . import lib

_I64 = ctypes.c_int64


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """ops.py:20-31."""
    return [(input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
            for i in range(len(input_size))]


def get_indice_pairs(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm,
                     order="cpu"):
    """Returns (outids, indice_pairs [K,2,N], indice_num [K]) like ops.get_indice_pairs.

    order="cpu": reference CPU path order (outputs first-touch, pairs ascending input row).
    order="gpu": outputs relabelled to sorted flat (b,z,y,x) index (reference GPU path,
                 spconv_ops.h:129-137); pair slots still ascending in the input row (canonical).
    """
    indices = _i32(indices)
    n = indices.shape[0]
    ksize, stride, padding, dilation = map(list, (ksize, stride, padding, dilation))
    out_shape = list(spatial_shape) if subm else get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    kvol = int(np.prod(ksize))
    pairs = np.full((kvol, 2, n), -1, np.int32)
    num = np.zeros((kvol,), np.int32)
    outids = np.zeros((max(n * kvol, 1), 4), np.int32)
    fn = lib().oracle_get_indice_pairs
    fn.restype = _I64
    a = [_i32(out_shape), _i32(ksize), _i32(stride), _i32(padding), _i32(dilation)]
    m = fn(_p(indices), _I64(n), *[_p(x) for x in a], ctypes.c_int(int(subm)), _p(outids), _p(pairs), _p(num))
    if subm:
        return indices, pairs, num, out_shape
    outids = outids[:m].copy()
    if order == "gpu" and m > 0:
        flat = outids[:, 0].astype(np.int64)
        for d in range(3):
            flat = flat * out_shape[d] + outids[:, 1 + d]
        perm = np.argsort(flat, kind="stable")      # new row r holds old row perm[r]
        inv = np.empty_like(perm)
        inv[perm] = np.arange(m)
        outids = outids[perm]
        for k in range(kvol):
            pairs[k, 1, :num[k]] = inv[pairs[k, 1, :num[k]]]
    return outids, pairs, num, out_shape


def indice_conv(features, filters, pairs, num, n_out, inverse=False):
    features = np.ascontiguousarray(features, np.float32)
    kvol = pairs.shape[0]
    cin, cout = filters.shape[-2], filters.shape[-1]
    filters = np.ascontiguousarray(filters, np.float32).reshape(kvol, cin, cout)
    pairs, num = _i32(pairs), _i32(num)
    out = np.empty((n_out, cout), np.float32)
    fn = lib().oracle_indice_conv_f32
    fn.restype = None
    fn(_p(features), _p(filters), _p(pairs), _p(num), _I64(pairs.shape[2]), _I64(kvol), _I64(cin),
       _I64(cout), _I64(n_out), ctypes.c_int(int(inverse)), _p(out))
    return out


def indice_conv_backward(features, filters, out_grad, pairs, num, inverse=False):
    features = np.ascontiguousarray(features, np.float32)
    out_grad = np.ascontiguousarray(out_grad, np.float32)
    kvol = pairs.shape[0]
    cin, cout = filters.shape[-2], filters.shape[-1]
    fshape = filters.shape
    filters = np.ascontiguousarray(filters, np.float32).reshape(kvol, cin, cout)
    pairs, num = _i32(pairs), _i32(num)
    gin = np.empty_like(features)
    gw = np.empty_like(filters)
    fn = lib().oracle_indice_conv_backward_f32
    fn.restype = None
    fn(_p(features), _p(filters), _p(out_grad), _p(pairs), _p(num), _I64(pairs.shape[2]), _I64(kvol),
       _I64(cin), _I64(cout), _I64(features.shape[0]), ctypes.c_int(int(inverse)), _p(gin), _p(gw))
    return gin, gw.reshape(fshape)


def dense(features, indices, spatial_shape, batch_size):
    """SparseConvTensor.dense() (structure.py:55-64): [B, C, D, H, W]."""
    C = features.shape[1]
    out = np.zeros((batch_size, *spatial_shape, C), features.dtype)
    out[indices[:, 0], indices[:, 1], indices[:, 2], indices[:, 3]] = features
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))
