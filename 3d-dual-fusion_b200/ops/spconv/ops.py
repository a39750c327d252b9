"""Rulebook / convolution op wrappers with the reference's names and argument meaning
(TransFusion/mmdet3d/ops/spconv/ops.py:20-183) on top of the C-ABI (include/ddf_b200.h)."""
import ctypes

import torch

from ... import lib as _lib


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        if kernel_size[i] == -1:
            output_size.append(1)
            continue
        output_size.append((input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1)
                           // stride[i] + 1)
    return output_size


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation, output_padding):
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        if kernel_size[i] == -1:
            raise ValueError("deconv don't support kernel_size < 0")
        output_size.append((input_size[i] - 1) * stride[i] - 2 * padding[i] + kernel_size[i]
                           + output_padding[i])
    return output_size


def _i64x3(v):
    return (ctypes.c_int64 * 3)(*[int(x) for x in v])


def _listify(v, ndim):
    return list(v) if isinstance(v, (list, tuple)) else [v] * ndim


class Rulebook(object):
    """Reference-format rulebook + the row-major tables of the fused kernels. For SubM layers whose
    three conv kernels all walk the tables the pair lists are built only when somebody reads them."""
    __slots__ = ("outids", "_pairs", "_num", "gather_table", "scatter_table", "out_spatial_shape", "subm",
                 "kvol", "_build_pairs", "key_indices")

    def __init__(self, outids, indice_pairs, indice_pair_num, gather_table, scatter_table,
                 out_spatial_shape, kvol=None, build_pairs=None):
        self.outids = outids
        self._pairs = indice_pairs
        self._num = indice_pair_num
        self.gather_table = gather_table
        self.scatter_table = scatter_table
        self.out_spatial_shape = out_spatial_shape
        self.subm = False
        self.kvol = kvol if kvol is not None else indice_pairs.shape[0]
        self._build_pairs = build_pairs
        self.key_indices = None     # input indices the cache key was derived from (kept alive by SparseConvolution)

    def _ensure_pairs(self):
        if self._pairs is None:
            self._pairs, self._num = self._build_pairs()
            self._build_pairs = None

    @property
    def indice_pairs(self):
        self._ensure_pairs()
        return self._pairs

    @property
    def indice_pair_num(self):
        self._ensure_pairs()
        return self._num


def build_rulebook(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                   out_padding=0, subm=False, transpose=False, with_tables=True, with_pairs=True):
    """Build a rulebook on the device. One host read (num_act_out) for regular convs, none for SubM.
    ``with_pairs=False`` (SubM with tables): the reference-format pair lists are built lazily."""
    _lib.require_cuda(indices)
    if indices.dtype != torch.int32 or not indices.is_contiguous():
        raise RuntimeError("indices must be a contiguous int32 [N, 4] tensor")
    ndim = indices.shape[1] - 1
    if ndim != 3:
        raise NotImplementedError("only 3-D rulebooks are implemented")
    if transpose:
        raise NotImplementedError("transposed rulebooks are not implemented (unused by 3D-DF)")
    ksize, stride, padding = _listify(ksize, ndim), _listify(stride, ndim), _listify(padding, ndim)
    dilation = _listify(dilation, ndim)
    for d, s in zip(dilation, stride):
        assert any([s == 1, d == 1]), "don't support this."
    out_shape = list(spatial_shape) if subm else get_conv_output_size(spatial_shape, ksize, stride,
                                                                      padding, dilation)
    L = _lib.get_lib()
    n = indices.shape[0]
    kvol = ksize[0] * ksize[1] * ksize[2]
    dev = indices.device
    geo = [_i64x3(out_shape), _i64x3(spatial_shape), _i64x3(ksize), _i64x3(stride), _i64x3(padding),
           _i64x3(dilation)]
    ws_bytes = L.ddf_indice_pairs_workspace_bytes(n, batch_size, *geo, int(subm))
    if ws_bytes < 0:
        raise RuntimeError("get_indice_pairs: bad geometry")
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)
    lazy = subm and with_tables and not with_pairs and n > 0
    pairs = None if lazy else torch.empty((kvol, 2, n), dtype=torch.int32, device=dev)
    num = None if lazy else torch.empty((kvol,), dtype=torch.int32, device=dev)
    scatter_t = torch.empty((n, kvol), dtype=torch.int32, device=dev) if with_tables else None
    stream = _lib.current_stream()
    with _lib.on_device(dev):
        if subm:
            gather_t = torch.empty((n, kvol), dtype=torch.int32, device=dev) if with_tables else None
            rc = L.ddf_subm_indice_pairs(_lib.ptr(indices), n, batch_size, geo[1], geo[2], geo[5],
                                         _lib.ptr(pairs), _lib.ptr(num), _lib.ptr(gather_t),
                                         _lib.ptr(scatter_t), _lib.ptr(ws), int(ws_bytes), stream)
            _lib.check(rc, "subm_indice_pairs")

            def build_pairs():
                full = build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation,
                                      out_padding, True, False, with_tables=False)
                return full.indice_pairs, full.indice_pair_num

            rb = Rulebook(indices, pairs, num, gather_t, scatter_t, out_shape, kvol, build_pairs if lazy else None)
            rb.subm = True
            return rb
        cnt = torch.empty(1, dtype=torch.int32, device=dev)
        rc = L.ddf_conv_count_outputs(_lib.ptr(indices), n, batch_size, *geo, _lib.ptr(cnt),
                                      _lib.ptr(ws), int(ws_bytes), stream)
        _lib.check(rc, "conv_count_outputs")
        n_out = int(cnt.item())
        outids = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
        gather_t = torch.empty((n_out, kvol), dtype=torch.int32, device=dev) if with_tables else None
        rc = L.ddf_conv_indice_pairs(_lib.ptr(indices), n, batch_size, *geo, n_out, _lib.ptr(outids),
                                     _lib.ptr(pairs), _lib.ptr(num), _lib.ptr(gather_t),
                                     _lib.ptr(scatter_t), _lib.ptr(ws), int(ws_bytes), stream)
        _lib.check(rc, "conv_indice_pairs")
    return Rulebook(outids, pairs, num, gather_t, scatter_t, out_shape)


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     out_padding=0, subm=False, transpose=False, grid=None):
    """Same contract as ops.py:46-105: returns (outids, indice_pairs, indice_pair_num)."""
    rb = build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation,
                        out_padding, subm, transpose, with_tables=False)
    return rb.outids, rb.indice_pairs, rb.indice_pair_num


def _check_conv_args(features, filters):
    _lib.require_cuda(features, filters)
    if features.dtype != torch.float32 or filters.dtype != torch.float32:
        raise RuntimeError("sparse conv: float32 only")  # reference also exports half; not built here


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse=False,
                subm=False):
    """Same contract as ops.py:108-121 -> sparse_conv_ext.indice_conv_fp32."""
    _check_conv_args(features, filters)
    features, filters = features.contiguous(), filters.contiguous()
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = indice_pairs.shape[0]
    out = torch.empty((num_activate_out, cout), dtype=features.dtype, device=features.device)
    table_ws = torch.empty((max(num_activate_out, 1), kvol), dtype=torch.int32, device=features.device)
    with _lib.on_device(features.device):
        rc = _lib.get_lib().ddf_indice_conv(
            _lib.ptr(features), _lib.ptr(filters), _lib.ptr(indice_pairs), _lib.ptr(indice_pair_num),
            indice_pairs.shape[2], _lib.ptr(out), num_activate_out, kvol, cin, cout, int(inverse),
            int(subm), _lib.ptr(table_ws), _lib.current_stream())
    _lib.check(rc, "indice_conv")
    return out


def indice_conv_backward(features, filters, out_bp, indice_pairs, indice_pair_num, inverse=False,
                         subm=False):
    """Same contract as ops.py:139-152 -> sparse_conv_ext.indice_conv_backward_fp32."""
    _check_conv_args(features, filters)
    features, filters, out_bp = features.contiguous(), filters.contiguous(), out_bp.contiguous()
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = indice_pairs.shape[0]
    n_in = features.shape[0]
    gin = torch.empty_like(features)
    gw = torch.empty_like(filters)
    table_ws = torch.empty((max(n_in, 1), kvol), dtype=torch.int32, device=features.device)
    wt_ws = torch.empty_like(filters)
    with _lib.on_device(features.device):
        rc = _lib.get_lib().ddf_indice_conv_backward(
            _lib.ptr(features), _lib.ptr(filters), _lib.ptr(out_bp), _lib.ptr(indice_pairs),
            _lib.ptr(indice_pair_num), indice_pairs.shape[2], _lib.ptr(gin), _lib.ptr(gw), n_in, kvol,
            cin, cout, int(inverse), int(subm), _lib.ptr(table_ws), _lib.ptr(wt_ws),
            _lib.current_stream())
    _lib.check(rc, "indice_conv_backward")
    return gin, gw


def tc_mode(kvol, cin, cout):
    """Bit mask of the conv kernels of this layer that run on the tensor cores (1 fwd, 2 dgrad, 4 wgrad, 8 table
    wgrad; 16 / 32: fwd / dgrad take bf16x3 split operands)."""
    return int(_lib.get_lib().ddf_sparse_conv_tc_mode(int(kvol), int(cin), int(cout)))


def padded_cin(kvol, cin, cout):
    """Input-channel count the hot-path Function actually runs a layer with: narrow, unaligned
    contractions (e.g. the 5-channel input layer) are zero-padded to a multiple of 8 when that puts
    the layer on the tensor-core kernels."""
    if cin < 32 and cin % 8 and tc_mode(kvol, cin + (-cin) % 8, cout):
        return cin + (-cin) % 8
    return cin


def round_tf32(x):
    """Copy of ``x`` rounded to the nearest tf32 (operand preparation for the tensor-core kernels)."""
    x = x.contiguous()
    out = torch.empty_like(x)
    with _lib.on_device(x.device):
        rc = _lib.get_lib().ddf_round_tf32(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.current_stream())
    _lib.check(rc, "round_tf32")
    return out


def bf16x3_on(C):
    """True when a SubM 3x3x3 conv of C -> C channels runs on the bf16x3 tcgen05 kernels (mode bit 16)."""
    return bool(tc_mode(27, C, C) & 16)


def split_bf16x3(x, want_rounded=False):
    """``x`` fp32 [rows, C] (C % 32 == 0) -> (split, rounded): ``split`` holds every row as blocks of
    [32 x bf16 hi | 32 x bf16 lo] (same bytes as fp32; carried in a float32 tensor of x's shape), the operand
    layout of the bf16x3 conv kernels; ``rounded`` = tf32-rounded copy for the wgrad kernels (or None)."""
    x = x.contiguous()
    split = torch.empty_like(x)
    rounded = torch.empty_like(x) if want_rounded else None
    with _lib.on_device(x.device):
        rc = _lib.get_lib().ddf_split_bf16x3(_lib.ptr(x), _lib.ptr(split), _lib.ptr(rounded), x.shape[0], x.shape[1],
                                             _lib.current_stream())
    _lib.check(rc, "split_bf16x3")
    return split, rounded


def sparse_conv_forward(features, filters, gather_table, bias, n_out, operand_format=0):
    _check_conv_args(features, filters)
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = gather_table.shape[1] if gather_table.numel() else filters.numel() // (cin * cout)
    out = torch.empty((n_out, cout), dtype=features.dtype, device=features.device)
    wt_ws = torch.empty_like(filters)
    with _lib.on_device(features.device):
        rc = _lib.get_lib().ddf_sparse_conv_forward(
            _lib.ptr(features), _lib.ptr(filters), _lib.ptr(gather_table), _lib.ptr(bias), _lib.ptr(out),
            _lib.ptr(wt_ws), n_out, features.shape[0], kvol, cin, cout, int(operand_format),
            _lib.current_stream())
    _lib.check(rc, "sparse_conv_forward")
    return out


def sparse_conv_dgrad(filters, grad_out, scatter_table, n_in, operand_format=0):
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = filters.numel() // (cin * cout)
    gin = torch.empty((n_in, cin), dtype=grad_out.dtype, device=grad_out.device)
    wt_ws = torch.empty_like(filters)
    with _lib.on_device(grad_out.device):
        rc = _lib.get_lib().ddf_sparse_conv_dgrad(_lib.ptr(grad_out), _lib.ptr(filters),
                                                  _lib.ptr(scatter_table), _lib.ptr(gin), _lib.ptr(wt_ws),
                                                  n_in, grad_out.shape[0], kvol, cin, cout,
                                                  int(operand_format), _lib.current_stream())
    _lib.check(rc, "sparse_conv_dgrad")
    return gin


def sparse_conv_wgrad(features, filters, grad_out, indice_pairs, indice_pair_num):
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = indice_pairs.shape[0]
    gw = torch.empty_like(filters)
    with _lib.on_device(features.device):
        rc = _lib.get_lib().ddf_sparse_conv_wgrad(_lib.ptr(features), _lib.ptr(grad_out),
                                                  _lib.ptr(indice_pairs), _lib.ptr(indice_pair_num),
                                                  indice_pairs.shape[2], _lib.ptr(gw), kvol, cin, cout, 0,
                                                  _lib.current_stream())
    _lib.check(rc, "sparse_conv_wgrad")
    return gw


def sparse_conv_wgrad_table(features, filters, grad_out, gather_table):
    """wgrad through the forward gather table (dense SubM layers; see ddf_sparse_conv_wgrad_table)."""
    cin, cout = filters.shape[-2], filters.shape[-1]
    kvol = gather_table.shape[1]
    gw = torch.empty_like(filters)
    with _lib.on_device(features.device):
        rc = _lib.get_lib().ddf_sparse_conv_wgrad_table(_lib.ptr(features), _lib.ptr(grad_out), _lib.ptr(gather_table),
                                                        _lib.ptr(gw), grad_out.shape[0], features.shape[0], kvol, cin,
                                                        cout, _lib.current_stream())
    _lib.check(rc, "sparse_conv_wgrad_table")
    return gw
