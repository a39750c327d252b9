"""Multi-scale deformable attention op: the reference's ``MSDeformAttnFunction`` name and call
signature (``<proj>/models/model_utils/ops/functions/ms_deform_attn_func.py:21-38``) on top of
``ddf_ms_deform_attn_forward/backward`` (include/ddf_b200.h).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib as _lib

_DTYPES = {torch.float32: 0, torch.float64: 1}


def _check_inputs(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    # reference asserts: ms_deform_attn_cuda.cu:28-38
    names = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
    for n, t in zip(names, (value, spatial_shapes, level_start_index, sampling_locations,
                            attention_weights)):
        if not t.is_contiguous():
            raise RuntimeError("%s tensor has to be contiguous" % n)
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % n)
    if value.dtype not in _DTYPES:
        raise RuntimeError("ms_deform_attn: unsupported dtype %s" % value.dtype)
    if sampling_locations.dtype != value.dtype or attention_weights.dtype != value.dtype:
        raise RuntimeError("ms_deform_attn: value / sampling_loc / attn_weight dtypes differ")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("ms_deform_attn: spatial_shapes / level_start_index must be int64")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step):
    """Same contract as ``MSDA.ms_deform_attn_forward`` (ops/src/ms_deform_attn.h:20-39)."""
    _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with _lib.on_device(value.device):
        rc = _lib.get_lib().ddf_ms_deform_attn_forward(
            _lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index),
            _lib.ptr(sampling_loc), _lib.ptr(attn_weight), _lib.ptr(out), N, S, M, D, L, Lq, P,
            int(im2col_step), _DTYPES[value.dtype], _lib.current_stream())
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, im2col_step):
    """Same contract as ``MSDA.ms_deform_attn_backward`` (ops/src/ms_deform_attn.h:41-62)."""
    _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    if not grad_output.is_contiguous():
        raise RuntimeError("grad_output tensor has to be contiguous")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    grad_value = torch.empty_like(value)  # zeroed inside the library call
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with _lib.on_device(value.device):
        rc = _lib.get_lib().ddf_ms_deform_attn_backward(
            _lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index),
            _lib.ptr(sampling_loc), _lib.ptr(attn_weight), _lib.ptr(grad_output),
            _lib.ptr(grad_value), _lib.ptr(grad_loc), _lib.ptr(grad_attn), N, S, M, D, L, Lq, P,
            int(im2col_step), _DTYPES[value.dtype], _lib.current_stream())
    _lib.check(rc, "ms_deform_attn_backward")
    return grad_value, grad_loc, grad_attn


class MSDeformAttnFunction(Function):
    """Drop-in for the reference autograd Function (ms_deform_attn_func.py:21-38)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                        sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = ms_deform_attn_backward(
            value, shapes, lsi, loc, attn, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


# ---- tile-staged dual-query form (hot path) -------------------------------------------------------
def tile_supported(M, D, L, P, dtype=torch.float32):
    """True when the fused tile-staged kernels take this shape (one level, 4 points, D in {8, 16}, fp32)."""
    return dtype == torch.float32 and bool(_lib.get_lib().ddf_msda_tile_supported(int(M), int(D), int(L), int(P)))


class TilePlan(object):
    """Queries binned by feature-map tile (ddf_msda_plan). Built once per encoder forward from the reference
    points; shared by every layer's forward and backward.

    Regular layout: ``reference_points`` (N, Lq, [1,] 2), query (b, q) samples image b.  Ragged layout
    (``query_batch`` given): ``reference_points`` (NQ, 2) are only the real queries of the padded per-camera layout and
    ``query_batch`` (NQ,) int32 names the image each one samples; the per-query tensors handed to the kernels are then
    (1, NQ, ...)."""

    def __init__(self, reference_points, H, W, query_batch=None, n_images=None):
        ref = reference_points
        _lib.require_cuda(ref)
        if query_batch is None:
            if ref.dim() == 4:           # (N, Lq, L=1, 2)
                ref = ref[:, :, 0]
            self.N, self.Lq = ref.shape[0], ref.shape[1]
            self.NQ = self.N * self.Lq
            self.qbatch = None
        else:
            ref = ref.reshape(-1, 2)
            self.N, self.Lq, self.NQ = int(n_images), 0, ref.shape[0]
            self.qbatch = query_batch.to(torch.int32).contiguous()
        self.ref = ref.detach().contiguous().float()
        self.H, self.W = int(H), int(W)
        L = _lib.get_lib()
        nbytes = int(L.ddf_msda_plan_bytes(self.N, self.NQ, self.H, self.W))
        self.buf = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=ref.device)
        with _lib.on_device(ref.device):
            rc = L.ddf_msda_plan(_lib.ptr(self.ref), _lib.ptr(self.qbatch), _lib.ptr(self.buf), self.N, self.NQ, self.Lq,
                                 self.H, self.W, _lib.current_stream())
        _lib.check(rc, "msda_plan")


def msda_tile_forward(value, plan, offsets, logits):
    N, S, M, D = value.shape
    out = torch.empty(tuple(offsets.shape[:2]) + (M * D,), dtype=value.dtype, device=value.device)
    with _lib.on_device(value.device):
        rc = _lib.get_lib().ddf_msda_tile_forward(
            _lib.ptr(value), _lib.ptr(plan.ref), _lib.ptr(offsets), _lib.ptr(logits), _lib.ptr(plan.buf),
            _lib.ptr(out), N, plan.H, plan.W, M, D, plan.NQ, _lib.current_stream())
    _lib.check(rc, "msda_tile_forward")
    return out


def msda_tile_backward(value, plan, offsets, logits, grad_out):
    N, S, M, D = value.shape
    grad_value = torch.empty_like(value)   # zeroed inside the library call
    grad_off = torch.empty_like(offsets)
    grad_logit = torch.empty_like(logits)
    with _lib.on_device(value.device):
        rc = _lib.get_lib().ddf_msda_tile_backward(
            _lib.ptr(value), _lib.ptr(plan.ref), _lib.ptr(offsets), _lib.ptr(logits), _lib.ptr(grad_out),
            _lib.ptr(plan.buf), _lib.ptr(grad_value), _lib.ptr(grad_off), _lib.ptr(grad_logit), N, plan.H, plan.W,
            M, D, plan.NQ, _lib.current_stream())
    _lib.check(rc, "msda_tile_backward")
    return grad_value, grad_off, grad_logit


class MSDeformAttnTileFunction(Function):
    """out = MSDA(value, loc = ref + offsets / (W, H), softmax(logits)): the module arithmetic of
    ops/modules/ms_deform_attn.py:149-188 as one op. value (N, H*W, M, D), offsets (N, Lq, M, 1, 4, 2) raw,
    logits (N, Lq, M, 4) - or (1, NQ, ...) for a ragged plan; reference points live in ``plan`` (no gradient flows
    to them)."""

    @staticmethod
    def forward(ctx, value, offsets, logits, plan):
        for t in (value, offsets, logits):
            if t.dtype != torch.float32 or not t.is_cuda:
                raise RuntimeError("msda tile kernels: CUDA float32 tensors only")
        value, offsets, logits = value.contiguous(), offsets.contiguous(), logits.contiguous()
        if (value.shape[0] != plan.N or value.shape[1] != plan.H * plan.W
                or offsets.shape[0] * offsets.shape[1] != plan.NQ or logits.shape[0] * logits.shape[1] != plan.NQ):
            raise RuntimeError("msda tile kernels: plan was built for another problem size")
        ctx.plan = plan
        ctx.save_for_backward(value, offsets, logits)
        return msda_tile_forward(value, plan, offsets, logits)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, offsets, logits = ctx.saved_tensors
        gv, go, gl = msda_tile_backward(value, ctx.plan, offsets, logits, grad_output.contiguous())
        return gv, go, gl, None
