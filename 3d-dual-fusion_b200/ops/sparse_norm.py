"""BatchNorm1d (+ residual add) (+ ReLU) over sparse-voxel features ``[N, C]`` as one fused op
(``ddf_sparse_bn_forward`` / ``ddf_sparse_bn_backward``, include/ddf_b200.h).

The reference runs ``nn.BatchNorm1d`` -> ``+ identity`` -> ``nn.ReLU`` as separate modules
(TransFusion/mmdet3d/ops/sparse_block.py:102-120, 153-185); ``batch_norm_act`` takes the SAME
``nn.BatchNorm1d`` module (parameters, buffers and state-dict keys untouched) and computes the
chain in two launches forward and two backward.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib as _lib

_workspaces = {}


def _workspace(device, C):
    """Zero-initialised once per (device, stream, C); the kernels leave it reusable."""
    key = (device.index, torch._C._cuda_getCurrentRawStream(device.index), C)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(_lib.get_lib().ddf_sparse_bn_workspace_bytes(C))
        if nbytes < 0:
            raise RuntimeError("sparse_bn: unsupported channel count %d" % C)
        ws = _workspaces[key] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    return ws


def supported(x):
    C = x.shape[1] if x.dim() == 2 else 0
    return x.dim() == 2 and x.dtype == torch.float32 and 4 <= C <= 1024 and (C & (C - 1)) == 0


class _BatchNormAct(Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu, emit):
        """``emit``: 0 = y only; 1 = also y in the bf16x3 operand layout of the convs; 2 = that and the tf32-rounded
        copy for the wgrad kernels (ddf_sparse_bn_forward_split)."""
        _lib.require_cuda(x, residual, weight, bias, running_mean, running_var)
        x = x.contiguous()
        residual = residual.contiguous() if residual is not None else None
        n, C = x.shape
        y = torch.empty_like(x)
        split = torch.empty_like(x) if emit else None
        rounded = torch.empty_like(x) if emit > 1 else None
        ws = _workspace(x.device, C) if training else None
        if training:
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        else:
            mean = invstd = None
        with _lib.on_device(x.device):
            if emit:
                rc = _lib.get_lib().ddf_sparse_bn_forward_split(
                    _lib.ptr(x), _lib.ptr(residual), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(running_mean),
                    _lib.ptr(running_var), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(split),
                    _lib.ptr(rounded), n, C, int(training), float(momentum), float(eps), int(relu), _lib.ptr(ws),
                    _lib.current_stream())
            else:
                rc = _lib.get_lib().ddf_sparse_bn_forward(
                    _lib.ptr(x), _lib.ptr(residual), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(running_mean),
                    _lib.ptr(running_var), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(invstd), n, C, int(training),
                    float(momentum), float(eps), int(relu), _lib.ptr(ws), _lib.current_stream())
        _lib.check(rc, "sparse_bn_forward")
        if not training:
            mean = running_mean
            invstd = torch.rsqrt(running_var + eps)
        ctx.training, ctx.relu, ctx.has_res = bool(training), bool(relu), residual is not None
        ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
        if not emit:
            return y
        extras = [t for t in (split, rounded) if t is not None]
        ctx.mark_non_differentiable(*extras)
        ctx.set_materialize_grads(False)     # no zero tensors for the operand copies in backward
        return (y, split, rounded) if rounded is not None else (y, split)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y, *_unused):
        if grad_y is None:
            return (None,) * 11
        x, y, weight, mean, invstd = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        n, C = x.shape
        need_x, need_res, need_w, need_b = ctx.needs_input_grad[:4]
        need_res = need_res and ctx.has_res
        gx = torch.empty_like(x) if need_x else None
        if need_res and not ctx.relu:
            gres = grad_y       # identity branch without a mask: the incoming gradient itself
            gres_out = None
        else:
            gres = gres_out = torch.empty_like(x) if need_res else None
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if (need_w and weight is not None) else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if need_b else None
        ws = _workspace(x.device, C)
        with _lib.on_device(x.device):
            rc = _lib.get_lib().ddf_sparse_bn_backward(
                _lib.ptr(grad_y), _lib.ptr(y), _lib.ptr(x), _lib.ptr(weight), _lib.ptr(mean), _lib.ptr(invstd),
                _lib.ptr(gx), _lib.ptr(gres_out), _lib.ptr(gw), _lib.ptr(gb), n, C, int(ctx.training),
                int(ctx.relu), _lib.ptr(ws), _lib.current_stream())
        _lib.check(rc, "sparse_bn_backward")
        return gx, gres, gw, gb, None, None, None, None, None, None, None


def batch_norm_act(bn, x, residual=None, relu=False):
    """``relu(bn(x) + residual)`` with ``bn`` an ``nn.BatchNorm1d`` (train or eval mode)."""
    if not supported(x) or x.shape[0] == 0:
        # shapes the fused kernels do not take (C not a power of two): the library modules, on the GPU
        _lib.require_cuda(x)
        y = bn(x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    use_batch_stats = bn.training or not bn.track_running_stats
    momentum = 0.0
    if bn.training and bn.track_running_stats:
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        momentum = (1.0 / float(bn.num_batches_tracked)) if bn.momentum is None else bn.momentum
    rm = bn.running_mean if (bn.track_running_stats and (bn.training or not use_batch_stats)) else None
    rv = bn.running_var if rm is not None else None
    # when the bf16x3 tensor-core convs are on, the consumer of y is (almost always) a conv that wants y in its operand
    # layout (+ the tf32-rounded copy for wgrad in training): written by the same pass and handed over on the tensor
    # (TableConvFunction.forward picks ``_ddf_operands`` up; anybody else just sees y)
    emit = 0
    if x.shape[1] % 32 == 0 and _conv_ops().bf16x3_on(x.shape[1]):
        emit = 2 if torch.is_grad_enabled() else 1
    out = _BatchNormAct.apply(x, residual, bn.weight, bn.bias, rm, rv, use_batch_stats, momentum, bn.eps, relu, emit)
    if not emit:
        return out
    y = out[0]
    y._ddf_operands = (out[1], out[2] if len(out) > 2 else None)
    return y


def _conv_ops():
    from .spconv import ops
    return ops
