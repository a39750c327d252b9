"""SparseConvTensor with the attribute surface the three reference backbones touch
(TransFusion/mmdet3d/ops/spconv/structure.py:21-69; spconv-2 style ``replace_feature`` probed by
CenterPoint/det3d/models/backbones/scn.py:17-23 and VoxelRCNN/pcdet/utils/spconv_utils.py:28-34)."""
import numpy as np
import torch
from torch.autograd import Function

from ... import lib as _lib


class _ToDense(Function):
    """dense() as one kernel writing NCDHW directly (ddf_sparse_to_dense / ddf_dense_to_sparse)."""

    @staticmethod
    def forward(ctx, features, indices, spatial_shape, batch_size):
        _lib.require_cuda(features, indices)
        features = features.contiguous()
        if features.dtype != torch.float32:
            raise RuntimeError("dense(): float32 features only")
        n, C = features.shape
        D, H, W = [int(s) for s in spatial_shape]
        out = torch.empty((batch_size, C, D, H, W), dtype=features.dtype, device=features.device)
        with _lib.on_device(features.device):
            rc = _lib.get_lib().ddf_sparse_to_dense(_lib.ptr(features), _lib.ptr(indices), _lib.ptr(out),
                                                    n, C, batch_size, D, H, W, _lib.current_stream())
        _lib.check(rc, "sparse_to_dense")
        ctx.save_for_backward(indices)
        ctx.dims = (n, C, batch_size, D, H, W)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (indices,) = ctx.saved_tensors
        n, C, B, D, H, W = ctx.dims
        grad_out = grad_out.contiguous()
        gfeat = torch.empty((n, C), dtype=grad_out.dtype, device=grad_out.device)
        with _lib.on_device(grad_out.device):
            rc = _lib.get_lib().ddf_dense_to_sparse(_lib.ptr(grad_out), _lib.ptr(indices), _lib.ptr(gfeat),
                                                    n, C, B, D, H, W, _lib.current_stream())
        _lib.check(rc, "dense_to_sparse")
        return gfeat, None, None, None


class _ToBevNhwcBf16(Function):
    """[N, C] fp32 rows -> the BEV map [B, C*D, H, W] in bf16 with channels-last storage, one kernel
    (ddf_sparse_to_bev_nhwc_bf16); backward gathers the bf16 gradient back into fp32 rows."""

    @staticmethod
    def forward(ctx, features, indices, spatial_shape, batch_size):
        _lib.require_cuda(features, indices)
        features = features.contiguous()
        if features.dtype != torch.float32:
            raise RuntimeError("dense_bev_bf16(): float32 features only")
        n, C = features.shape
        D, H, W = [int(s) for s in spatial_shape]
        out = torch.empty((batch_size, C * D, H, W), dtype=torch.bfloat16, device=features.device,
                          memory_format=torch.channels_last)
        with _lib.on_device(features.device):
            rc = _lib.get_lib().ddf_sparse_to_bev_nhwc_bf16(_lib.ptr(features), _lib.ptr(indices), _lib.ptr(out),
                                                            n, C, batch_size, D, H, W, _lib.current_stream())
        _lib.check(rc, "sparse_to_bev_nhwc_bf16")
        ctx.save_for_backward(indices)
        ctx.dims = (n, C, batch_size, D, H, W)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (indices,) = ctx.saved_tensors
        n, C, B, D, H, W = ctx.dims
        grad_out = grad_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gfeat = torch.empty((n, C), dtype=torch.float32, device=grad_out.device)
        with _lib.on_device(grad_out.device):
            rc = _lib.get_lib().ddf_bev_nhwc_bf16_to_sparse(_lib.ptr(grad_out), _lib.ptr(indices), _lib.ptr(gfeat),
                                                            n, C, B, D, H, W, _lib.current_stream())
        _lib.check(rc, "bev_nhwc_bf16_to_sparse")
        return gfeat, None, None, None


def scatter_nd(indices, updates, shape):
    """Same contract as structure.py:5-18 (index-put into zeros), kept for API parity; the hot
    path (SparseConvTensor.dense) uses the fused kernel instead."""
    ret = torch.zeros(*shape, dtype=updates.dtype, device=updates.device)
    ndim = indices.shape[-1]
    flat = indices.reshape(-1, ndim)
    out_shape = list(indices.shape[:-1]) + list(shape[ndim:])
    ret[tuple(flat[:, i] for i in range(ndim)) + (Ellipsis,)] = updates.view(*out_shape)
    return ret


class SparseConvTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        if indices.dtype != torch.int32:
            indices = indices.int()
        self.indices = indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid

    def replace_feature(self, feature):
        """spconv-2 style functional update: a new tensor sharing indices and rulebooks."""
        out = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid)
        out.indice_dict = self.indice_dict
        return out

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        if len(self.spatial_shape) != 3:
            raise RuntimeError("dense(): 3-D sparse tensors only")
        res = _ToDense.apply(self.features, self.indices, self.spatial_shape, self.batch_size)
        if channels_first:
            return res
        return res.permute(0, 2, 3, 4, 1).contiguous()

    def dense_bev_bf16(self):
        """``dense().view(B, C*D, H, W)`` for a bf16 channels-last 2-D backbone: same values rounded to bf16, written
        directly in channels-last storage (no fp32 NCDHW tensor, no permute)."""
        if len(self.spatial_shape) != 3:
            raise RuntimeError("dense_bev_bf16(): 3-D sparse tensors only")
        return _ToBevNhwcBf16.apply(self.features, self.indices, self.spatial_shape, self.batch_size)

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size
