"""Point-set ops with the reference's names and call signatures
(<proj>/ops/furthest_point_sample/furthest_point_sample.py:7-40, points_sampler.py:34-152,
<proj>/ops/ball_query/ball_query.py:7-47, <proj>/ops/group_points/group_points.py:11-208 — the
CenterPoint / Voxel-RCNN flavour of QueryAndGroup that can return absolute grouped xyz and the
indices, which LocalTransformer needs — <proj>/ops/gather_points/gather_points.py:7-52)."""
from typing import List

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib as _lib


def _chk(*ts):
    _lib.require_cuda(*ts)
    for t in ts:
        assert t.is_contiguous()


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, points_xyz, num_points):
        _chk(points_xyz)
        B, N = points_xyz.size()[:2]
        output = torch.empty((B, num_points), dtype=torch.int32, device=points_xyz.device)
        with _lib.on_device(points_xyz.device):
            rc = _lib.get_lib().ddf_furthest_point_sampling(_lib.ptr(points_xyz.float()), None, _lib.ptr(output),
                                                           B, N, num_points, _lib.current_stream())
        _lib.check(rc, "furthest_point_sampling")
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, min_radius, max_radius, sample_num, xyz, center_xyz):
        _chk(xyz, center_xyz)
        assert min_radius < max_radius
        B, N, _ = xyz.size()
        npoint = center_xyz.size(1)
        idx = torch.zeros((B, npoint, sample_num), dtype=torch.int32, device=xyz.device)
        with _lib.on_device(xyz.device):
            rc = _lib.get_lib().ddf_ball_query(_lib.ptr(center_xyz), _lib.ptr(xyz), _lib.ptr(idx), B, N, npoint,
                                              float(min_radius), float(max_radius), sample_num,
                                              _lib.current_stream())
        _lib.check(rc, "ball_query")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


ball_query = BallQuery.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, indices):
        _chk(features, indices)
        B, nfeatures, nsample = indices.size()
        _, C, N = features.size()
        output = torch.empty((B, C, nfeatures, nsample), dtype=features.dtype, device=features.device)
        with _lib.on_device(features.device):
            rc = _lib.get_lib().ddf_group_points(_lib.ptr(features), _lib.ptr(indices), _lib.ptr(output), B, C, N,
                                                nfeatures, nsample, _lib.current_stream())
        _lib.check(rc, "group_points")
        ctx.for_backwards = (indices, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_out = grad_out.contiguous()
        grad_features = torch.empty((B, C, N), dtype=grad_out.dtype, device=grad_out.device)
        with _lib.on_device(grad_out.device):
            rc = _lib.get_lib().ddf_group_points_grad(_lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(grad_features),
                                                     B, C, N, npoint, nsample, _lib.current_stream())
        _lib.check(rc, "group_points_grad")
        return grad_features, None


grouping_operation = GroupingOperation.apply


class GatherPoints(Function):
    @staticmethod
    def forward(ctx, features, indices):
        _chk(features, indices)
        B, npoint = indices.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=features.dtype, device=features.device)
        with _lib.on_device(features.device):
            rc = _lib.get_lib().ddf_gather_points(_lib.ptr(features), _lib.ptr(indices), _lib.ptr(output), B, C, N,
                                                 npoint, _lib.current_stream())
        _lib.check(rc, "gather_points")
        ctx.for_backwards = (indices, C, N)
        ctx.mark_non_differentiable(indices)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_out = grad_out.contiguous()
        grad_features = torch.empty((B, C, N), dtype=grad_out.dtype, device=grad_out.device)
        with _lib.on_device(grad_out.device):
            rc = _lib.get_lib().ddf_gather_points_grad(_lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(grad_features),
                                                      B, C, N, npoint, _lib.current_stream())
        _lib.check(rc, "gather_points_grad")
        return grad_features, None


gather_points = GatherPoints.apply


class ScatterFirst(Function):
    """LocalTransformer.scatter with the 'unique' rule as device kernels: ``features`` (B, C, N) with the columns of
    every voxel that occurs in ``idx`` (B, npoint, nsample) replaced by ``feats`` (B, C, npoint, nsample) at the
    voxel's first occurrence in flattened order. Returns a new tensor (the reference overwrites in place)."""

    @staticmethod
    def forward(ctx, features, feats, idx):
        _lib.require_cuda(features, feats, idx)
        B, C, N = features.shape
        E = idx.shape[1] * idx.shape[2]
        feats = feats.contiguous().view(B, C, E)
        idx = idx.contiguous()
        if features.dtype != torch.float32 or feats.dtype != torch.float32 or idx.dtype != torch.int32:
            raise RuntimeError("scatter_first: float32 features / int32 indices only")
        out = features.clone(memory_format=torch.contiguous_format)
        first = torch.empty((B, N), dtype=torch.int32, device=features.device)
        L = _lib.get_lib()
        with _lib.on_device(features.device):
            rc = L.ddf_first_occurrence(_lib.ptr(idx), _lib.ptr(first), B, N, E, _lib.current_stream())
            _lib.check(rc, "first_occurrence")
            rc = L.ddf_scatter_first(_lib.ptr(feats), _lib.ptr(first), _lib.ptr(out), B, C, N, E, _lib.current_stream())
        _lib.check(rc, "scatter_first")
        ctx.save_for_backward(first)
        ctx.dims = (B, C, N, E, idx.shape[1], idx.shape[2])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (first,) = ctx.saved_tensors
        B, C, N, E, npoint, nsample = ctx.dims
        grad_out = grad_out.contiguous()
        g_feats = torch.empty((B, C, E), dtype=grad_out.dtype, device=grad_out.device)
        g_features = torch.empty_like(grad_out)
        with _lib.on_device(grad_out.device):
            rc = _lib.get_lib().ddf_scatter_first_grad(_lib.ptr(grad_out), _lib.ptr(first), _lib.ptr(g_feats),
                                                       _lib.ptr(g_features), B, C, N, E, _lib.current_stream())
        _lib.check(rc, "scatter_first_grad")
        return g_features, g_feats.view(B, C, npoint, nsample), None


scatter_first = ScatterFirst.apply


def first_occurrence(idx, N):
    """idx (B, npoint, nsample) int32 -> first (B, N) int32: flattened position of every voxel's first occurrence,
    npoint * nsample where the voxel is in no group."""
    _lib.require_cuda(idx)
    idx = idx.contiguous()
    B, E = idx.shape[0], idx.shape[1] * idx.shape[2]
    first = torch.empty((B, N), dtype=torch.int32, device=idx.device)
    with _lib.on_device(idx.device):
        rc = _lib.get_lib().ddf_first_occurrence(_lib.ptr(idx), _lib.ptr(first), B, N, E, _lib.current_stream())
    _lib.check(rc, "first_occurrence")
    return first


def local_attn_supported(heads, head_dim, group_size):
    return bool(_lib.get_lib().ddf_local_attn_supported(int(heads), int(head_dim), int(group_size)))


class LocalAttention(Function):
    """softmax(q k^T / sqrt(hd)) v inside every group of 32 tokens, per head, on token-major projections:
    qkv (T, 3C) -> (T, C), T = groups * 32 (csrc/local_attn.cu)."""

    @staticmethod
    def forward(ctx, qkv, heads, group_size):
        _lib.require_cuda(qkv)
        qkv = qkv.contiguous()
        if qkv.dtype != torch.float32:
            raise RuntimeError("local attention: float32 only")
        T, C3 = qkv.shape
        C = C3 // 3
        out = torch.empty((T, C), dtype=qkv.dtype, device=qkv.device)
        with _lib.on_device(qkv.device):
            rc = _lib.get_lib().ddf_local_attn_forward(_lib.ptr(qkv), _lib.ptr(out), T // group_size, heads, C // heads,
                                                       group_size, _lib.current_stream())
        _lib.check(rc, "local_attn_forward")
        ctx.save_for_backward(qkv)
        ctx.dims = (heads, group_size)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (qkv,) = ctx.saved_tensors
        heads, group_size = ctx.dims
        grad_out = grad_out.contiguous()
        T, C3 = qkv.shape
        g = torch.empty_like(qkv)
        with _lib.on_device(qkv.device):
            rc = _lib.get_lib().ddf_local_attn_backward(_lib.ptr(qkv), _lib.ptr(grad_out), _lib.ptr(g), T // group_size,
                                                        heads, C3 // 3 // heads, group_size, _lib.current_stream())
        _lib.check(rc, "local_attn_backward")
        return g, None, None


local_attention = LocalAttention.apply


class DFPS_Sampler(nn.Module):
    def forward(self, points, features, npoint):
        return furthest_point_sample(points.contiguous(), npoint)


def get_sampler_type(sampler_type):
    if sampler_type == "D-FPS":
        return DFPS_Sampler
    raise ValueError('Only "D-FPS" is used by 3D-DF (F-FPS / FS samplers are off the hot path), got %s'
                     % sampler_type)


class Points_Sampler(nn.Module):
    def __init__(self, num_point: List[int], fps_mod_list: List[str] = ["D-FPS"],
                 fps_sample_range_list: List[int] = [-1]):
        super(Points_Sampler, self).__init__()
        assert len(num_point) == len(fps_mod_list) == len(fps_sample_range_list)
        self.num_point = num_point
        self.fps_sample_range_list = fps_sample_range_list
        self.samplers = nn.ModuleList([get_sampler_type(m)() for m in fps_mod_list])
        self.fp16_enabled = False

    def forward(self, points_xyz, features):
        indices = []
        last = 0
        for rng, sampler, npoint in zip(self.fps_sample_range_list, self.samplers, self.num_point):
            assert rng < points_xyz.shape[1]
            if rng == -1:
                xyz = points_xyz[:, last:]
                feats = features[:, :, last:] if features is not None else None
            else:
                xyz = points_xyz[:, last:rng]
                feats = features[:, :, last:rng] if features is not None else None
            indices.append(sampler(xyz.contiguous(), feats, npoint) + last)
            last += rng
        return torch.cat(indices, dim=1)


class QueryAndGroup(nn.Module):
    def __init__(self, max_radius, sample_num, min_radius=0, use_xyz=True, return_grouped_xyz=False,
                 normalize_xyz=False, uniform_sample=False, return_unique_cnt=False,
                 return_grouped_idx=False):
        super(QueryAndGroup, self).__init__()
        if max_radius is None:
            raise NotImplementedError("kNN grouping is off the 3D-DF hot path (radius is always given)")
        if uniform_sample or return_unique_cnt:
            raise NotImplementedError("uniform_sample is off the 3D-DF hot path")
        self.max_radius = max_radius
        self.min_radius = min_radius
        self.sample_num = sample_num
        self.use_xyz = use_xyz
        self.return_grouped_xyz = return_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.return_grouped_idx = return_grouped_idx

    def forward(self, points_xyz, center_xyz, features=None):
        idx = ball_query(self.min_radius, self.max_radius, self.sample_num, points_xyz, center_xyz)
        xyz_trans = points_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz_diff = grouped_xyz - center_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz_diff = grouped_xyz_diff / self.max_radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz_diff, grouped_features], dim=1) if self.use_xyz \
                else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz_diff
        ret = [new_features]
        if self.return_grouped_xyz:
            ret.append(grouped_xyz)
        if self.return_grouped_idx:
            ret.append(idx)
        return ret[0] if len(ret) == 1 else tuple(ret)
