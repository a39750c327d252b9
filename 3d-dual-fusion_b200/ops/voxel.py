"""Voxelization ops: the reference's ``voxelization`` / ``Voxelization`` names and call signatures
(``TransFusion/mmdet3d/ops/voxel/voxelize.py:13-122``) on top of ``ddf_hard_voxelize`` /
``ddf_dynamic_voxelize`` (include/ddf_b200.h).
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from .. import lib as _lib


def _f32_array(vals, n):
    vals = [float(v) for v in vals]
    if len(vals) != n:
        raise RuntimeError("expected %d floats, got %d" % (n, len(vals)))
    return (ctypes.c_float * n)(*vals)


def _check_points(points):
    _lib.require_cuda(points)
    if points.dim() != 2 or points.size(1) < 3:
        raise RuntimeError("points must be [N, >=3]")
    if points.dtype != torch.float32:
        raise RuntimeError("points must be float32")
    return points.contiguous()


def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """Same contract as ``voxel_layer.dynamic_voxelize`` (voxelization.h:71-86): fills ``coors``."""
    points = _check_points(points)
    if NDim != 3:
        raise RuntimeError("only NDim=3 is supported")
    with _lib.on_device(points.device):
        rc = _lib.get_lib().ddf_dynamic_voxelize(
            _lib.ptr(points), _lib.ptr(coors), _f32_array(voxel_size, 3), _f32_array(coors_range, 6),
            points.size(0), points.size(1), _lib.current_stream())
    _lib.check(rc, "dynamic_voxelize")


def hard_voxelize_device(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                         max_points, max_voxels):
    """Enqueue hard voxelization; returns the voxel count as a 1-element int32 DEVICE tensor
    (no host synchronisation)."""
    points = _check_points(points)
    L = _lib.get_lib()
    n, f = points.shape
    ws_bytes = L.ddf_hard_voxelize_workspace_bytes(n, max_points, max_voxels)
    if ws_bytes < 0:
        raise RuntimeError("hard_voxelize: bad sizes")
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=points.device)
    voxel_num = torch.empty(1, dtype=torch.int32, device=points.device)
    with _lib.on_device(points.device):
        rc = L.ddf_hard_voxelize(
            _lib.ptr(points), _lib.ptr(voxels), _lib.ptr(coors), _lib.ptr(num_points_per_voxel),
            _lib.ptr(voxel_num), _f32_array(voxel_size, 3), _f32_array(coors_range, 6), n, f,
            int(max_points), int(max_voxels), _lib.ptr(ws), int(ws_bytes), _lib.current_stream())
    _lib.check(rc, "hard_voxelize")
    return voxel_num


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                  max_voxels, NDim=3):
    """Same contract as ``voxel_layer.hard_voxelize`` (voxelization.h:51-69): fills the three
    caller-allocated outputs and returns ``voxel_num`` as a Python int (one D2H read)."""
    if NDim != 3:
        raise RuntimeError("only NDim=3 is supported")
    return int(hard_voxelize_device(points, voxels, coors, num_points_per_voxel, voxel_size,
                                    coors_range, max_points, max_voxels).item())


def hard_voxelize_mean(points, voxel_size, coors_range, max_points, max_voxels, num_features):
    """Hard voxelization fused with ``HardSimpleVFE``: returns (mean [M, num_features], coors [M, 3], num_points [M])
    without materialising the padded [M, max_points, F] voxel tensor (one D2H read of the voxel count, like
    ``hard_voxelize``)."""
    points = _check_points(points)
    L = _lib.get_lib()
    n, f = points.shape
    cap = min(int(max_voxels), n)
    mean = points.new_empty((cap, int(num_features)))
    coors = points.new_empty((cap, 3), dtype=torch.int)
    num = points.new_empty((cap,), dtype=torch.int)
    ws_bytes = L.ddf_hard_voxelize_workspace_bytes(n, max_points, cap)
    if ws_bytes < 0:
        raise RuntimeError("hard_voxelize_mean: bad sizes")
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=points.device)
    voxel_num = torch.empty(1, dtype=torch.int32, device=points.device)
    with _lib.on_device(points.device):
        rc = L.ddf_hard_voxelize_mean(
            _lib.ptr(points), _lib.ptr(mean), _lib.ptr(coors), _lib.ptr(num), _lib.ptr(voxel_num),
            _f32_array(voxel_size, 3), _f32_array(coors_range, 6), n, f, int(num_features), int(max_points), cap,
            _lib.ptr(ws), int(ws_bytes), _lib.current_stream())
    _lib.check(rc, "hard_voxelize_mean")
    m = int(voxel_num.item())
    return mean[:m], coors[:m], num[:m]


class _Voxelization(Function):
    """Drop-in for voxelize.py:11-58."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_empty(size=(points.size(0), 3), dtype=torch.int)
            dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        # rows [0, voxel_num) are fully written by the library, so empty (not zeros) buffers do
        cap = min(int(max_voxels), points.size(0))
        voxels = points.new_empty(size=(cap, max_points, points.size(1)))
        coors = points.new_empty(size=(cap, 3), dtype=torch.int)
        num_points_per_voxel = points.new_empty(size=(cap,), dtype=torch.int)
        voxel_num = hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
                                  coors_range, max_points, cap, 3)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """Drop-in for voxelize.py:64-122 (same constructor arguments and attributes)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        super(Voxelization, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        if isinstance(max_voxels, tuple):
            self.max_voxels = max_voxels
        else:
            self.max_voxels = _pair(max_voxels)
        point_cloud_range = torch.tensor(point_cloud_range, dtype=torch.float32)
        voxel_size = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = torch.round(grid_size).long()
        input_feat_shape = grid_size[:2]
        self.grid_size = grid_size
        self.pcd_shape = [*input_feat_shape, 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            max_voxels)

    def forward_mean(self, input, num_features):
        """``HardSimpleVFE(num_features)(*self(input))`` as one pass (ddf_hard_voxelize_mean)."""
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return hard_voxelize_mean(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels,
                                  num_features)

    def __repr__(self):
        tmpstr = self.__class__.__name__ + '('
        tmpstr += 'voxel_size=' + str(self.voxel_size)
        tmpstr += ', point_cloud_range=' + str(self.point_cloud_range)
        tmpstr += ', max_num_points=' + str(self.max_num_points)
        tmpstr += ', max_voxels=' + str(self.max_voxels)
        tmpstr += ')'
        return tmpstr
