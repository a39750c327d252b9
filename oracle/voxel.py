"""ORACLE (test infrastructure): numpy front-end of oracle/voxel_ref.c — restates
TransFusion/mmdet3d/ops/voxel/src/voxelization_cpu.cpp:8-102 (hard / dynamic voxelization)."""
import ctypes

import numpy as np

from . import lib

_I64 = ctypes.c_int64


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32))


def dynamic_voxelize(points, voxel_size, coors_range):
    points = _f32(points)
    n, F = points.shape
    coors = np.empty((n, 3), np.int32)
    fn = lib().oracle_dynamic_voxelize
    fn.restype = None
    vs, rg = _f32(voxel_size), _f32(coors_range)
    fn(_p(points), _p(coors), _p(vs), _p(rg), _I64(n), _I64(F))
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    """Returns (voxels[:M], coors[:M], num_points[:M]) like voxelize.py:46-58."""
    points = _f32(points)
    n, F = points.shape
    voxels = np.zeros((max_voxels, max_points, F), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    fn = lib().oracle_hard_voxelize
    fn.restype = _I64
    vs, rg = _f32(voxel_size), _f32(coors_range)
    m = fn(_p(points), _p(voxels), _p(coors), _p(num), _p(vs), _p(rg), _I64(n), _I64(F),
           _I64(max_points), _I64(max_voxels))
    return voxels[:m], coors[:m], num[:m]
