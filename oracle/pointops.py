"""ORACLE (test infrastructure): numpy restatements of the point-set CUDA ops used by the 3D local
self-attention. The reference has no CPU implementation of these; each function follows the kernel:

  furthest_point_sample  <proj>/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141
                         (incl. its tie-break: the per-thread strided scan keeps the lowest index; the pairwise
                         tree - strides B/2 ... 1, the LOWER slot keeps a tie - decides equal distances by the
                         low bits of the thread index first => smallest bit-reversed (k mod B), then lowest k,
                         B = largest power of two <= n capped at 1024, :9-13; verified against the compiled
                         reference kernel on lattice clouds, where exact ties are common)
  ball_query             <proj>/ops/ball_query/src/ball_query_cuda.cu:11-54
  grouping_operation     <proj>/ops/group_points/src/group_points_cuda.cu:10-31,56-79
  gather_points          <proj>/ops/gather_points/src/gather_points_cuda.cu:8-26,51-70

Pinned by the known-answer vectors of TransFusion/tests/test_models/test_common_modules/
test_pointnet_ops.py:9-24 (FPS), :26-73 (ball query incl. dilated), :126-196 (grouping),
:198-238 (gather) through tests/golden/pointops_golden.npz, and on the GPU box by the reference CUDA kernels themselves
(oracle/_ref/*_ext.so, tests/test_reference_cuda_gpu.py). Distances in fp32 with the reference kernels' FMA contraction.
"""
import numpy as np


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64, one rounding to float32 at the
    end (the float64 sum can itself round when the exponents are > 29 apart: harmless double rounding)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _sqdist(a, b):
    """Squared distance with the contraction nvcc gives the reference kernels (SASS of oracle/_ref/
    furthest_point_sample_ext.so, ball_query_ext.so built from the reference sources, default -fmad=true):
    d = fma(dz, dz, fma(dx, dx, dy * dy))."""
    d = (b - a).astype(np.float32)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return _fma32(dz, dz, _fma32(dx, dx, (dy * dy).astype(np.float32)))


def furthest_point_sample(xyz, npoint, temp=None):
    xyz = np.ascontiguousarray(xyz, np.float32)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    if npoint == 0:
        return idx
    block = 1
    while block * 2 <= N and block < 1024:
        block *= 2
    k = np.arange(N)
    bits = int(block).bit_length() - 1
    rev = np.zeros(N, np.int64)
    for bit in range(bits):                                        # bit reversal of (k mod block) in log2(block) bits
        rev |= (((k % block) >> bit) & 1) << (bits - 1 - bit)
    tie = rev * (1 << 21) + k // block   # smaller wins
    for b in range(B):
        dist = np.full((N,), 1e10, np.float32) if temp is None else temp[b]
        old = 0
        for j in range(1, npoint):
            d = _sqdist(xyz[b, old][None, :], xyz[b]).astype(np.float32)
            dist = np.minimum(d, dist)
            best = dist.max()
            cand = np.nonzero(dist == best)[0]
            old = int(cand[np.argmin(tie[cand])])
            idx[b, j] = old
        if temp is not None:
            temp[b] = dist
    return idx


def ball_query(min_radius, max_radius, nsample, xyz, new_xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((B, m, nsample), np.int32)
    r0 = np.float32(min_radius) * np.float32(min_radius)
    r1 = np.float32(max_radius) * np.float32(max_radius)
    for b in range(B):
        for p in range(m):
            d2 = _sqdist(xyz[b], new_xyz[b, p][None, :])  # (new - x)^2, same squares
            hits = np.nonzero((d2 == 0) | ((d2 >= r0) & (d2 < r1)))[0][:nsample]
            if len(hits):
                idx[b, p, :] = hits[0]
                idx[b, p, :len(hits)] = hits
    return idx


def grouping_operation(features, idx):
    """features (B,C,N), idx (B,np,ns) -> (B,C,np,ns)."""
    B, C, N = features.shape
    return np.stack([features[b][:, idx[b]] for b in range(B)])


def grouping_operation_grad(grad_out, idx, N):
    B, C = grad_out.shape[:2]
    g = np.zeros((B, C, N), grad_out.dtype)
    for b in range(B):
        np.add.at(g[b], (slice(None), idx[b].reshape(-1)), grad_out[b].reshape(C, -1))
    return g


def gather_points(points, idx):
    """points (B,C,N), idx (B,M) -> (B,C,M)."""
    return np.stack([points[b][:, idx[b]] for b in range(points.shape[0])])


def gather_points_grad(grad_out, idx, N):
    B, C = grad_out.shape[:2]
    g = np.zeros((B, C, N), grad_out.dtype)
    for b in range(B):
        np.add.at(g[b], (slice(None), idx[b]), grad_out[b])
    return g
