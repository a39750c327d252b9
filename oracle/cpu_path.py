"""ORACLE (test infrastructure): run the hot path on the HOST cores with the reference's own CPU code.

Used only by bench.py (`cpu_baseline`, `--impl reference`) and tests. It assembles the same module
graph as the product (so the work is identical) but swaps every native op for the reference's CPU
implementation from OUTSIDE the product package:

    hard_voxelize           -> oracle/_ref/voxel_layer.so (reference voxelization_cpu.cpp, unmodified)
                               or the C restatement oracle/voxel_ref.c when _ref was never built
    rulebook / conv fwd/bwd -> oracle/_ref/sparse_conv_ext.so CPU path (reference geometry.h,
                               spconv_ops.h, unmodified) or oracle/spconv_ref.c
    dense()                 -> scatter_nd + permute (reference structure.py:5-18,55-64)
    MSDA                    -> the reference's pure-PyTorch ms_deform_attn_core_pytorch algorithm
                               (grid_sample; ms_deform_attn_func.py:41-61) with autograd backward —
                               the reference has no CPU kernel for it

The product package itself contains no CPU branch; nothing here is reachable from it.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import ref_build
from . import pointops as opointops
from . import spconv as ospconv
from . import voxel as ovoxel


def kind():
    """'reference' when the reference's own extensions are available, else 'port'."""
    return "reference" if (ref_build.load("voxel_layer") is not None
                           and ref_build.load("sparse_conv_ext") is not None) else "port"


# ---- MSDA: the reference's pure-PyTorch core ---------------------------------------------------
def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    N_, S_, M_, D_ = value.shape
    _, Lq_, M_, L_, P_, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in value_spatial_shapes.tolist()]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampled = []
    for lid, (H_, W_) in enumerate(shapes):
        value_l = value_list[lid].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        grid_l = sampling_grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(value_l, grid_l, mode="bilinear", padding_mode="zeros", align_corners=False))
    attention_weights = attention_weights.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * attention_weights).sum(-1).view(N_, M_ * D_, Lq_)
    return out.transpose(1, 2).contiguous()


class _CpuMSDA(object):
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return ms_deform_attn_core_pytorch(value, shapes, loc, attn)


# ---- voxelization ------------------------------------------------------------------------------
def cpu_voxelization(points, voxel_size, coors_range, max_points=35, max_voxels=20000):
    ext = ref_build.load("voxel_layer")
    if ext is not None:
        voxels = points.new_zeros((max_voxels, max_points, points.size(1)))
        coors = points.new_zeros((max_voxels, 3), dtype=torch.int)
        num = points.new_zeros((max_voxels,), dtype=torch.int)
        m = ext.hard_voxelize(points, voxels, coors, num, list(voxel_size), list(coors_range),
                              max_points, max_voxels, 3)
        return voxels[:m], coors[:m], num[:m]
    v, c, n = ovoxel.hard_voxelize(points.numpy(), voxel_size, coors_range, max_points, max_voxels)
    return torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(n)


def cpu_hard_voxelize_mean(points, voxel_size, coors_range, max_points, max_voxels, num_features):
    """Reference voxelization followed by the reference's HardSimpleVFE arithmetic (voxel_encoder.py:42-44)."""
    v, c, n = cpu_voxelization(points, voxel_size, coors_range, max_points, max_voxels)
    mean = v[:, :, :num_features].sum(dim=1, keepdim=False) / n.type_as(v).view(-1, 1)
    return mean.contiguous(), c, n


# ---- sparse conv -------------------------------------------------------------------------------
class _CpuRulebook(object):
    def __init__(self, outids, pairs, num, out_shape):
        self.outids, self.indice_pairs, self.indice_pair_num = outids, pairs, num
        self.out_spatial_shape = out_shape
        self.gather_table = self.scatter_table = None


def cpu_build_rulebook(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                       out_padding=0, subm=False, transpose=False, with_tables=True):
    as3 = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * 3
    ksize, stride, padding, dilation = as3(ksize), as3(stride), as3(padding), as3(dilation)
    out_shape = list(spatial_shape) if subm else ospconv.get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    ext = ref_build.load("sparse_conv_ext")
    if ext is not None:
        outids, pairs, num = ext.get_indice_pairs_3d(indices, batch_size, out_shape, list(spatial_shape), ksize,
                                                     stride, padding, dilation, [0, 0, 0], int(subm), 0)
    else:
        o, p, n, _ = ospconv.get_indice_pairs(indices.numpy(), batch_size, spatial_shape, ksize, stride, padding, dilation, subm)
        outids, pairs, num = torch.from_numpy(np.ascontiguousarray(o)), torch.from_numpy(p), torch.from_numpy(n)
    if not subm and outids.shape[0] > 0:
        # the reference CPU path numbers output voxels in first-touch order, its GPU path (and the
        # product) in sorted flat (b,z,y,x) order (spconv_ops.h:129-137); ops that depend on the row
        # order (FPS start / ball-query order in the LocalTransformer) need the same order on both sides
        flat = outids[:, 0].long()
        for d in range(3):
            flat = flat * out_shape[d] + outids[:, 1 + d].long()
        perm = torch.argsort(flat, stable=True)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel())
        outids = outids[perm].contiguous()
        pairs = pairs.clone()
        for k in range(pairs.shape[0]):
            nk = int(num[k])
            pairs[k, 1, :nk] = inv[pairs[k, 1, :nk].long()].int()
    return _CpuRulebook(outids, pairs, num, out_shape)


class _CpuTableConv(Function):
    @staticmethod
    def forward(ctx, features, filters, bias, rb, n_out):
        ctx.rb = rb
        ctx.subm = int(rb.outids.shape[0] == features.shape[0] and rb.out_spatial_shape is not None and rb.subm)
        ctx.save_for_backward(features, filters)
        ext = ref_build.load("sparse_conv_ext")
        if ext is not None:
            out = ext.indice_conv_fp32(features, filters, rb.indice_pairs, rb.indice_pair_num, n_out, 0, ctx.subm)
        else:
            out = torch.from_numpy(ospconv.indice_conv(features.numpy(), filters.numpy(), rb.indice_pairs.numpy(),
                                                       rb.indice_pair_num.numpy(), n_out))
        return out if bias is None else out + bias

    @staticmethod
    def backward(ctx, grad_output):
        features, filters = ctx.saved_tensors
        rb = ctx.rb
        grad_output = grad_output.contiguous()
        ext = ref_build.load("sparse_conv_ext")
        if ext is not None:
            gin, gw = ext.indice_conv_backward_fp32(features, filters, grad_output, rb.indice_pairs,
                                                    rb.indice_pair_num, 0, ctx.subm)
        else:
            gi, gw_ = ospconv.indice_conv_backward(features.numpy(), filters.numpy(), grad_output.numpy(),
                                                   rb.indice_pairs.numpy(), rb.indice_pair_num.numpy())
            gin, gw = torch.from_numpy(gi), torch.from_numpy(gw_)
        return gin, gw, None, None, None


def _cpu_table_conv(features, filters, bias, rb, n_out):
    return _CpuTableConv.apply(features, filters, bias, rb, n_out)


def _cpu_dense(self, channels_first=True):
    shape = [self.batch_size] + list(self.spatial_shape) + [self.features.shape[1]]
    ret = torch.zeros(*shape, dtype=self.features.dtype)
    idx = self.indices.long()
    ret[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
    return ret.permute(0, 4, 1, 2, 3).contiguous() if channels_first else ret


# ---- point ops (numpy restatements of the CUDA kernels; the reference has no CPU version) -------
class _CpuGroup(Function):
    @staticmethod
    def forward(ctx, features, indices):
        ctx.save_for_backward(indices)
        ctx.n = features.shape[2]
        return torch.from_numpy(opointops.grouping_operation(features.detach().numpy(), indices.numpy()))

    @staticmethod
    def backward(ctx, g):
        (indices,) = ctx.saved_tensors
        return torch.from_numpy(opointops.grouping_operation_grad(g.contiguous().numpy(), indices.numpy(), ctx.n)), None


class _CpuGather(Function):
    @staticmethod
    def forward(ctx, features, indices):
        ctx.save_for_backward(indices)
        ctx.n = features.shape[2]
        return torch.from_numpy(opointops.gather_points(features.detach().numpy(), indices.numpy()))

    @staticmethod
    def backward(ctx, g):
        (indices,) = ctx.saved_tensors
        return torch.from_numpy(opointops.gather_points_grad(g.contiguous().numpy(), indices.numpy(), ctx.n)), None


def _cpu_fps(xyz, npoint):
    return torch.from_numpy(opointops.furthest_point_sample(xyz.detach().numpy(), npoint))


def _cpu_ball_query(min_r, max_r, ns, xyz, center):
    return torch.from_numpy(opointops.ball_query(min_r, max_r, ns, xyz.detach().numpy(), center.detach().numpy()))


def _cpu_batch_norm_act(bn, x, residual=None, relu=False):
    """The reference's module chain: nn.BatchNorm1d -> (+ identity) -> ReLU (sparse_block.py:102-120)."""
    y = bn(x)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y


def _cpu_ffn_hidden(linear, dropout, x):
    """The reference's module chain (actr_transformer.py:383-384)."""
    return dropout(F.relu(linear(x)))


def _cpu_add_dropout_layer_norm(norm, dropout, a, b):
    """norm(a + dropout(b)) (actr_transformer.py:385)."""
    return norm(a + dropout(b))


@contextlib.contextmanager
def reference_cpu_ops():
    """Patch the product's op entry points with the reference CPU implementations (from outside)."""
    import ddf_b200.fusion.ms_deform_attn as m_msda
    import ddf_b200.ops.spconv.conv as m_conv
    import ddf_b200.ops.spconv.ops as m_ops
    import ddf_b200.ops.spconv.structure as m_struct
    import ddf_b200.fusion.pointformer as m_pf
    import ddf_b200.ops.pointops as m_po
    import ddf_b200.ops.voxel as m_voxel
    import ddf_b200.ops.sparse_norm as m_norm
    import ddf_b200.ops.fused as m_fused

    def build_rb(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, out_padding, subm, transposed,
                 **_unused):
        rb = cpu_build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, out_padding, subm)
        rb.subm = bool(subm)
        return rb

    saved = [(m_msda, "MSDeformAttnFunction", m_msda.MSDeformAttnFunction),
             (m_ops, "build_rulebook", m_ops.build_rulebook),
             (m_conv.Fsp, "table_conv", m_conv.Fsp.table_conv),
             (m_struct.SparseConvTensor, "dense", m_struct.SparseConvTensor.dense),
             (m_voxel, "voxelization", m_voxel.voxelization),
             (m_voxel, "hard_voxelize_mean", m_voxel.hard_voxelize_mean),
             (m_norm, "batch_norm_act", m_norm.batch_norm_act),
             (m_fused, "ffn_hidden", m_fused.ffn_hidden),
             (m_fused, "add_dropout_layer_norm", m_fused.add_dropout_layer_norm),
             (m_po, "furthest_point_sample", m_po.furthest_point_sample),
             (m_po, "ball_query", m_po.ball_query),
             (m_po, "grouping_operation", m_po.grouping_operation),
             (m_po, "gather_points", m_po.gather_points),
             (m_pf, "gather_points", m_pf.gather_points)]
    try:
        m_msda.MSDeformAttnFunction = _CpuMSDA
        m_ops.build_rulebook = build_rb
        m_conv.Fsp.table_conv = _cpu_table_conv
        m_struct.SparseConvTensor.dense = _cpu_dense
        m_voxel.voxelization = cpu_voxelization
        m_voxel.hard_voxelize_mean = cpu_hard_voxelize_mean
        m_norm.batch_norm_act = _cpu_batch_norm_act
        m_fused.ffn_hidden = _cpu_ffn_hidden
        m_fused.add_dropout_layer_norm = _cpu_add_dropout_layer_norm
        m_po.furthest_point_sample = _cpu_fps
        m_po.ball_query = _cpu_ball_query
        m_po.grouping_operation = _CpuGroup.apply
        m_po.gather_points = _CpuGather.apply
        m_pf.gather_points = _CpuGather.apply
        yield
    finally:
        for obj, name, val in saved:
            setattr(obj, name, val)
