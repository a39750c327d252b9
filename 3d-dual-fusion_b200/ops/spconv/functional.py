"""autograd Functions with the reference's names (TransFusion/mmdet3d/ops/spconv/functional.py:20-98)
plus the table-driven Function the SparseConvolution module uses on the hot path."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class _PairConv(Function):
    INVERSE = False
    SUBM = False

    @classmethod
    def _fwd(cls, ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                               cls.INVERSE, cls.SUBM)

    @classmethod
    def _bwd(cls, ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output, indice_pairs,
                                                        indice_pair_num, cls.INVERSE, cls.SUBM)
        return input_bp, filters_bp, None, None, None


class SparseConvFunction(_PairConv):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SparseConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num,
                                       num_activate_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return SparseConvFunction._bwd(ctx, grad_output)


class SparseInverseConvFunction(_PairConv):
    INVERSE = True

    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SparseInverseConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num,
                                              num_activate_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return SparseInverseConvFunction._bwd(ctx, grad_output)


class SubMConvFunction(_PairConv):
    SUBM = True

    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SubMConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num,
                                     num_activate_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return SubMConvFunction._bwd(ctx, grad_output)


class TableConvFunction(Function):
    """Hot-path conv: forward through the gather table, dgrad through the scatter table, wgrad over
    the pair lists. ``bias`` may be None."""

    @staticmethod
    def forward(ctx, features, filters, bias, rulebook, num_activate_out):
        features = features.contiguous()
        filters = filters.contiguous()
        cin, cout = filters.shape[-2], filters.shape[-1]
        ctx.cin = cin
        kvol = rulebook.gather_table.shape[1] if rulebook.gather_table is not None else rulebook.indice_pairs.shape[0]
        pad = ops.padded_cin(kvol, cin, cout) - cin
        if pad:
            # e.g. the 5-channel input layer: zero-pad the contraction to a multiple of 8 so the rows
            # are 16-byte aligned and the layer runs on the tensor-core kernels like all the others
            features = torch.nn.functional.pad(features, (0, pad))
            filters = torch.nn.functional.pad(filters, (0, 0, 0, pad))
            cin += pad
        ctx.mode = mode = ops.tc_mode(kvol, cin, cout)
        saved = features
        fmt = 0
        if features.shape[0]:
            if mode & 16:
                # bf16x3 forward: hi/lo split operands; the same pass writes the tf32-rounded copy wgrad reads.  A
                # BatchNorm that produced these features has written both already (ops/sparse_norm.py)
                pre = getattr(features, "_ddf_operands", None) if not pad else None
                want_rounded = bool(mode & 12) and ctx.needs_input_grad[1]
                if pre is not None and pre[0].shape == features.shape and (pre[1] is not None or not want_rounded):
                    features, rounded = pre[0], (pre[1] if want_rounded else None)
                else:
                    features, rounded = ops.split_bf16x3(features, want_rounded=bool(mode & 12))
                saved = rounded if rounded is not None else saved
                fmt = 1
            elif mode & 5:
                # tensor-core operands are made exact tf32 once; forward and wgrad share the copy. A forward
                # that runs in fp32 (narrow layers) keeps the unrounded features.
                saved = ops.round_tf32(features)
                if mode & 1:
                    features = saved
        ctx.rulebook = rulebook
        ctx.has_bias = bias is not None
        ctx.save_for_backward(saved, filters)
        return ops.sparse_conv_forward(features, filters, rulebook.gather_table, bias, num_activate_out, fmt)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        features, filters = ctx.saved_tensors
        rb = ctx.rulebook
        grad_output = grad_output.contiguous()
        gb = grad_output.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        grad_exact = grad_output
        grad_split = None
        if grad_output.shape[0]:
            if ctx.mode & 32 and ctx.needs_input_grad[0]:
                grad_split, rounded = ops.split_bf16x3(grad_output, want_rounded=bool(ctx.mode & 12)
                                                       and ctx.needs_input_grad[1])
                grad_output = rounded if rounded is not None else grad_output
            elif ctx.mode & 14:
                grad_output = ops.round_tf32(grad_output)  # shared by dgrad and wgrad
        gin = gw = None
        if ctx.needs_input_grad[0]:
            if grad_split is not None:
                gin = ops.sparse_conv_dgrad(filters, grad_split, rb.scatter_table, features.shape[0], 1)
            else:
                gin = ops.sparse_conv_dgrad(filters, grad_output if ctx.mode & 2 else grad_exact, rb.scatter_table,
                                            features.shape[0])
        if ctx.needs_input_grad[1]:
            if ctx.mode & 8 and rb.gather_table is not None and grad_output.shape[0]:
                # Cin == Cout layers (SubM or not: the gather table of a strided / (3,1,1) conv is the same object):
                # walk the output rows once through the gather table
                gw = ops.sparse_conv_wgrad_table(features, filters, grad_output, rb.gather_table)
            else:
                gw = ops.sparse_conv_wgrad(features, filters, grad_output, rb.indice_pairs, rb.indice_pair_num)
        if ctx.cin != filters.shape[-2]:  # drop the zero-padded input channels again
            gin = gin[:, :ctx.cin].contiguous() if gin is not None else None
            gw = gw[..., :ctx.cin, :].contiguous() if gw is not None else None
        return gin, gw, gb, None, None


indice_conv = SparseConvFunction.apply
indice_inverse_conv = SparseInverseConvFunction.apply
indice_subm_conv = SubMConvFunction.apply
table_conv = TableConvFunction.apply
