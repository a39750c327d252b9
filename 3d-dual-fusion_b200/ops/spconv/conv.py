"""Sparse convolution modules with the reference's class names, constructor arguments, parameter
names and weight layout [kd, kh, kw, Cin, Cout] (TransFusion/mmdet3d/ops/spconv/conv.py:48-455)."""
import math

import numpy as np
import torch
from torch.nn import init
from torch.nn.parameter import Parameter

from ...registry import CONV_LAYERS
from . import functional as Fsp
from . import ops
from .modules import SparseModule
from .structure import SparseConvTensor


def _fan_in_hwio(tensor):
    if tensor.dim() < 2:
        raise ValueError("fan in can not be computed for tensor with fewer than 2 dimensions")
    rf = tensor[..., 0, 0].numel() if tensor.dim() > 2 else 1
    return tensor.size(-2) * rf


class _IndiceTuple(object):
    """The reference's ``indice_dict[key]`` 5-tuple (outids, indices, indice_pairs, indice_pair_num,
    spatial_shape; conv.py:181-186) over a rulebook whose pair lists may be built on first access."""

    def __init__(self, rb, indices, spatial_shape):
        self._rb, self._indices, self._shape = rb, indices, spatial_shape

    def _items(self):
        return (self._rb.outids, self._indices, self._rb.indice_pairs, self._rb.indice_pair_num, self._shape)

    def __iter__(self):
        return iter(self._items())

    def __getitem__(self, i):
        return self._items()[i]

    def __len__(self):
        return 5


def conv1x1_only(kernel_size, ndim):
    ks = list(kernel_size) if isinstance(kernel_size, (list, tuple)) else [kernel_size] * ndim
    return all(k == 1 for k in ks)


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, output_padding=0, transposed=False, inverse=False,
                 indice_key=None, fused_bn=False):
        super(SparseConvolution, self).__init__()
        assert groups == 1
        # the names of the reference's 2-D / 4-D / transposed layers resolve (configs that merely import them keep
        # loading), but only what the 3D-DF hot path uses is built: fail at construction, not at the first forward
        if ndim != 3 and not (conv1x1_only(kernel_size, ndim)):
            raise NotImplementedError("%d-D sparse convolutions are not part of the 3D-DF hot path (3-D only)" % ndim)
        if transposed:
            raise NotImplementedError("transposed sparse convolutions are not part of the 3D-DF hot path")
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * ndim
        kernel_size, stride, padding = as_list(kernel_size), as_list(stride), as_list(padding)
        dilation, output_padding = as_list(dilation), as_list(output_padding)
        for d, s in zip(dilation, stride):
            assert any([s == 1, d == 1]), "don't support this."
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.conv1x1 = np.prod(kernel_size) == 1
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = output_padding
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.fused_bn = fused_bn
        self.weight = Parameter(torch.Tensor(*kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(_fan_in_hwio(self.weight))
            init.uniform_(self.bias, -bound, bound)

    def _rulebook(self, input):
        """Find or build the rulebook. Besides the reference's ``indice_key`` sharing
        (conv.py:159-186), convs WITHOUT a key that see the same indices/geometry share one build
        (the reference rebuilds for each of them: 21 builds per SparseEncoder forward where 8 do)."""
        key = self.indice_key
        if key is None:
            key = ("__auto__", input.indices.data_ptr(), input.indices.shape[0],
                   tuple(input.spatial_shape), tuple(self.kernel_size), tuple(self.stride),
                   tuple(self.padding), tuple(self.dilation), self.subm)
        rb = input.indice_dict.get(("__rulebook__", key))
        if rb is None:
            kvol = int(np.prod(self.kernel_size))
            # SubM layers whose wgrad walks the gather table never read the pair lists
            # (host tensors only occur in the oracle-driven CPU path of the tests / the reference arm of bench.py, which
            # must not touch the CUDA library)
            need_pairs = not (self.subm and input.indices.is_cuda
                              and ops.tc_mode(kvol, self.in_channels, self.out_channels) & 8)
            rb = ops.build_rulebook(input.indices, input.batch_size, input.spatial_shape,
                                    self.kernel_size, self.stride, self.padding, self.dilation,
                                    self.output_padding, self.subm, self.transposed, with_pairs=need_pairs)
            # the auto key holds indices.data_ptr(): keep that tensor alive as long as the cached rulebook, or a freed
            # and re-allocated buffer with the same address / row count would hit a stale entry
            rb.key_indices = input.indices
            input.indice_dict[("__rulebook__", key)] = rb
            if self.indice_key is not None:  # the reference's 5-tuple, for code that reads it
                input.indice_dict[self.indice_key] = _IndiceTuple(rb, input.indices, input.spatial_shape)
        return rb

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        features = input.features
        if self.conv1x1:
            features = torch.mm(features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                features = features + self.bias
            out_tensor = SparseConvTensor(features, input.indices, input.spatial_shape, input.batch_size)
            out_tensor.indice_dict = input.indice_dict
            out_tensor.grid = input.grid
            return out_tensor
        if self.inverse:
            datas = input.find_indice_pair(self.indice_key)
            assert datas is not None and self.indice_key is not None
            _, outids, indice_pairs, indice_pair_num, out_spatial_shape = datas
            assert indice_pairs.shape[0] == np.prod(self.kernel_size), \
                "inverse conv must have same kernel size as its couple conv"
            out_features = Fsp.indice_inverse_conv(features, self.weight, indice_pairs, indice_pair_num,
                                                   outids.shape[0])
            if self.bias is not None:
                out_features = out_features + self.bias
        else:
            rb = self._rulebook(input)
            outids, out_spatial_shape = rb.outids, rb.out_spatial_shape
            out_features = Fsp.table_conv(features, self.weight, self.bias, rb, outids.shape[0])
        out_tensor = SparseConvTensor(out_features, outids, out_spatial_shape, input.batch_size)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        return out_tensor


def _make(name, ndim, **fixed):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        SparseConvolution.__init__(self, ndim, in_channels, out_channels, kernel_size, stride, padding,
                                   dilation, groups, bias, indice_key=indice_key, **fixed)
    cls = type(name, (SparseConvolution,), {"__init__": __init__})
    cls.__module__ = __name__
    return CONV_LAYERS.register_module()(cls)


SparseConv2d = _make("SparseConv2d", 2)
SparseConv3d = _make("SparseConv3d", 3)
SparseConv4d = _make("SparseConv4d", 4)
SparseConvTranspose2d = _make("SparseConvTranspose2d", 2, transposed=True)
SparseConvTranspose3d = _make("SparseConvTranspose3d", 3, transposed=True)
SubMConv2d = _make("SubMConv2d", 2, subm=True)
SubMConv3d = _make("SubMConv3d", 3, subm=True)
SubMConv4d = _make("SubMConv4d", 4, subm=True)


def _make_inverse(name, ndim):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        SparseConvolution.__init__(self, ndim, in_channels, out_channels, kernel_size, bias=bias,
                                   inverse=True, indice_key=indice_key)
    cls = type(name, (SparseConvolution,), {"__init__": __init__})
    cls.__module__ = __name__
    return CONV_LAYERS.register_module()(cls)


SparseInverseConv2d = _make_inverse("SparseInverseConv2d", 2)
SparseInverseConv3d = _make_inverse("SparseInverseConv3d", 3)
