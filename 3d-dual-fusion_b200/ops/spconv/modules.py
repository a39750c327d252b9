"""SparseModule / SparseSequential (TransFusion/mmdet3d/ops/spconv/modules.py:46-137): a container
that feeds SparseConvTensor to sparse modules and ``.features`` to ordinary nn.Modules."""
from collections import OrderedDict

import torch
from torch import nn

from .structure import SparseConvTensor


class SparseModule(nn.Module):
    """Marker base class: subclasses take and return SparseConvTensor."""
    pass


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def is_sparse_conv(module):
    from .conv import SparseConvolution
    return isinstance(module, SparseConvolution)


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super(SparseSequential, self).__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self._sparity_dict = {}

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    @property
    def sparity_dict(self):
        return self._sparity_dict

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        from .. import sparse_norm
        mods = list(self._modules.items())
        skip = False
        for i, (k, module) in enumerate(mods):
            if skip:
                skip = False
                continue
            if (isinstance(module, nn.BatchNorm1d) and isinstance(input, SparseConvTensor)
                    and input.indices.shape[0] != 0):
                # BatchNorm1d [-> ReLU] on sparse features: one fused op (same modules, same state dict)
                skip = i + 1 < len(mods) and isinstance(mods[i + 1][1], nn.ReLU)
                input.features = sparse_norm.batch_norm_act(module, input.features, None, skip)
                continue
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                self._sparity_dict[k] = input.sparity
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
        return input
