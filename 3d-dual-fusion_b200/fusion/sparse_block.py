"""Sparse conv building blocks (TransFusion/mmdet3d/ops/sparse_block.py:67-185). ``SparseBasicBlock``
keeps mmdet BasicBlock's sub-module names (conv1 / bn1 / conv2 / bn2) so state dicts line up."""
from torch import nn

from ..ops import sparse_norm, spconv
from ..registry import CONV_LAYERS, NORM_LAYERS

NORM_LAYERS.register_module("BN1d", module=nn.BatchNorm1d)
NORM_LAYERS.register_module("BN", module=nn.BatchNorm2d)
NORM_LAYERS.register_module("BN2d", module=nn.BatchNorm2d)


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    layer = CONV_LAYERS.get(cfg.pop("type"))
    return layer(*args, **kwargs, **cfg)


def build_norm_layer(cfg, num_features, postfix=""):
    """mmcv.cnn.build_norm_layer subset: returns (name, module)."""
    cfg = dict(cfg)
    t = cfg.pop("type")
    cfg.pop("requires_grad", None)
    cfg.setdefault("eps", 1e-5)
    layer = NORM_LAYERS.get(t)(num_features, **cfg)
    return "bn" + str(postfix), layer


class SparseBasicBlock(spconv.SparseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
        super().__init__()
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride, padding=1,
                                      dilation=1, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)

    def forward(self, x):
        identity = x.features
        assert x.features.dim() == 2, "x.features.dim()=%d" % x.features.dim()
        out = self.conv1(x)
        # BN -> ReLU and BN -> + identity -> ReLU run as fused kernels on the same BatchNorm1d modules
        out.features = sparse_norm.batch_norm_act(self.norm1, out.features, None, True)
        out = self.conv2(out)
        if self.downsample is not None:
            identity = self.downsample(x)
        out.features = sparse_norm.batch_norm_act(self.norm2, out.features, identity, True)
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0,
                           conv_type="SubMConv3d", norm_cfg=None, order=("conv", "norm", "act")):
    assert isinstance(order, tuple) and len(order) <= 3
    assert set(order) | {"conv", "norm", "act"} == {"conv", "norm", "act"}
    conv_cfg = dict(type=conv_type, indice_key=indice_key)
    layers = []
    for layer in order:
        if layer == "conv":
            if conv_type in ("SparseInverseConv3d", "SparseInverseConv2d", "SparseInverseConv1d"):
                layers.append(build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, bias=False))
            else:
                layers.append(build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                               stride=stride, padding=padding, bias=False))
        elif layer == "norm":
            layers.append(build_norm_layer(norm_cfg, out_channels)[1])
        elif layer == "act":
            layers.append(nn.ReLU(inplace=True))
    return spconv.SparseSequential(*layers)
