"""Import the REFERENCE's fusion-encoder Python (build container only: needs /root/reference) on the CPU.

The three forks keep the 3D-DF code in ``<pkg>/models/model_utils``. Their imports
need packages that are not in this image (mmcv, mmdet3d / det3d / pcdet with compiled extensions), so the
files are loaded from where they lie with these stand-ins, all registered in ``sys.modules`` BEFORE import:

* ``<pkg>``, ``<pkg>.models``, ``<pkg>.ops``: empty namespace modules whose ``__path__`` is the real
  directory, so the real ``model_utils/*.py`` and the real Python wrappers in ``ops/{ball_query,
  furthest_point_sample,gather_points,group_points}/*.py`` are what runs (their package ``__init__``
  files included), without executing the forks' heavy top-level ``__init__.py``.
* the compiled extensions ``*_ext`` -> the numpy restatements in ``oracle/pointops.py`` (pinned to the
  reference's own known-answer vectors, tests/test_oracle_pointops.py), writing into the caller-allocated
  outputs exactly like the pybind wrappers.
* ``MultiScaleDeformableAttention`` -> empty module; ``MSDeformAttnFunction`` inside
  ``ops/modules/ms_deform_attn.py`` is re-pointed at the reference's own pure-PyTorch
  ``ms_deform_attn_core_pytorch`` (ops/functions/ms_deform_attn_func.py:41-61).
* ``mmcv.cnn.ConvModule`` -> the minimal conv -> norm -> activation module with mmcv's sub-module names
  (``conv``, ``bn``, ``activate``; bias only without a norm layer); ``mmcv.runner.force_fp32`` -> identity.
* ``torch.cuda.IntTensor / FloatTensor`` (used by the wrappers to allocate outputs) -> CPU constructors;
  ``torchvision.__version__`` is read as "1.0" while ``actr_utils.py`` is imported (its ``[:3]`` parse of
  "0.26" would take the torchvision<0.5 branch).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import pointops as opo  # noqa: E402

REFERENCE = "/root/reference"
FLAVOURS = {
    "TF": ("mmdet3d", os.path.join(REFERENCE, "TransFusion", "mmdet3d")),
    "CP": ("det3d", os.path.join(REFERENCE, "CenterPoint", "det3d")),
    "VR": ("pcdet", os.path.join(REFERENCE, "VoxelRCNN", "pcdet")),
}


class AttrDict(dict):
    """mmcv ConfigDict / EasyDict stand-in: attribute access + .get()."""
    __getattr__ = dict.__getitem__


def _ns(name, path=None, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    if path is not None:
        mod.__path__ = [path]
    for k, v in attrs.items():
        setattr(mod, k, v)
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod


class ConvModule(nn.Module):
    """Subset of mmcv.cnn.ConvModule the reference's LocalTransformer uses (order conv, norm, act)."""

    def __init__(self, in_channels, out_channels, kernel_size, norm_cfg=None, act_cfg=dict(type="ReLU"), **kw):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=not self.with_norm)
        if self.with_norm:
            assert norm_cfg["type"] in ("BN2d", "BN")
            self.bn = nn.BatchNorm2d(out_channels)
        if self.with_activation:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def _ext_modules():
    def fps_wrapper(B, N, m, xyz, temp, out):
        out.copy_(torch.from_numpy(opo.furthest_point_sample(xyz.detach().numpy(), m)))

    def fps_with_dist_wrapper(*a):
        raise NotImplementedError("F-FPS is not used by 3D-DF")

    def ball_query_wrapper(B, N, m, min_r, max_r, ns, center, xyz, idx):
        idx.copy_(torch.from_numpy(opo.ball_query(min_r, max_r, ns, xyz.detach().numpy(), center.detach().numpy())))

    def group_forward(B, C, N, npoints, nsample, feats, idx, out):
        out.copy_(torch.from_numpy(opo.grouping_operation(feats.detach().numpy(), idx.numpy())))

    def group_backward(B, C, N, npoints, nsample, grad_out, idx, grad_points):
        grad_points.copy_(torch.from_numpy(opo.grouping_operation_grad(grad_out.contiguous().numpy(), idx.numpy(), N)))

    def gather_wrapper(B, C, N, npoint, feats, idx, out):
        out.copy_(torch.from_numpy(opo.gather_points(feats.detach().numpy(), idx.numpy())))

    def gather_grad_wrapper(B, C, N, npoint, grad_out, idx, grad_feats):
        grad_feats.copy_(torch.from_numpy(opo.gather_points_grad(grad_out.contiguous().numpy(), idx.numpy(), N)))

    return {
        "furthest_point_sample.furthest_point_sample_ext": dict(
            furthest_point_sampling_wrapper=fps_wrapper,
            furthest_point_sampling_with_dist_wrapper=fps_with_dist_wrapper),
        "ball_query.ball_query_ext": dict(ball_query_wrapper=ball_query_wrapper),
        "group_points.group_points_ext": dict(forward=group_forward, backward=group_backward),
        "gather_points.gather_points_ext": dict(gather_points_wrapper=gather_wrapper,
                                                gather_points_grad_wrapper=gather_grad_wrapper),
    }


_loaded = {}


def load(flavour):
    """Returns an AttrDict of the reference modules of one fork: actr, actr_transformer, attentions,
    position_encoding, ms_deform_attn (module file), pointformer."""
    if flavour in _loaded:
        return _loaded[flavour]
    pkg, base = FLAVOURS[flavour]
    if not os.path.isdir(base):
        raise RuntimeError("reference tree not present: %s" % base)

    torch.cuda.IntTensor = torch.IntTensor
    torch.cuda.FloatTensor = torch.FloatTensor

    _ns("mmcv")
    _ns("mmcv.cnn", ConvModule=ConvModule)
    _ns("mmcv.runner", force_fp32=lambda *a, **k: (lambda f: f))
    _ns("MultiScaleDeformableAttention")

    _ns(pkg)
    _ns(pkg + ".models")
    _ns(pkg + ".models.model_utils", os.path.join(base, "models", "model_utils"))
    _ns(pkg + ".ops", os.path.join(base, "ops"))
    _ns(pkg + ".ops.knn", knn=None)
    # pcdet's attentions.py imports pcdet.utils.common_utils for an image-gate class 3D-DF's encoder never builds
    _ns(pkg + ".utils", common_utils=_ns(pkg + ".utils.common_utils"))
    for name, attrs in _ext_modules().items():
        _ns(pkg + ".ops." + name.split(".")[0], os.path.join(base, "ops", name.split(".")[0]))
        # the namespace above shadows the sub-package __init__; import its real python files by name below
        ext = types.ModuleType(pkg + ".ops." + name)
        for k, v in attrs.items():
            setattr(ext, k, v)
        sys.modules[pkg + ".ops." + name] = ext
        setattr(sys.modules[pkg + ".ops." + name.split(".")[0]], name.split(".")[1], ext)
    # sub-package attributes the wrappers import from their parents (`from ..ball_query import ball_query`)
    bq = importlib.import_module(pkg + ".ops.ball_query.ball_query")
    sys.modules[pkg + ".ops.ball_query"].ball_query = bq.ball_query

    import torchvision
    tv_version = torchvision.__version__
    torchvision.__version__ = "1.0.0"
    try:
        mu = pkg + ".models.model_utils."
        mods = AttrDict(
            actr_utils=importlib.import_module(mu + "actr_utils"),
            attentions=importlib.import_module(mu + "attentions"),
            position_encoding=importlib.import_module(mu + "position_encoding"),
            ms_deform_attn_func=importlib.import_module(mu + "ops.functions.ms_deform_attn_func"),
            ms_deform_attn=importlib.import_module(mu + "ops.modules.ms_deform_attn"),
            pointformer=importlib.import_module(mu + "pointformer"),
            actr_transformer=importlib.import_module(mu + "actr_transformer"),
            actr=importlib.import_module(mu + "actr"),
        )
    finally:
        torchvision.__version__ = tv_version

    core = mods.ms_deform_attn_func.ms_deform_attn_core_pytorch

    class _CoreMSDA(object):
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return core(value, shapes, loc, attn)

    mods.ms_deform_attn.MSDeformAttnFunction = _CoreMSDA
    mods.pkg = pkg
    _loaded[flavour] = mods
    return mods
