"""spconv-compatible surface (TransFusion/mmdet3d/ops/spconv/__init__.py; also what
``import spconv`` / ``spconv.pytorch`` offer to the CenterPoint and Voxel-RCNN backbones)."""
from . import functional, ops  # noqa: F401
from .conv import (SparseConv2d, SparseConv3d, SparseConvolution, SparseConvTranspose2d,
                   SparseConvTranspose3d, SparseInverseConv2d, SparseInverseConv3d, SubMConv2d,
                   SubMConv3d)
from .modules import SparseModule, SparseSequential
from .structure import SparseConvTensor, scatter_nd

__all__ = [
    "SparseConv2d", "SparseConv3d", "SubMConv2d", "SubMConv3d", "SparseConvTranspose2d",
    "SparseConvTranspose3d", "SparseInverseConv2d", "SparseInverseConv3d", "SparseModule",
    "SparseSequential", "SparseConvTensor", "scatter_nd", "SparseConvolution",
]
