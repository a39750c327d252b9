"""Golden vectors for the CenterPoint fusion wrapper from the REFERENCE classes (build container only):
``FUSION['VoxelWithPointProjection']`` (CenterPoint/det3d/models/fusion/voxel_with_point_projection.py:14-385, fuse_mode
'pfat') with its ``Point2ImageProjection`` (point_to_image_projection.py:17-231), the IFAT gate
(model_utils/attention.py:9-61) and the reference's own encoder (tests/golden/ref_loader.py), run on the CPU.

Stand-ins: kornia's ``transform_points`` (homogeneous transform, divide by w: the documented behaviour) and
``create_meshgrid3d`` (unused on this path); ``det3d.models.registry.FUSION`` (decorator), ``det3d.core.bbox.box_np_ops``
and ``det3d.datasets.nuscenes.nusc_common`` (imported, unused by the 'pfat' path), ``losses.auxseg_loss`` (segmentation
auxiliary, off); ``Tensor.cuda`` is a no-op and ``Point2ImageProjection``'s default device becomes "cpu".
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import detfill  # noqa: E402
import recipes  # noqa: E402
import ref_loader  # noqa: E402
from ref_loader import _ns  # noqa: E402

DEPTH_THRES = {"CAM_FRONT": 1, "CAM_FRONT_LEFT": 0, "CAM_FRONT_RIGHT": 0, "CAM_BACK": 0.5, "CAM_BACK_LEFT": 0,
               "CAM_BACK_RIGHT": 0}
CASES = {
    # nusc_centerpoint_..._pfat_hybrid7_ifat.py:86-108 (hybrid dual query + IFAT image gate)
    "cp_wrapper_hybrid_ifat": dict(
        pfat_cfg=dict(fusion_method="sum", feature_modal="hybrid",
                      hybrid_cfg=dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"]),
                      num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=26000,
                      pos_encode_method="depth"),
        ifat_cfg=dict(fusion_method="Basicgate_patch_iv_multivoxel", img_num_channel=32, pts_num_channel=64,
                      voxel_feat_channel=[8, 16, 64], voxel_idx=[0, 2])),
    # the plain 'lidar' modal encoder without the gate
    "cp_wrapper_lidar": dict(
        pfat_cfg=dict(fusion_method="sum", num_channels=[32], query_num_feat=64, num_enc_layers=2,
                      max_num_ne_voxel=26000, pos_encode_method="depth"),
        ifat_cfg=None),
}


class SparseStub(object):
    def __init__(self, indices, features):
        self.indices, self.features = indices, features


def transform_points(trans_01, points_1):
    """kornia.geometry.linalg.transform_points: (B, 4, 4) x (B, N, 3) -> (B, N, 3), homogeneous divide."""
    ones = torch.ones_like(points_1[..., :1])
    ph = torch.cat([points_1, ones], -1)
    out = torch.matmul(trans_01, ph.transpose(-1, -2)).transpose(-1, -2)
    return from_homogeneous(out)


def to_homogeneous(points):
    """kornia.geometry.conversions.convert_points_to_homogeneous: append a 1."""
    return torch.nn.functional.pad(points, [0, 1], "constant", 1.0)


def from_homogeneous(points, eps=1e-8):
    """kornia.geometry.conversions.convert_points_from_homogeneous: divide by the last coordinate (1 where it is ~0)."""
    z = points[..., -1:]
    scale = torch.where(z.abs() > eps, 1.0 / z, torch.ones_like(z))
    return scale * points[..., :-1]


def load():
    mods = ref_loader.load("CP")
    base = ref_loader.FLAVOURS["CP"][1]
    torch.Tensor.cuda = lambda self, *a, **k: self

    class _Reg(object):
        def register_module(self, cls):
            return cls

    _ns("kornia")
    _ns("kornia.utils")
    _ns("kornia.utils.grid", create_meshgrid3d=None)
    _ns("kornia.geometry")
    _ns("kornia.geometry.linalg", transform_points=transform_points)
    _ns("kornia.geometry.conversions", convert_points_to_homogeneous=to_homogeneous,
        convert_points_from_homogeneous=from_homogeneous)
    _ns("det3d.models.registry", FUSION=_Reg())
    _ns("det3d.models.losses")
    _ns("det3d.models.losses.auxseg_loss", SEGLOSS=object)
    _ns("det3d.models.utils", os.path.join(base, "models", "utils"))
    _ns("det3d.models.fusion", os.path.join(base, "models", "fusion"))
    _ns("det3d.core")
    _ns("det3d.core.bbox", box_np_ops=types.ModuleType("box_np_ops"))
    _ns("det3d.datasets")
    _ns("det3d.datasets.nuscenes")
    _ns("det3d.datasets.nuscenes.nusc_common", get_lidar2cam_matrix=None, view_points=None)
    p2i = importlib.import_module("det3d.models.fusion.point_to_image_projection")
    p2i.Point2ImageProjection.__init__.__defaults__ = ({}, False, "cpu")
    return importlib.import_module("det3d.models.fusion.voxel_with_point_projection")


def main():
    vw = load()
    out = {}
    for name, spec in CASES.items():
        torch.manual_seed(0)
        fuse = vw.VoxelWithPointProjection("pfat", False, recipes.CP_VOXEL, recipes.CP_RANGE, recipes.CP_CAMS,
                                           image_scale=2.0 / 3, depth_thres=DEPTH_THRES,
                                           pfat_cfg=ref_loader.AttrDict(spec["pfat_cfg"]),
                                           ifat_cfg=ref_loader.AttrDict(spec["ifat_cfg"]) if spec["ifat_cfg"] else None)
        detfill.fill_state_dict(fuse)
        fuse.eval()
        data = recipes.cp_wrapper_case(name)
        tensors = [SparseStub(i.clone(), f.clone()) for i, f in data["tensors"]]
        bd = dict(calib=data["calib"], image_shape=data["image_shape"], img_feat=data["img_feat"])
        with torch.no_grad():
            res = fuse(bd, {}, encoded_voxel_list=tensors, layer_name="layer1_ori", fuse_mode="pfat", d_factor_list=[2, 4, 8])
        out[name + "/features"] = res.features.numpy()
        delta = res.features - data["tensors"][-1][1]
        print(name, tuple(res.features.shape), "rows updated", int((delta.abs().sum(1) > 0).sum()),
              "keys", len(fuse.state_dict()))
        out[name + "/keys"] = np.array(sorted(fuse.state_dict()))
    np.savez_compressed(os.path.join(HERE, "cp_wrapper_golden.npz"), **out)
    print("wrote cp_wrapper_golden.npz")


if __name__ == "__main__":
    main()
