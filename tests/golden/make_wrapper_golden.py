"""Golden vectors for the TransFusion fusion wrapper from the REFERENCE class (build container only):
``FUSION_LAYERS['ACTR']`` — TransFusion/mmdet3d/models/fusion_layers/point_fusion.py:315-507 with its
``get_2d_coor_multi`` (:509-549), ``projection`` (:551-643), ``split_param`` (:342-382) and ``agg_param``
(:384-394) — run UNMODIFIED on the CPU around the reference's own encoder (tests/golden/ref_loader.py).

The nuScenes devkit and its database are not in this image. ``projection()`` only needs ``nusc.get(table,
token)`` records, ``pyquaternion.Quaternion(q).rotation_matrix`` and ``nuscenes.utils.geometry_utils.view_points``;
the stand-ins below are a dict-backed ``NuScenes`` over the synthetic database of tests/golden/recipes.py, the
textbook quaternion -> matrix formula and the devkit's documented view_points (K @ points, divide by depth).
The product is driven by ``img_metas['lidar2img']`` = the same lidar -> ego -> global -> ego' -> camera -> pixel
chain composed into one 4x4 per camera (recipes.nusc_database), so this fixture pins the projection rule, the
"last camera wins / unseen -> camera 0 at (0, 0)" assignment, scale / crop / flip handling, the per-camera zero
padded layout and the un-pad + residual sum to the reference's code.
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

CASES = {
    "tf_wrapper_hybrid": dict(
        pfat_cfg=dict(fusion_method="sum", feature_modal="hybrid",
                      hybrid_cfg=dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"]),
                      num_bins=80, num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=26000,
                      pos_encode_method="depth"),
        case=dict()),
    "tf_wrapper_replace_relu": dict(
        pfat_cfg=dict(fusion_method="replace", num_bins=80, num_channels=[32], query_num_feat=64, num_enc_layers=1,
                      max_num_ne_voxel=26000, pos_encode_method="image_coor"),
        activate_out=True,
        case=dict(n_pts=(150, 90, 210), flip=(True, False, False), crop=(None, None, (1.0, 3.0)))),
}


class FakeNuScenes(object):
    def __init__(self, tables):
        self.tables = tables

    def get(self, table, token):
        return self.tables[table][token]


class Quaternion(object):
    def __init__(self, q):
        self.q = np.asarray(q, np.float64)

    @property
    def rotation_matrix(self):
        return recipes.matrix_from_quat(self.q)


def view_points(points, view, normalize):
    """nuscenes.utils.geometry_utils.view_points for a 3x3 intrinsic matrix."""
    viewpad = np.eye(4)
    viewpad[:view.shape[0], :view.shape[1]] = view
    nbr = points.shape[1]
    pts = np.dot(viewpad, np.concatenate((points, np.ones((1, nbr)))))[:3, :]
    if normalize:
        pts = pts / pts[2:3, :].repeat(3, 0).reshape(3, nbr)
    return pts


def load_point_fusion():
    mods = ref_loader.load("TF")
    base = ref_loader.FLAVOURS["TF"][1]

    class _Reg(object):
        def register_module(self, *a, **k):
            return lambda cls: cls

    _ns("mmcv.cnn", xavier_init=lambda *a, **k: None)
    _ns("mmdet3d.models.registry", FUSION_LAYERS=_Reg())
    # no 3-D augmentation in the fixtures: apply_3d_transformation(reverse=True) is the identity
    _ns("mmdet3d.models.fusion_layers", os.path.join(base, "models", "fusion_layers"),
        apply_3d_transformation=lambda pts, coord_type, img_meta, reverse=False: pts)
    _ns("mmdet3d.core")
    _ns("mmdet3d.core.bbox")
    _ns("mmdet3d.core.bbox.structures", get_proj_mat_by_coord_type=lambda meta, coord_type: meta["lidar2img"])
    _ns("nuscenes")
    _ns("nuscenes.utils")
    _ns("nuscenes.utils.geometry_utils", view_points=view_points)
    _ns("nuscenes.nuscenes", NuScenes=FakeNuScenes)
    _ns("pyquaternion", Quaternion=Quaternion)
    pf = importlib.import_module("mmdet3d.models.fusion_layers.point_fusion")
    # projection() does `points.transpose(1, 0).cpu().numpy()` and rotates/translates that array IN PLACE
    # (point_fusion.py:585-611). With CUDA points (the reference's only use) `.cpu()` is a copy; on this CPU run
    # it would alias the caller's tensor and feed camera k the points already moved into camera k-1's frame.
    # Hand every call its own copy = the CUDA behaviour.
    inner = pf.projection
    pf.projection = lambda points, *a, **k: inner(points.clone(), *a, **k)
    return pf, mods


def main():
    pf, _ = load_point_fusion()
    out = {}
    for name, spec in CASES.items():
        torch.manual_seed(0)
        layer = pf.ACTR(ref_loader.AttrDict(spec["pfat_cfg"]), activate_out=spec.get("activate_out", False))
        detfill.fill_state_dict(layer)
        layer.eval()
        data = recipes.tf_wrapper_case(name, **spec["case"])
        layer.nusc = FakeNuScenes(data["tables"])
        seen = {}
        inner = layer.actr.forward

        def spy(v_feat, grid, i_feats, lidar_grid=None, v_i_feat=None, _inner=inner):
            seen.update(v_feat=v_feat.clone(), grid=grid.clone(), lidar_grid=lidar_grid.clone(),
                        v_i_feat=v_i_feat.clone())
            return _inner(v_feat=v_feat, grid=grid, i_feats=i_feats, lidar_grid=lidar_grid, v_i_feat=v_i_feat)
        layer.actr.forward = spy
        with torch.no_grad():
            res = layer(data["img_feats"], data["pts"], data["pts_feats"].clone(), data["img_metas"], None)
        out[name + "/out"] = res.numpy()
        for k, v in seen.items():
            out[name + "/padded_" + k] = v.numpy()
        print(name, res.shape, "padded", tuple(seen["v_feat"].shape),
              "unseen", int((seen["grid"].abs().sum(-1) == 0).sum()))
    np.savez_compressed(os.path.join(HERE, "wrapper_golden.npz"), **out)
    print("wrote wrapper_golden.npz")


if __name__ == "__main__":
    main()
