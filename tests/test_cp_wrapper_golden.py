"""CenterPoint fusion wrapper (``FUSION['VoxelWithPointProjection']``, 'pfat' mode, with and without the IFAT image
gate) against the REFERENCE classes (tests/golden/make_cp_wrapper_golden.py): per-camera projection with depth thresholds,
integer pixel grids and their rescaling to the feature map, per-(sample, camera) query lists (a voxel seen by two
cameras is a query twice), zero padding, the encoder, one additive update per (voxel, camera). State-dict keys equal."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import detfill  # noqa: E402
import recipes  # noqa: E402

GOLD = np.load(os.path.join(GOLDEN, "cp_wrapper_golden.npz"))
DEPTH_THRES = {"CAM_FRONT": 1, "CAM_FRONT_LEFT": 0, "CAM_FRONT_RIGHT": 0, "CAM_BACK": 0.5, "CAM_BACK_LEFT": 0,
               "CAM_BACK_RIGHT": 0}
# same specs as tests/golden/make_cp_wrapper_golden.py:CASES
CASES = {
    "cp_wrapper_hybrid_ifat": dict(
        pfat_cfg=dict(fusion_method="sum", feature_modal="hybrid",
                      hybrid_cfg=dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"]),
                      num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=26000,
                      pos_encode_method="depth"),
        ifat_cfg=dict(fusion_method="Basicgate_patch_iv_multivoxel", img_num_channel=32, pts_num_channel=64,
                      voxel_feat_channel=[8, 16, 64], voxel_idx=[0, 2])),
    "cp_wrapper_lidar": dict(
        pfat_cfg=dict(fusion_method="sum", num_channels=[32], query_num_feat=64, num_enc_layers=2,
                      max_num_ne_voxel=26000, pos_encode_method="depth"),
        ifat_cfg=None),
}


def check(name, device, tol):
    import ddf_b200.ops.spconv as sp
    from ddf_b200.fusion.centerpoint import VoxelWithPointProjection
    spec = CASES[name]
    fuse = VoxelWithPointProjection("pfat", False, recipes.CP_VOXEL, recipes.CP_RANGE, recipes.CP_CAMS,
                                    image_scale=2.0 / 3, depth_thres=DEPTH_THRES, pfat_cfg=spec["pfat_cfg"],
                                    ifat_cfg=spec["ifat_cfg"])
    assert sorted(fuse.state_dict()) == list(GOLD[name + "/keys"])
    detfill.fill_state_dict(fuse)
    fuse = fuse.to(device).eval()
    data = recipes.cp_wrapper_case(name)
    shapes = ([21, 720, 720], [11, 360, 360], [6, 180, 180])
    tensors = [sp.SparseConvTensor(f.to(device), i.to(device), s, 2) for (i, f), s in zip(data["tensors"], shapes)]
    move = lambda d: {k: ({kk: vv.to(device) for kk, vv in v.items()} if isinstance(v, dict) else v.to(device))
                      for k, v in d.items()}
    bd = dict(calib=move(data["calib"]), image_shape=move(data["image_shape"]),
              img_feat={"layer1_ori_feat2d": move(data["img_feat"]["layer1_ori_feat2d"])})
    with torch.no_grad():
        out = fuse(bd, {}, encoded_voxel_list=tensors, layer_name="layer1_ori", fuse_mode="pfat", d_factor_list=[2, 4, 8])
    ref = GOLD[name + "/features"]
    got = out.features.cpu().numpy()
    assert got.shape == ref.shape
    # the same voxels were touched, then the values
    base = data["tensors"][-1][1].numpy()
    assert np.array_equal(np.abs(got - base).sum(1) > 0, np.abs(ref - base).sum(1) > 0)
    assert float(np.abs(got - ref).max() / np.abs(ref).max()) <= tol


@pytest.mark.parametrize("name", sorted(CASES))
def test_cp_wrapper_matches_reference_class_cpu(name):
    from oracle import cpu_path
    with cpu_path.reference_cpu_ops():
        check(name, "cpu", 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cp_wrapper_matches_reference_class_cuda(name):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    check(name, "cuda", 1e-3)
