"""TransFusion fusion wrapper (``FUSION_LAYERS['ACTR']``, row a-9) against the REFERENCE class run on a synthetic
nuScenes database (tests/golden/make_wrapper_golden.py): projection through ``lidar2img``, camera assignment,
scale / crop / flip, the zero-padded per-camera layout handed to the encoder (bit-exact), un-pad + fusion."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import detfill  # noqa: E402
import recipes  # noqa: E402

GOLD = np.load(os.path.join(GOLDEN, "wrapper_golden.npz"))

# same specs as tests/golden/make_wrapper_golden.py:CASES
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


def check(name, device, tol):
    from ddf_b200.fusion.point_fusion import ACTR
    spec = CASES[name]
    layer = ACTR(spec["pfat_cfg"], activate_out=spec.get("activate_out", False))
    detfill.fill_state_dict(layer)
    layer = layer.to(device).eval()
    data = recipes.tf_wrapper_case(name, **spec["case"])
    seen = {}
    inner = layer.actr.forward

    def spy(v_feat, grid, i_feats, v_i_feat=None, lidar_grid=None, valid_index=None):
        seen.update(v_feat=v_feat, grid=grid, lidar_grid=lidar_grid, v_i_feat=v_i_feat)
        return inner(v_feat, grid, i_feats, v_i_feat=v_i_feat, lidar_grid=lidar_grid, valid_index=valid_index)
    layer.actr.forward = spy
    with torch.no_grad():
        out = layer([t.to(device) for t in data["img_feats"]], [p.to(device) for p in data["pts"]],
                    data["pts_feats"].to(device), data["img_metas"], None)
    # the padded layout is part of the numerics (GroupNorm sees the padding): same shape, same rows
    for k in ("v_feat", "lidar_grid", "v_i_feat"):
        ref = GOLD["%s/padded_%s" % (name, k)]
        got = seen[k].cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert np.array_equal(got, ref), k
    ref = GOLD[name + "/padded_grid"]
    got = seen["grid"].cpu().numpy()
    # reference points: the reference projects in float64 NumPy, the product in fp32 from the composed matrix
    assert np.abs(got - ref).max() < 2e-5
    ref = GOLD[name + "/out"]
    err = float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err <= tol, err


@pytest.mark.parametrize("name", sorted(CASES))
def test_tf_wrapper_matches_reference_class_cpu(name):
    from oracle import cpu_path
    with cpu_path.reference_cpu_ops():
        check(name, "cpu", 1e-4)   # reference points come from a float64 NumPy chain there, fp32 here


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_tf_wrapper_matches_reference_class_cuda(name):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    check(name, "cuda", 1e-3)
