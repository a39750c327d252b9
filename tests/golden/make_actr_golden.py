"""Golden vectors for the 3D-DF fusion encoder from the REFERENCE's own classes (run in the build
container, where /root/reference exists; CPU, fp32):

  ACTR.forward / build                      <pkg>/models/model_utils/actr.py:40-187,619-657
  DeformableTransformerACTR / Encoder /
  (Fusion)EncoderLayer                      actr_transformer.py:22-141,275-511 (VoxelRCNN: gate before FFN)
  MSDeformAttn (dual query)                 ops/modules/ms_deform_attn.py:33-190
  BiGateSum1D_2                             attentions.py:96-117
  PositionEmbeddingSine*                    position_encoding.py
  LocalTransformer (+ scatter)              pointformer.py:10-44,250-380   (CenterPoint / VoxelRCNN forks)

loaded by tests/golden/ref_loader.py (stand-ins documented there). Weights and inputs are NOT stored: both
sides regenerate them with tests/golden/detfill.py (a pure function of key name / shape), so the fixture
holds the state-dict signature and the outputs only. Each case runs in eval() mode (dropout off, BatchNorm
running statistics) and, when ``train`` is set, again in train() mode with every dropout p forced to 0
(BatchNorm2d of the LocalTransformer position MLP then uses batch statistics over all grouped points,
padded centres included — SURVEY.md section 0.4).

One torch-version shim on the reference side: ``nn.TransformerEncoder.forward`` of torch >= 2.0 passes
``is_causal=`` to its layers, which the reference's ``TransformerEncoderLayerPreNorm.forward`` (written for
torch 1.x) does not accept; the torch-1.x behaviour — apply the layers in order, no final norm — is restored
for that instance.
"""
import json
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import detfill  # noqa: E402
import recipes  # noqa: E402
import ref_loader  # noqa: E402
from ref_loader import AttrDict  # noqa: E402

HYB = dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"])
LT = dict(npoint=24, radius=2.0, nsample=8, num_layers=1, attn_feat_agg_method="unique", feat_agg_method="replace")
LT32 = dict(LT, nsample=32)   # the shipped group size: on CUDA this is the token-major path + csrc/local_attn.cu

# name -> flavour, model_name, ACTR cfg, lt_cfg, hybrid_cfg, (B', Lq, H, W), valid rows per batch row, train too?
CASES = {
    # TransFusion headline (transfusion_nusc_voxel_F.py:203-226): hybrid + ACTR, depth pos-enc, 2 layers
    "tf_hybrid_actr": dict(flavour="TF", model_name="ACTR", train=True,
                           cfg=dict(num_channels=[256], query_num_feat=128, num_enc_layers=2, max_num_ne_voxel=26000,
                                    pos_encode_method="depth", feature_modal="hybrid", hybrid_cfg=HYB),
                           dims=(3, 90, 9, 14), valid=(90, 61, 0)),
    # CenterPoint pfatv2 (nusc_centerpoint_..._pfatv2.py:75-90): lidar + ACTRv2
    "cp_lidar_actrv2": dict(flavour="CP", model_name="ACTRv2", train=True, lt=LT32,
                            cfg=dict(num_channels=[64], query_num_feat=128, num_enc_layers=1, max_num_ne_voxel=26000,
                                     pos_encode_method="depth"),
                            dims=(2, 70, 8, 12), valid=(70, 43)),
    # Voxel-RCNN (voxel_rcnn_car_mm_mvx+actrv2_hybrid_ifat.yaml:53-76): hybrid + ACTRv2, d_model 64, 4 layers,
    # gate BEFORE the FFNs, hybrid_cfg passed separately
    "vr_hybrid_actrv2": dict(flavour="VR", model_name="ACTRv2", train=True, lt=dict(LT32, num_layers=2),
                             cfg=dict(num_channels=[48], query_num_feat=64, num_enc_layers=4, max_num_ne_voxel=20000,
                                      pos_encode_method="depth", feature_modal="hybrid"), hybrid=HYB,
                             dims=(2, 80, 7, 16), valid=(80, 55)),
    # the remaining query-mixing modes of MSDeformAttn (ms_deform_attn.py:129-147)
    "tf_q_image_offset": dict(flavour="TF", model_name="ACTR",
                              cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=100,
                                       pos_encode_method="image_coor", feature_modal="hybrid",
                                       hybrid_cfg=dict(HYB, q_method="image", q_rep_place=["offset"])),
                              dims=(2, 40, 6, 10), valid=(40, 29)),
    "tf_q_gating_both": dict(flavour="TF", model_name="ACTR",
                             cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=100,
                                      pos_encode_method="depth_learn", feature_modal="hybrid",
                                      hybrid_cfg=dict(HYB, q_method="gating", q_rep_place=["offset", "weight"])),
                             dims=(2, 40, 6, 10), valid=(33, 40)),
    "tf_q_sum_offset_gate2": dict(flavour="TF", model_name="ACTR",
                                  cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=100,
                                           pos_encode_method="depth", feature_modal="hybrid",
                                           hybrid_cfg=dict(HYB, attn_layer="BiGate1D_2", q_rep_place=["offset"])),
                                  dims=(2, 40, 6, 10), valid=(40, 17)),
    # the generic LocalTransformer path (group size 8: module graph with nn.MultiheadAttention)
    "cp_lidar_actrv2_ns8": dict(flavour="CP", model_name="ACTRv2", train=True, lt=LT,
                                cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=100,
                                         pos_encode_method="depth"),
                                dims=(2, 70, 6, 10), valid=(70, 43)),
    # single-stream modes
    "tf_image_actr": dict(flavour="TF", model_name="ACTR",
                          cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=100,
                                   pos_encode_method="image_coor", feature_modal="image"),
                          dims=(2, 40, 6, 10), valid=(40, 21)),
    "cp_lidar_actr": dict(flavour="CP", model_name="ACTR",
                          cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=100,
                                   pos_encode_method="depth"),
                          dims=(2, 40, 6, 10), valid=(36, 40)),
    # the remaining gate blocks (attentions.py:27-94)
    "tf_gate_bigate1d": dict(flavour="TF", model_name="ACTR",
                             cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=100,
                                      pos_encode_method="depth", feature_modal="hybrid",
                                      hybrid_cfg=dict(HYB, attn_layer="BiGate1D")),
                             dims=(2, 40, 6, 10), valid=(40, 17)),
    "tf_gate_bigatesum1d": dict(flavour="TF", model_name="ACTR",
                                cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=2, max_num_ne_voxel=100,
                                         pos_encode_method="depth", feature_modal="hybrid",
                                         hybrid_cfg=dict(HYB, attn_layer="BiGateSum1D")),
                                dims=(2, 40, 6, 10), valid=(40, 17)),
    # the other scatter / aggregation rule of the LocalTransformer (pointformer.py:331-347,371-376). Only the
    # Voxel-RCNN fork's version runs: TransFusion / CenterPoint divide an [C, n_hit] slice by the whole bincount
    # (pointformer.py:345) and raise unless every voxel below the largest grouped index is hit
    "vr_actrv2_sum_sum": dict(flavour="VR", model_name="ACTRv2",
                              lt=dict(LT, attn_feat_agg_method="sum", feat_agg_method="sum"),
                              cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=100,
                                       pos_encode_method="depth"),
                              dims=(2, 60, 6, 10), valid=(60, 41)),
    "vr_actrv2_sum_replace": dict(flavour="VR", model_name="ACTRv2",
                                  lt=dict(LT, attn_feat_agg_method="sum", feat_agg_method="replace"),
                                  cfg=dict(num_channels=[32], query_num_feat=64, num_enc_layers=1, max_num_ne_voxel=100,
                                           pos_encode_method="depth"),
                                  dims=(2, 60, 6, 10), valid=(60, 41)),
}


def make_inputs(name, cfg_or_case, dims=None, valid=None):
    if dims is None:
        cfg_or_case, dims, valid = cfg_or_case["cfg"], cfg_or_case["dims"], cfg_or_case["valid"]
    return recipes.actr_inputs(name, cfg_or_case, dims, valid)


def build_reference(case):
    mods = ref_loader.load(case["flavour"])
    cfg = AttrDict(case["cfg"])
    lt = AttrDict(case["lt"]) if case.get("lt") else None
    if case["flavour"] == "VR":
        net = mods.actr.build(cfg, model_name=case["model_name"], lt_cfg=lt, hybrid_cfg=case.get("hybrid"))
    else:
        net = mods.actr.build(cfg, model_name=case["model_name"], lt_cfg=lt)
    if case["cfg"]["pos_encode_method"] == "depth_learn":
        # the reference calls PositionEmbeddingLearnedDepth.forward(feat, depth) with ONE argument
        # (actr.py:161-163 vs position_encoding.py:134) and raises TypeError; `feat` is unused by the class, so the
        # evident intent — forward(None, depth) — is what is pinned here
        pe = net.q_position_embedding
        pe.forward = (lambda depth, _f=pe.forward: _f(None, depth))
    for m in net.modules():
        if isinstance(m, nn.TransformerEncoder):
            def old_forward(src, _m=m):
                for layer in _m.layers:
                    src = layer(src)
                return src
            m.forward = old_forward
    return net


def run(net, inputs, feature_modal):
    v_feat, grid, i_feat, v_i_feat, lidar = [t.clone() for t in inputs]
    with torch.no_grad():
        return net(v_feat, grid, [i_feat], v_i_feat if feature_modal in ("image", "hybrid") else None, lidar)


def main():
    out, meta = {}, {}
    for name, case in CASES.items():
        torch.manual_seed(0)
        net = build_reference(case)
        detfill.fill_state_dict(net)
        inputs = make_inputs(name, case["cfg"], case["dims"], case["valid"])
        modal = case["cfg"].get("feature_modal", "lidar")
        net.eval()
        out[name + "/eval"] = run(net, inputs, modal).numpy()
        if case.get("train"):
            for m in net.modules():
                if isinstance(m, nn.Dropout):
                    m.p = 0.0
                if isinstance(m, nn.MultiheadAttention):
                    m.dropout = 0.0
            net.train()
            out[name + "/train"] = run(net, inputs, modal).numpy()
        meta[name] = dict(case, signature=[[k, list(s)] for k, s in detfill.state_dict_signature(net)])
        print(name, out[name + "/eval"].shape, float(np.abs(out[name + "/eval"]).mean()),
              len(meta[name]["signature"]), "keys")
    np.savez_compressed(os.path.join(HERE, "actr_golden.npz"), **out)
    with open(os.path.join(HERE, "actr_golden.json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)
    print("wrote actr_golden.npz / actr_golden.json")


if __name__ == "__main__":
    main()
