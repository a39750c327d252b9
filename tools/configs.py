"""The reference's model configs for the hot path, as plain dicts (values copied from the config
FILES of the reference, which are the API contract — not code):
  transfusion_f   TransFusion/configs/transfusion_nusc_voxel_F.py:164-227 (pts_voxel_layer,
                  pts_voxel_encoder, pts_middle_encoder incl. fusion_layer)
"""

NUSC_VOXEL_SIZE = [0.075, 0.075, 0.2]
NUSC_PC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]


def transfusion_f(model_name="ACTR"):
    fusion_layer = dict(
        type="ACTR",
        pfat_cfg=dict(
            fusion_method="sum", feature_modal="hybrid",
            hybrid_cfg=dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"]),
            num_bins=80, num_channels=[256], query_num_feat=128, num_enc_layers=2,
            max_num_ne_voxel=26000, pos_encode_method="depth"),
        lt_cfg=dict(npoint=2048, radius=2.0, nsample=32, num_layers=2, attn_feat_agg_method="unique",
                    feat_agg_method="replace"))
    if model_name != "ACTR":
        fusion_layer["model_name"] = model_name
    return dict(
        pts_voxel_layer=dict(max_num_points=10, voxel_size=NUSC_VOXEL_SIZE, max_voxels=(120000, 160000),
                             point_cloud_range=NUSC_PC_RANGE),
        pts_voxel_encoder=dict(type="HardSimpleVFE", num_features=5),
        pts_middle_encoder=dict(
            type="SparseEncoderFusion", in_channels=5, sparse_shape=[41, 1440, 1440], output_channels=128,
            order=("conv", "norm", "act"),
            encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
            encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
            block_type="basicblock", fusion_pos=[3], voxel_size=NUSC_VOXEL_SIZE,
            point_cloud_range=NUSC_PC_RANGE, fusion_layer=fusion_layer))
