"""``ACTR``: the 3D-DF fusion encoder front (<proj>/models/model_utils/actr.py:40-187) and its
``build`` (:619-657). Hyper-parameters that the reference hides in a DETR argparse namespace
(actr_utils.py:546-644: nheads 8, enc_n_points 4, dim_feedforward 1024, dropout 0.1) are the
defaults of ``ActrArgs`` here."""
import torch
from torch import nn

from ..ops import fused as _fused
from .actr_transformer import build_deformable_transformer
from .position_encoding import (PositionEmbeddingLearnedDepth, PositionEmbeddingSine,
                                PositionEmbeddingSineSparse, PositionEmbeddingSineSparseDepth)


class ActrArgs(object):
    """The subset of the reference's argparse defaults that build() reads."""
    nheads = 8
    dim_feedforward = 1024
    dropout = 0.1
    enc_n_points = 4
    two_stage = False
    num_queries = 300
    hidden_dim = 256
    query_num_feat = 256
    enc_layers = 6
    num_feature_levels = 5
    feature_modal = "lidar"
    hybrid_cfg = None
    gate_first = False


class ACTR(nn.Module):
    def __init__(self, transformer, num_channels, num_feature_levels, max_num_ne_voxel,
                 p_num_channels=None, pos_encode_method="image_coor", feature_modal="lidar"):
        super().__init__()
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.num_feature_levels = num_feature_levels
        self.num_backbone_outs = len(num_channels)
        assert self.num_backbone_outs == num_feature_levels
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim))
            for c in num_channels])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        if feature_modal in ["image", "hybrid"]:
            self.i_input_proj = nn.Sequential(
                nn.Conv1d(num_channels[0], hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim))
            nn.init.xavier_uniform_(self.i_input_proj[0].weight, gain=1)
            nn.init.constant_(self.i_input_proj[0].bias, 0)
        self.feature_modal = feature_modal
        self.max_num_ne_voxel = max_num_ne_voxel
        self.pos_encode_method = pos_encode_method
        assert pos_encode_method in ["image_coor", "depth", "depth_learn"]
        q_model = transformer.q_model
        if pos_encode_method == "image_coor":
            self.q_position_embedding = PositionEmbeddingSineSparse(num_pos_feats=q_model // 2, normalize=True)
        elif pos_encode_method == "depth":
            self.q_position_embedding = PositionEmbeddingSineSparseDepth(num_pos_feats=q_model, normalize=True)
        else:
            self.q_position_embedding = PositionEmbeddingLearnedDepth(num_pos_feats=q_model)
        # kept for attribute parity; its output never reaches the encoder layers (see transformer)
        self.v_position_embedding = PositionEmbeddingSine(num_pos_feats=hidden_dim // 2, normalize=True)

    def forward(self, v_feat, grid, i_feats, v_i_feat=None, lidar_grid=None, valid_index=None):
        """v_feat (B', Lq, C) voxel-query features (zero-padded rows included), grid (B', Lq, 2)
        reference points in [0,1], i_feats list of (B', Cimg, H, W), v_i_feat (B', Lq, Cimg) the
        camera feature under each query, lidar_grid (B', Lq, 3) xyz.  Returns (B', Lq, C).
        ``valid_index`` (ours, optional): flat positions of the real queries in the padded layout; the encoder layers
        then skip the padding (see DeformableTransformerEncoder.forward). The input projections / GroupNorm above the
        encoder always see the padded layout, whose statistics include the padded rows."""
        q_feat = v_feat
        q_i_feat = None
        if self.feature_modal in ["image", "hybrid"]:
            assert v_i_feat is not None
            # Conv1d(k=1) on (B', C, Lq) == Linear on (B', Lq, C): skip the two transposes around it
            conv, gn = self.i_input_proj[0], self.i_input_proj[1]
            q_i_feat = _fused.linear_wb(v_i_feat, conv.weight.squeeze(-1), conv.bias)
            q_i_feat = _fused.group_norm_rows(gn, q_i_feat)       # GroupNorm on (B', C, Lq) without the two transposes
            if self.feature_modal == "image":
                q_feat = q_i_feat
        if self.pos_encode_method == "image_coor":
            q_pos = self.q_position_embedding(grid).transpose(1, 2)
        else:
            q_pos = self.q_position_embedding(lidar_grid[..., 0].clone()).transpose(1, 2)
        srcs = [self._project_map(l, src) for l, src in enumerate(i_feats)]
        return self.transformer(srcs, None, None, q_feat, q_pos, grid, q_lidar_grid=lidar_grid,
                                q_i_feat_flatten=q_i_feat, valid_index=valid_index)


def _project_map(self, lvl, src):
    """input_proj[lvl] = Conv2d(k=1) + GroupNorm on one camera map (actr.py:172-187). On the device the map becomes
    rows [B', H*W, Cin] once, the 1x1 convolution is a row-major GEMM and GroupNorm runs on rows; the result is handed
    on as the NCHW VIEW of those rows, so the transformer's flatten(2).transpose(1, 2) is the rows again (no copy)."""
    proj = self.input_proj[lvl]
    conv, gn = proj[0], proj[1]
    rows_in = isinstance(src, _fused.CameraRows)
    t = src.rows if rows_in else src
    fast = (t.is_cuda and isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
            and conv.padding == (0, 0) and conv.groups == 1 and conv.weight.dtype == torch.float32
            and (rows_in or t.dtype == torch.float32 or not t.requires_grad)
            and (rows_in or t.is_contiguous()) and _fused.group_norm_rows_ok(gn, conv.out_channels))
    if not fast:
        x = src.nchw() if rows_in else src
        return proj(x if x.dtype == conv.weight.dtype else x.to(conv.weight.dtype))
    cam = src if rows_in else _fused.nchw_to_rows(src)
    y = _fused.linear_wb(cam.rows, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias)
    y = _fused.group_norm_rows(gn, y)
    return _fused.CameraRows(y, cam.H, cam.W).nchw()


ACTR._project_map = _project_map


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def build(model_cfg, model_name="ACTR", lt_cfg=None, hybrid_cfg=None, gate_first=False):
    """Same call as the reference ``build`` (actr.py:619; Voxel-RCNN passes ``hybrid_cfg``
    separately, VoxelRCNN/pcdet/models/model_utils/actr.py:623,646)."""
    if model_name not in ("ACTR", "ACTRv2"):
        raise NotImplementedError("%s: only ACTR / ACTRv2 are live in the shipped configs" % model_name)
    args = ActrArgs()
    num_channels = _cfg_get(model_cfg, "num_channels")
    args.query_num_feat = _cfg_get(model_cfg, "query_num_feat")
    args.hidden_dim = _cfg_get(model_cfg, "query_num_feat")
    args.enc_layers = _cfg_get(model_cfg, "num_enc_layers")
    args.pos_encode_method = _cfg_get(model_cfg, "pos_encode_method")
    args.max_num_ne_voxel = _cfg_get(model_cfg, "max_num_ne_voxel")
    args.num_feature_levels = len(num_channels)
    args.feature_modal = _cfg_get(model_cfg, "feature_modal", "lidar")
    args.hybrid_cfg = hybrid_cfg if hybrid_cfg is not None else _cfg_get(model_cfg, "hybrid_cfg", None)
    args.gate_first = gate_first
    transformer = build_deformable_transformer(args, model_name=model_name, lt_cfg=lt_cfg)
    return ACTR(transformer, num_feature_levels=args.num_feature_levels,
                p_num_channels=_cfg_get(model_cfg, "p_num_channels", None), num_channels=num_channels,
                max_num_ne_voxel=args.max_num_ne_voxel, pos_encode_method=args.pos_encode_method,
                feature_modal=args.feature_modal)
