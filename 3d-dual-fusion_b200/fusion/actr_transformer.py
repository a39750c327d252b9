"""Deformable fusion transformer of 3D-DF (<proj>/models/model_utils/actr_transformer.py):
``DeformableTransformerACTR`` (:22-141), single-query encoder layer (:275-336), dual-query fusion
encoder layer (:338-426; Voxel-RCNN applies the gate BEFORE the FFNs,
VoxelRCNN/pcdet/models/model_utils/actr_transformer.py:497-513), encoder with the optional
3D local self-attention in front of every layer (:428-511)."""
import copy

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import normal_

from ..ops import fused
from ..ops import msda as _msda
from .attentions import attn_dict
from .ms_deform_attn import MSDeformAttn


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def _get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError("activation should be relu/gelu, not {}.".format(activation))


def _with_pos(t, pos):
    return t if pos is None else t + pos


class DeformableTransformerEncoderLayer(nn.Module):
    """Single query stream ('lidar' / 'image' feature_modal)."""

    def __init__(self, d_model=256, q_model=256, d_ffn=1024, dropout=0.1, activation="relu",
                 n_levels=4, n_heads=8, n_points=4, hybrid_cfg=None):
        super().__init__()
        self.d_model = d_model
        self.self_attn = MSDeformAttn(d_model, q_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    with_pos_embed = staticmethod(_with_pos)

    def forward_ffn(self, src):
        if self.activation is F.relu:
            src2 = fused.ffn(self.linear1, self.dropout2, self.linear2, src)     # one kernel forward where supported
        else:
            src2 = fused.linear(self.linear2, self.dropout2(self.activation(self.linear1(src))))
        return fused.add_dropout_layer_norm(self.norm2, self.dropout3, src, src2)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index,
                padding_mask=None, q_pos=None, q_feat=None, q_i_feat=None, plan=None):
        src2 = self.self_attn(_with_pos(q_feat, q_pos), reference_points, src, spatial_shapes,
                              level_start_index, padding_mask, plan=plan)
        q_feat = fused.add_dropout_layer_norm(self.norm1, self.dropout1, q_feat, src2)
        return self.forward_ffn(q_feat), q_i_feat


class DeformableTransformerFusionEncoderLayer(nn.Module):
    """Dual query streams: MSDA updates the image stream, each stream has its FFN, then the
    bi-directional gate mixes them. ``gate_first`` selects the Voxel-RCNN ordering."""

    def __init__(self, d_model=256, q_model=256, d_ffn=1024, dropout=0.1, activation="relu",
                 n_levels=4, n_heads=8, n_points=4, hybrid_cfg=None, gate_first=False):
        super().__init__()
        self.attn_layer = hybrid_cfg["attn_layer"]
        self.q_method = hybrid_cfg.get("q_method", None)
        self.q_rep_place = hybrid_cfg.get("q_rep_place", None)
        self.gate_first = gate_first
        self.d_model = d_model
        self.self_attn = MSDeformAttn(d_model, q_model, n_levels, n_heads, n_points,
                                      q_method=self.q_method, q_rep_place=self.q_rep_place)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        # image-stream FFN
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        # LiDAR-stream FFN
        self.linear3 = nn.Linear(d_model, d_ffn)
        self.dropout4 = nn.Dropout(dropout)
        self.linear4 = nn.Linear(d_ffn, d_model)
        self.dropout5 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.fusion_layer = attn_dict[self.attn_layer](q_model, q_model)

    with_pos_embed = staticmethod(_with_pos)

    def _hidden(self, linear, dropout, src):
        # bias + ReLU + dropout of the [tokens, d_ffn] hidden activation as one in-place pass
        if self.activation is F.relu:
            return fused.ffn_hidden(linear, dropout, src)
        return dropout(self.activation(linear(src)))

    def _ffn(self, linear_a, dropout, linear_b, src):
        if self.activation is F.relu:
            return fused.ffn(linear_a, dropout, linear_b, src)                   # one kernel forward where supported
        return fused.linear(linear_b, self._hidden(linear_a, dropout, src))

    def forward_i_ffn(self, src):
        src2 = self._ffn(self.linear1, self.dropout2, self.linear2, src)
        return fused.add_dropout_layer_norm(self.norm2, self.dropout3, src, src2)

    def forward_p_ffn(self, src):
        src2 = self._ffn(self.linear3, self.dropout4, self.linear4, src)
        return fused.add_dropout_layer_norm(self.norm3, self.dropout5, src, src2)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index,
                padding_mask=None, q_pos=None, q_feat=None, q_i_feat=None, plan=None):
        src2 = self.self_attn(_with_pos(q_feat, q_pos), reference_points, src, spatial_shapes,
                              level_start_index, padding_mask, i_query=_with_pos(q_i_feat, q_pos), plan=plan)
        q_i_feat = fused.add_dropout_layer_norm(self.norm1, self.dropout1, q_i_feat, src2)
        if self.gate_first:
            q_feat, q_i_feat = self.fusion_layer(q_feat, q_i_feat)
            return self.forward_p_ffn(q_feat), self.forward_i_ffn(q_i_feat)
        q_i_feat = self.forward_i_ffn(q_i_feat)
        q_feat = self.forward_p_ffn(q_feat)
        return self.fusion_layer(q_feat, q_i_feat)


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, model_name="ACTR", lt_cfg=None):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.model_name = model_name
        self.spatial_hw = None   # (H, W) of the single feature level, set by DeformableTransformerACTR.forward
        if model_name == "ACTRv2":
            from .pointformer import LocalTransformer
            get = lt_cfg.get if hasattr(lt_cfg, "get") else lambda k, d=None: getattr(lt_cfg, k, d)
            self.lidar_attns = _get_clones(
                LocalTransformer(get("npoint"), get("radius"), get("nsample"), encoder_layer.d_model,
                                 encoder_layer.d_model, num_layers=get("num_layers"),
                                 attn_feat_agg_method=get("attn_feat_agg_method", "unique"),
                                 feat_agg_method=get("feat_agg_method", "replace")), num_layers)

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None,
                padding_mask=None, q_feat=None, q_pos=None, q_reference_points=None,
                q_lidar_grid=None, q_i_feat=None, valid_index=None):
        """``valid_index`` (optional, 1-D long): flat positions (row * Lq + col) of the REAL queries inside the
        zero-padded (B', Lq) layout. Everything the encoder layers do is row-wise (Linear, LayerNorm, FFN, the gate,
        MSDA per query), so with it the layers run on the real rows only and the padded rows of the returned tensor
        are zero - the fusion wrappers drop them anyway (agg_param). Without it (or with the 3D local self-attention,
        whose FPS / ball query see the padding) the padded layout is processed as in the reference."""
        if q_reference_points is None:
            raise NotImplementedError("image->point direction (IACTR) is not part of the 3D-DF hot path")
        reference_points = q_reference_points[:, :, None] * valid_ratios[:, None]
        if valid_index is not None and self.model_name != "ACTRv2":
            out = self._forward_compact(src, pos, reference_points, spatial_shapes, level_start_index, padding_mask,
                                        q_feat, q_pos, q_i_feat, valid_index)
            if out is not None:
                return out
        plan = self._tile_plan(src, reference_points, spatial_shapes)
        geom = None
        for idx, layer in enumerate(self.layers):
            if self.model_name == "ACTRv2":
                lt = self.lidar_attns[idx]
                if geom is None and lt._token_path_ok(q_feat.permute(0, 2, 1)):
                    geom = lt.geometry(q_lidar_grid)   # FPS / ball query / scatter table: once for all layers
                q_feat = lt(q_lidar_grid, q_feat.permute(0, 2, 1), geom)
            q_feat, q_i_feat = layer(src, pos, reference_points, spatial_shapes, level_start_index,
                                     padding_mask, q_pos=q_pos, q_feat=q_feat, q_i_feat=q_i_feat, plan=plan)
        return q_feat

    def _forward_compact(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask, q_feat,
                         q_pos, q_i_feat, valid_index):
        if self._tile_plan(src, reference_points[:1, :1], spatial_shapes) is None:
            return None       # the generic MSDA op needs the regular (N, Lq) layout
        Bp, Lq, C = q_feat.shape
        pick = lambda t: None if t is None else t.reshape(Bp * Lq, -1).index_select(0, valid_index)[None]
        qf, qi, qp = pick(q_feat), pick(q_i_feat), pick(q_pos)
        ref = reference_points.reshape(Bp * Lq, -1, 2).index_select(0, valid_index)
        plan = _msda.TilePlan(ref[:, 0], *self.spatial_hw, query_batch=valid_index // Lq, n_images=Bp)
        ref = ref[None]
        for layer in self.layers:
            qf, qi = layer(src, pos, ref, spatial_shapes, level_start_index, padding_mask, q_pos=qp, q_feat=qf,
                           q_i_feat=qi, plan=plan)
        out = q_feat.new_zeros((Bp * Lq, C)).index_copy(0, valid_index, qf[0])
        return out.view(Bp, Lq, C)

    def _tile_plan(self, src, reference_points, spatial_shapes):
        """One tile binning of the queries for all layers (and their backward): the reference points do not change
        inside the encoder. None when the shape is not the tile kernels' (then every layer takes the generic op)."""
        attn = self.layers[0].self_attn
        if (not src.is_cuda or src.dtype != torch.float32 or reference_points.shape[-1] != 2
                or self.spatial_hw is None
                or not _msda.tile_supported(attn.n_heads, attn.d_model // attn.n_heads, attn.n_levels, attn.n_points)):
            return None
        return _msda.TilePlan(reference_points, *self.spatial_hw)


class DeformableTransformerACTR(nn.Module):
    def __init__(self, d_model=256, query_num_feat=256, nhead=8, num_encoder_layers=6,
                 dim_feedforward=1024, dropout=0.1, activation="relu", return_intermediate_dec=False,
                 num_feature_levels=4, enc_n_points=4, two_stage=False, two_stage_num_proposals=300,
                 model_name="ACTR", lt_cfg=None, feature_modal="lidar", hybrid_cfg=None,
                 gate_first=False):
        super().__init__()
        self.d_model = d_model
        self.q_model = query_num_feat
        self.nhead = nhead
        self.two_stage = two_stage
        self.two_stage_num_proposals = two_stage_num_proposals
        self.feature_modal = feature_modal
        if feature_modal in ["hybrid"]:
            encoder_layer = DeformableTransformerFusionEncoderLayer(
                d_model, self.q_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                enc_n_points, hybrid_cfg, gate_first=gate_first)
        else:
            encoder_layer = DeformableTransformerEncoderLayer(
                d_model, self.q_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                enc_n_points, hybrid_cfg)
        self.encoder = DeformableTransformerEncoder(encoder_layer, num_encoder_layers,
                                                    model_name=model_name, lt_cfg=lt_cfg)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, masks, pos_embeds, q_feat_flatten, q_pos, q_ref_coors, q_lidar_grid=None,
                q_i_feat_flatten=None, valid_index=None):
        """srcs: per-level (B', C, H, W) projected camera maps. ``masks`` may be None (= no padding,
        which is what ACTR always passes: all-False masks, actr.py:172-176) — then valid ratios are 1
        and no masked_fill is issued. ``pos_embeds`` is accepted for signature parity; the encoder
        layers never read it (actr_transformer.py:399-426), so it may be None."""
        src_flatten, spatial_shapes, mask_flatten = [], [], []
        for lvl, src in enumerate(srcs):
            bs, c, h, w = src.shape
            spatial_shapes.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            if masks is not None:
                mask_flatten.append(masks[lvl].flatten(1))
        # ONE token-major copy of the projected camera maps for all encoder layers (value_proj of every layer reads
        # it; each layer used to make its own copy of the transposed view, forward and backward)
        src_flatten = torch.cat(src_flatten, 1) if len(src_flatten) > 1 else src_flatten[0].contiguous()
        self.encoder.spatial_hw = spatial_shapes[0] if len(spatial_shapes) == 1 else None   # python ints: no sync
        device = src_flatten.device
        key = (tuple(spatial_shapes), device)
        cached = getattr(self, "_shape_cache", None)
        if cached is None or cached[0] != key:
            # level shapes / start offsets on the device: the same every step, so built once (no per-forward pageable copy)
            ss = torch.as_tensor(spatial_shapes, dtype=torch.long, device=device)
            lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
            cached = self._shape_cache = (key, ss, lsi)
        spatial_shapes, level_start_index = cached[1], cached[2]
        if masks is not None:
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
            mask_flatten = torch.cat(mask_flatten, 1)
        else:
            valid_ratios = src_flatten.new_ones((src_flatten.shape[0], len(srcs), 2))
            mask_flatten = None
        return self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, None,
                            mask_flatten, q_pos=q_pos, q_feat=q_feat_flatten,
                            q_reference_points=q_ref_coors, q_lidar_grid=q_lidar_grid,
                            q_i_feat=q_i_feat_flatten, valid_index=valid_index)


def build_deformable_transformer(args, model_name="ACTR", lt_cfg=None):
    if "IACTR" in model_name:
        raise NotImplementedError("IACTR* variants are dead code for every shipped 3D-DF config")
    return DeformableTransformerACTR(
        d_model=args.hidden_dim, query_num_feat=args.query_num_feat, nhead=args.nheads,
        num_encoder_layers=args.enc_layers, dim_feedforward=args.dim_feedforward, dropout=args.dropout,
        activation="relu", return_intermediate_dec=True, num_feature_levels=args.num_feature_levels,
        enc_n_points=args.enc_n_points, two_stage=args.two_stage,
        two_stage_num_proposals=args.num_queries, model_name=model_name, lt_cfg=lt_cfg,
        feature_modal=args.feature_modal, hybrid_cfg=args.hybrid_cfg,
        gate_first=getattr(args, "gate_first", False))
