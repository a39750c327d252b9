"""3D local self-attention over voxel queries: ``LocalTransformer``
(<proj>/models/model_utils/pointformer.py:250-380) and its pre-norm encoder layer (:10-44).
D-FPS picks ``npoint`` centres, a ball query groups ``nsample`` neighbours, a small MLP encodes the
(absolute) neighbour coordinates, a pre-norm transformer attends inside each group and the result
is written back to the voxels ("unique"/"replace": the FIRST grouped copy of a voxel in flattened
(group, slot) order wins — the reference's scatter_ with duplicate indices is nondeterministic on
CUDA; first occurrence is what it yields on CPU and is the contract here)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops import fused as _fused
from ..ops import pointops as _pointops
from ..ops import sparse_norm as _sparse_norm
from ..ops.pointops import Points_Sampler, QueryAndGroup, gather_points


class ConvModule(nn.Module):
    """The subset of mmcv.cnn.ConvModule LocalTransformer uses: Conv2d (+ BN2d) (+ ReLU), with
    mmcv's sub-module names ``conv`` / ``bn`` and bias only when there is no norm."""

    def __init__(self, in_channels, out_channels, kernel_size, norm_cfg=None, act_cfg=dict(type="ReLU")):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=norm_cfg is None)
        self.with_norm = norm_cfg is not None
        if self.with_norm:
            self.bn = nn.BatchNorm2d(out_channels)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            self.activate = nn.ReLU(inplace=True)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class TransformerEncoderLayerPreNorm(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu"):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout, inplace=True)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout, inplace=True)
        self.dropout2 = nn.Dropout(dropout, inplace=True)
        self.activation = nn.ReLU(inplace=True)

    def forward(self, src, src_mask=None, src_key_padding_mask=None, **kwargs):
        src = self.norm1(src)
        src2, _ = self.self_attn(src, src, src, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)
        src = src + self.dropout1(src2)
        src = self.norm2(src)
        src2 = self.linear2(self.dropout(self.activation(self.linear1(src))))
        return src + self.dropout2(src2)


class _EncoderStack(nn.Module):
    """nn.TransformerEncoder's state-dict layout (``layers.{i}.*``) without its fast-path checks."""

    def __init__(self, layer, num_layers):
        super().__init__()
        import copy
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    def forward(self, src):
        for layer in self.layers:
            src = layer(src)
        return src


def first_occurrence_scatter(attn_features, feats, idxs):
    """For every voxel that appears in ``idxs`` (B, np, ns), take the grouped feature of its first occurrence in
    flattened order; returns the updated (B, C, N) tensor. CUDA: two device kernels (atomicMin on the flattened
    position, then a column gather; ops.pointops.scatter_first). Host tensors (the oracle-driven CPU path of the
    tests): per-row scatter_reduce, updating ``attn_features`` in place like the reference."""
    if attn_features.is_cuda:
        return _pointops.scatter_first(attn_features, feats, idxs)
    B, C, N = attn_features.shape
    for b in range(B):
        idx_f = idxs[b].reshape(-1).long()
        feat_f = feats[b].reshape(C, -1)
        pos = torch.arange(idx_f.numel(), device=idx_f.device)
        first = torch.full((N,), idx_f.numel(), dtype=torch.long, device=idx_f.device)
        first.scatter_reduce_(0, idx_f, pos, reduce="amin", include_self=True)
        hit = first < idx_f.numel()
        attn_features[b][:, hit] = feat_f[:, first[hit]]
    return attn_features


class LocalTransformer(nn.Module):
    def __init__(self, npoint, radius, nsample, dim_feature, dim_out, nhead=4, num_layers=2,
                 norm_cfg=dict(type="BN2d"), ratio=1, drop=0.0, prenorm=True,
                 attn_feat_agg_method="unique", feat_agg_method="replace"):
        super().__init__()
        assert ratio == 1 and prenorm, "only the configuration 3D-DF uses is implemented"
        self.npoint = npoint
        self.nsample = nsample
        self.radius = radius
        self.nc_in = dim_feature
        self.nc_out = dim_out
        self.sampler = Points_Sampler([self.npoint], ["D-FPS"])
        self.grouper = QueryAndGroup(self.radius, self.nsample, use_xyz=False, return_grouped_xyz=True,
                                     return_grouped_idx=True, normalize_xyz=False)
        self.pe = nn.Sequential(ConvModule(3, self.nc_in // 2, 1, norm_cfg=norm_cfg),
                                ConvModule(self.nc_in // 2, self.nc_in, 1, act_cfg=None, norm_cfg=None))
        self.chunk = _EncoderStack(
            TransformerEncoderLayerPreNorm(d_model=self.nc_in, dim_feedforward=2 * self.nc_in, dropout=drop,
                                           nhead=nhead), num_layers)
        self.attn_feat_agg_method = attn_feat_agg_method
        self.feat_agg_method = feat_agg_method

    def scatter(self, attn_features, feats, idxs):
        if self.attn_feat_agg_method == "unique":
            return first_occurrence_scatter(attn_features, feats, idxs)
        elif self.attn_feat_agg_method == "sum":
            B, C, N = attn_features.shape
            for b in range(B):
                idx_f = idxs[b].reshape(-1).long()
                summed = torch.zeros_like(attn_features[b]).index_add_(1, idx_f, feats[b].reshape(C, -1))
                cnt = torch.bincount(idx_f, minlength=N)
                nz = cnt > 0
                attn_features[b][:, nz] = (attn_features[b][:, nz] + summed[:, nz]) / cnt[nz]
            return attn_features
        else:
            raise NotImplementedError(self.attn_feat_agg_method)

    # ---- CUDA hot path: everything token-major --------------------------------------------------------------
    def _token_path_ok(self, features):
        layer = self.chunk.layers[0]
        mha = layer.self_attn
        C = features.shape[1]
        return (features.is_cuda and features.dtype == torch.float32 and self.attn_feat_agg_method == "unique"
                and self.feat_agg_method == "replace" and mha._qkv_same_embed_dim and mha.in_proj_bias is not None
                and (mha.dropout == 0.0 or not self.training)
                and _pointops.local_attn_supported(mha.num_heads, C // mha.num_heads, self.nsample))

    def geometry(self, xyz):
        """Everything of the LocalTransformer that depends on the coordinates only: D-FPS centres, ball-query groups,
        the flattened row index of every grouped token, the grouped (absolute) coordinates and the first-occurrence
        table of the scatter. The encoder hands the same ``xyz`` to every layer's LocalTransformer, which the reference
        re-samples and re-groups each time (actr_transformer.py:496-498); here it is computed once per forward."""
        xyz = xyz.contiguous()
        B, N, _ = xyz.shape
        fps_idx = self.sampler(xyz, None)
        new_xyz = gather_points(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
        idx = _pointops.ball_query(0, self.radius, self.nsample, xyz, new_xyz)         # (B, np, ns) int32
        flat = (idx.long() + torch.arange(B, device=idx.device)[:, None, None] * N).reshape(-1)
        gxyz = xyz.reshape(B * N, 3).index_select(0, flat)                            # (T, 3) absolute coordinates
        first = _pointops.first_occurrence(idx, N).long()                              # (B, N), E where never grouped
        E = idx.shape[1] * idx.shape[2]
        hit = first < E
        src = (first.clamp(max=E - 1) + torch.arange(B, device=idx.device)[:, None] * E).reshape(-1)
        return dict(flat=flat, gxyz=gxyz, hit=hit.reshape(-1, 1), src=src, xyz=xyz)

    def _layer_tokens(self, layer, x):
        """TransformerEncoderLayerPreNorm.forward (pointformer.py:33-44) on (T, C) tokens; groups of ``nsample``
        consecutive rows are the sequences."""
        mha = layer.self_attn
        s = _fused.add_dropout_layer_norm(layer.norm1, None, x, None)       # plain LayerNorm, one warp-per-row pass
        qkv = _fused.linear_wb(s, mha.in_proj_weight, mha.in_proj_bias)
        o = _pointops.local_attention(qkv, mha.num_heads, self.nsample)
        s = s + layer.dropout1(_fused.linear_wb(o, mha.out_proj.weight, mha.out_proj.bias))
        s = _fused.add_dropout_layer_norm(layer.norm2, None, s, None)
        return s + layer.dropout2(_fused.ffn(layer.linear1, layer.dropout, layer.linear2, s))

    def forward_tokens(self, xyz, feats_nc, geom=None):
        """feats_nc (B, N, C) row-major voxel features -> (B, N, C). No (B, C, np, ns) tensor, no permutes: the grouped
        tokens are gathered as (T, C) rows, the position MLP (1x1 convs + BN2d = Linear + BatchNorm over the same T
        samples) and the transformer layers are row-major GEMMs + the fused kernels, the 32-token attention is
        csrc/local_attn.cu, the write-back is a row gather through the first-occurrence table."""
        B, N, C = feats_nc.shape
        if geom is None:
            geom = self.geometry(xyz)
        rows = feats_nc.reshape(B * N, C)
        x = rows.index_select(0, geom["flat"])
        conv1, bn, conv2 = self.pe[0].conv, self.pe[0].bn, self.pe[1].conv
        pe = F.linear(geom["gxyz"], conv1.weight.view(conv1.out_channels, 3))
        pe = _sparse_norm.batch_norm_act(bn, pe, relu=True)        # BatchNorm2d over (B, H, W) == over the T rows
        x = x + _fused.linear_wb(pe, conv2.weight.view(conv2.out_channels, conv2.in_channels), conv2.bias)
        for layer in self.chunk.layers:
            x = self._layer_tokens(layer, x)
        out = torch.where(geom["hit"], x.index_select(0, geom["src"]), rows)
        return out.view(B, N, C)

    def forward(self, xyz, features, geom=None):
        """xyz (B, N, 3), features (B, C, N) -> (B, N, C). Mutates ``features`` like the reference."""
        if self._token_path_ok(features):
            return self.forward_tokens(xyz, features.transpose(1, 2).contiguous(), geom)
        xyz = xyz.contiguous()
        fps_idx = self.sampler(xyz, features)
        new_xyz = gather_points(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2)
        group_features, group_xyz, group_idx = self.grouper(xyz, new_xyz.contiguous(), features.contiguous())
        input_features = group_features + self.pe(group_xyz)
        B, D, n_p, ns = input_features.shape
        tokens = input_features.permute(0, 2, 1, 3).reshape(-1, D, ns).permute(2, 0, 1)
        transformed = self.chunk(tokens).permute(1, 2, 0).reshape(B, n_p, D, ns).transpose(1, 2)
        if self.feat_agg_method == "replace":
            features = features.clone() if (features.requires_grad and not features.is_cuda) else features
            features = self.scatter(features, transformed, group_idx)
        elif self.feat_agg_method == "sum":
            attn_features = self.scatter(torch.zeros_like(features), transformed, group_idx)
            features = features + attn_features
        else:
            raise NotImplementedError(self.feat_agg_method)
        return features.permute(0, 2, 1)
