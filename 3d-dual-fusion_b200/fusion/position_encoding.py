"""Sine position encodings of the fusion encoder
(<proj>/models/model_utils/position_encoding.py:17-53 dense image map, :56-89 sparse image
coordinates, :91-120 sparse depth, :122-140 learned depth). Parameter-free except the learned one."""
import math

import torch
from torch import nn


def _dim_t(num_pos_feats, temperature, device):
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    return temperature ** (2 * (i // 2) / num_pos_feats)


def _sincos(v, dim_t):
    """v [...]. Returns [..., F]: sin on even feature slots, cos on odd ones (interleaved)."""
    p = v[..., None] / dim_t
    return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)


class PositionEmbeddingSine(nn.Module):
    """Dense (N, H, W) map encoding from a padding mask; takes any object with .tensors/.mask."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, tensor_list):
        x, mask = tensor_list.tensors, tensor_list.mask
        assert mask is not None
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = _dim_t(self.num_pos_feats, self.temperature, x.device)
        pos = torch.cat((_sincos(y_embed, dim_t), _sincos(x_embed, dim_t)), dim=3)
        return pos.permute(0, 3, 1, 2)


class PositionEmbeddingSineSparse(PositionEmbeddingSine):
    """(B, L, 2) normalised image coordinates -> (B, 2F, L)."""

    def forward(self, coor, depth=None):
        assert coor.dtype == torch.float32
        x_embed, y_embed = coor[..., 0], coor[..., 1]
        if self.normalize:
            y_embed = y_embed * self.scale
            x_embed = x_embed * self.scale
        dim_t = _dim_t(self.num_pos_feats, self.temperature, coor.device)
        pos = torch.cat((_sincos(y_embed, dim_t), _sincos(x_embed, dim_t)), dim=2)
        return pos.permute(0, 2, 1)


class PositionEmbeddingSineSparseDepth(PositionEmbeddingSine):
    """(B, L) metric depth -> (B, F, L); depth / 60 * 2*pi (position_encoding.py:105-113)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__(num_pos_feats, temperature, normalize, scale)
        self.norm_param = 60.

    def forward(self, depth):
        d = depth
        if self.normalize:
            d = d / self.norm_param * self.scale
        dim_t = _dim_t(self.num_pos_feats, self.temperature, depth.device)
        return _sincos(d, dim_t).permute(0, 2, 1)


class PositionEmbeddingLearnedDepth(nn.Module):
    """Learned embedding over depth bins (position_encoding.py:122-140): bin = trunc(depth / 60 * num_bin),
    parameter ``d_embed.weight [num_bin, F]``. The reference's ACTR calls it with one argument although the
    class takes ``(feat, depth)`` and so raises (actr.py:161-163); ``feat`` is unused, hence optional here.
    Depths >= 60 m index past the table in the reference (IndexError); they are clamped to the last bin."""

    def __init__(self, num_pos_feats=256, num_bin=120):
        super().__init__()
        self.d_embed = nn.Embedding(num_bin, num_pos_feats)
        self.num_bin = num_bin
        nn.init.uniform_(self.d_embed.weight)

    def forward(self, depth, feat=None):
        idx = (depth / 60. * self.num_bin).to(torch.long).clamp_(0, self.num_bin - 1)
        return self.d_embed(idx).permute(0, 2, 1)
