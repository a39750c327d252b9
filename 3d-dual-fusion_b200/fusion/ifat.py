"""IFAT / AGFN image-side gate: voxel features are scattered into the camera feature
plane and gate the image features that the 3D-DF encoder then samples.

Reference: ``Basicgate_patch_iv_multivoxel`` (CenterPoint/det3d/models/model_utils/attention.py:8-61) and
``pts2img`` (:422-466), called once per (camera, sample) from
CenterPoint/det3d/models/fusion/voxel_with_point_projection.py:277-293.  Same parameters and state-dict
keys (``reduced_dim.{i}``, ``reduced_dim2``, ``reduced_dim3``, ``spatial_basic``); here all (sample,
camera) planes go through the convolutions as one batch, and the scatter is deterministic: where several
voxels of a scale fall on one pixel the LAST one in query order wins (the reference's CPU semantics of
``i_pts_feat[y, x] = pts_feat``; its CUDA index_put picks an arbitrary one).
"""
import torch
from torch import nn


def pts2img(cell, feats, n_cells):
    """Scatter ``feats`` (n, C) into ``n_cells`` pixels (flat index ``cell`` (n,)); last query wins.
    Returns (n_cells, C), zeros where no voxel projects."""
    n = cell.shape[0]
    out = feats.new_zeros((n_cells, feats.shape[1]))
    if n == 0:
        return out
    winner = torch.full((n_cells,), -1, dtype=torch.long, device=cell.device)
    winner.scatter_reduce_(0, cell, torch.arange(n, device=cell.device), "amax", include_self=True)
    hit = winner >= 0
    out[hit] = feats[winner[hit]]
    return out


class Basicgate_patch_iv_multivoxel(nn.Module):
    def __init__(self, **kwarg):
        super(Basicgate_patch_iv_multivoxel, self).__init__()
        self.img_num_channel = kwarg["img_num_channel"]
        self.pts_num_channel = kwarg["pts_num_channel"] + 3
        self.voxel_feat_channel = list(kwarg["voxel_feat_channel"])
        self.voxel_idx = list(kwarg["voxel_idx"])
        if len(self.voxel_idx) == 1:
            self.pts_num_channel = self.voxel_feat_channel[self.voxel_idx[0]]
        last = self.voxel_feat_channel[self.voxel_idx[-1]] + 3
        self.reduced_dim2 = nn.Conv2d(last, last, kernel_size=1, stride=1, padding=0)
        self.reduced_dim3 = nn.Conv2d(self.img_num_channel, 1, kernel_size=1, stride=1, padding=0)
        self.spatial_basic = nn.Conv2d(last, 1, kernel_size=3, stride=1, padding=1)
        self.reduced_dim = nn.Sequential(*[
            nn.Conv2d(self.voxel_feat_channel[i] + 3, last, kernel_size=1, stride=1, padding=0)
            for i in range(self.voxel_idx[-1])])

    def forward(self, img_feat, voxel_feat, cell, voxel_coord):
        """img_feat (G, C_img, H, W): the camera planes of all (sample, camera) groups.
        voxel_feat[s] (n_s, C_s), cell[s] (n_s,) flat pixel index g * H * W + y * W + x, voxel_coord[s]
        (n_s, 3), one entry per backbone scale s (only the scales in ``voxel_idx`` are read)."""
        G, _, H, W = img_feat.shape
        pt_img = None
        for s in self.voxel_idx:
            f = torch.cat([voxel_feat[s], voxel_coord[s]], dim=-1)
            plane = pts2img(cell[s], f, G * H * W).view(G, H, W, -1).permute(0, 3, 1, 2).contiguous()
            if s != self.voxel_idx[-1]:
                plane = self.reduced_dim[s](plane)
            pt_img = plane if pt_img is None else pt_img + plane
        pt_img = self.reduced_dim2(pt_img)
        gated = self.reduced_dim3(img_feat)                 # (G, 1, H, W), broadcast over the channels
        attention_map = torch.sigmoid(self.spatial_basic(gated + pt_img))
        return img_feat * attention_map


__all__ = {"Basicgate_patch_iv_multivoxel": Basicgate_patch_iv_multivoxel}
