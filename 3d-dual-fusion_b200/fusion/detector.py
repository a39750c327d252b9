"""The hot path of ``TransFusionDetector.extract_pts_feat`` (TransFusion/mmdet3d/models/detectors/
transfusion.py:60-108): per-sample hard voxelization -> HardSimpleVFE -> SparseEncoderFusion (with the
3D-DF fusion hook) -> BEV feature map.  BEV backbone, neck and head are out of scope (dense cuDNN /
library code in the reference)."""
import torch
import torch.nn.functional as F
from torch import nn

from ..ops.voxel import Voxelization
from ..registry import build_middle_encoder, build_voxel_encoder
from .voxel_encoder import HardSimpleVFE


class TransFusionPtsBranch(nn.Module):
    def __init__(self, pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder):
        super().__init__()
        self.pts_voxel_layer = Voxelization(**pts_voxel_layer)
        self.pts_voxel_encoder = build_voxel_encoder(pts_voxel_encoder)
        self.pts_middle_encoder = build_middle_encoder(pts_middle_encoder)

    @torch.no_grad()
    def voxelize(self, points):
        voxels, coors, num_points = [], [], []
        for i, res in enumerate(points):
            v, c, n = self.pts_voxel_layer(res)
            voxels.append(v)
            coors.append(F.pad(c, (1, 0), mode="constant", value=i))
            num_points.append(n)
        return torch.cat(voxels, dim=0), torch.cat(num_points, dim=0), torch.cat(coors, dim=0)

    @torch.no_grad()
    def voxelize_mean(self, points):
        """voxelize + HardSimpleVFE in one pass per sample: the padded (M, T, F) tensor is never written."""
        feats, coors = [], []
        for i, res in enumerate(points):
            f, c, _ = self.pts_voxel_layer.forward_mean(res, self.pts_voxel_encoder.num_features)
            feats.append(f)
            coors.append(F.pad(c, (1, 0), mode="constant", value=i))
        return torch.cat(feats, dim=0), torch.cat(coors, dim=0)

    def extract_pts_feat(self, pts, img_feats, img_metas, img=None):
        if type(self.pts_voxel_encoder) is HardSimpleVFE and self.pts_voxel_layer.max_num_points > 0:
            voxel_features, coors = self.voxelize_mean(pts)
        else:
            voxels, num_points, coors = self.voxelize(pts)
            voxel_features = self.pts_voxel_encoder(voxels, num_points, coors)
        batch_size = len(pts)  # reference: coors[-1, 0] + 1 (a D2H read)
        if "Fusion" in self.pts_middle_encoder.__class__.__name__:
            return self.pts_middle_encoder(voxel_features, coors, batch_size, img_feats=img_feats,
                                           img_metas=img_metas, img=img)
        return self.pts_middle_encoder(voxel_features, coors, batch_size)

    forward = extract_pts_feat
