"""mmdet3d ``VOXEL_ENCODERS['HardSimpleVFE']`` (TransFusion/mmdet3d/models/voxel_encoders/
voxel_encoder.py:13-44): mean of the points of a voxel."""
from torch import nn

from ..registry import VOXEL_ENCODERS


@VOXEL_ENCODERS.register_module()
class HardSimpleVFE(nn.Module):
    def __init__(self, num_features=4):
        super(HardSimpleVFE, self).__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        points_mean = features[:, :, :self.num_features].sum(dim=1, keepdim=False) \
            / num_points.type_as(features).view(-1, 1)
        return points_mean.contiguous()
