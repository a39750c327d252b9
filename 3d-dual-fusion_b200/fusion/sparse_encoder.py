"""mmdet3d ``MIDDLE_ENCODERS['SparseEncoderFusion' | 'SparseEncoder']``: the sparse 3-D backbone
with the fusion hook (TransFusion/mmdet3d/models/middle_encoders/sparse_encoder.py:207-448)."""
import torch
from torch import nn

from ..ops import spconv
from ..registry import MIDDLE_ENCODERS, build_fusion_layer
from .sparse_block import SparseBasicBlock, make_sparse_convmodule


@MIDDLE_ENCODERS.register_module()
class SparseEncoderFusion(nn.Module):
    def __init__(self, in_channels, sparse_shape, order=("conv", "norm", "act"),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), base_channels=16,
                 output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                 block_type="conv_module", fusion_layer=None, fusion_pos=None, voxel_size=None,
                 point_cloud_range=None, ret_img_map=False):
        super().__init__()
        assert block_type in ["conv_module", "basicblock"]
        assert isinstance(order, tuple) and len(order) == 3
        assert set(order) == {"conv", "norm", "act"}
        self.sparse_shape = sparse_shape
        self.in_channels = in_channels
        self.order = order
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.stage_num = len(encoder_channels)
        self.fp16_enabled = False
        self.fusion_layer = None
        self.fusion_pos = None
        self.ret_img_map = ret_img_map
        if fusion_layer is not None:
            self.fusion_layer = (fusion_layer if isinstance(fusion_layer, nn.Module)
                                 else build_fusion_layer(fusion_layer))
            self.fusion_pos = fusion_pos
            self.voxel_size = voxel_size
            self.point_cloud_range = point_cloud_range
        first_order = ("conv",) if order[0] != "conv" else ("conv", "norm", "act")
        self.conv_input = make_sparse_convmodule(in_channels, base_channels, 3, norm_cfg=norm_cfg,
                                                 padding=1, indice_key="subm1", conv_type="SubMConv3d",
                                                 order=first_order)
        encoder_out_channels = self.make_encoder_layers(make_sparse_convmodule, norm_cfg, base_channels,
                                                        block_type=block_type)
        self.conv_out = make_sparse_convmodule(encoder_out_channels, output_channels, kernel_size=(3, 1, 1),
                                               stride=(2, 1, 1), norm_cfg=norm_cfg, padding=0,
                                               indice_key="spconv_down2", conv_type="SparseConv3d")

    def coor2pts(self, x, pad=0.0):
        """Metric (x, y, z) centres of the active voxels of ``x``, one tensor per sample
        (sparse_encoder.py:309-319; same fp32 expression order)."""
        ratio = self.sparse_shape[1] / x.spatial_shape[1]
        dev = x.indices.device
        consts = getattr(self, "_coor2pts_consts", None)
        if consts is None or consts[0].device != dev:
            # two configuration constants: made once per device, not by a pageable host->device copy in every forward
            consts = self._coor2pts_consts = (torch.tensor((list(self.voxel_size) + [1])[::-1], device=dev),
                                              torch.tensor(list(self.point_cloud_range[:3])[::-1], device=dev))
        scale, origin = consts
        pts = (x.indices.to(torch.float) + pad) * scale * ratio
        pts[:, 0] = pts[:, 0] / ratio - pad
        pts[:, 1:] += origin
        pts[:, 1:] = pts[:, [3, 2, 1]]
        # rows are grouped by sample already (voxelization order at stride 1, sorted flat index after
        # a strided conv), so per-sample lists are contiguous slices: one count read, no masks
        counts = torch.bincount(x.indices[:, 0].long(), minlength=x.batch_size).tolist()
        return list(torch.split(pts[:, 1:], counts))

    def forward(self, voxel_features, coors, batch_size, img_feats=None, img_metas=None, points=None,
                ret_lidar_features=False, img=None):
        coors = coors.int()
        x = spconv.SparseConvTensor(voxel_features, coors, self.sparse_shape, int(batch_size))
        x = self.conv_input(x)
        encode_features = []
        for idx, encoder_layer in enumerate(self.encoder_layers):
            x = encoder_layer(x)
            if self.fusion_pos is not None and idx in self.fusion_pos:
                c_pts = self.coor2pts(x, 0.5)
                x.features = self.fusion_layer(img_feats, c_pts, x.features, img_metas, img)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        spatial_features = out.dense()
        N, C, D, H, W = spatial_features.shape
        spatial_features = spatial_features.view(N, C * D, H, W)
        if ret_lidar_features:
            return spatial_features, encode_features[-1], img_feats
        return spatial_features

    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type="conv_module",
                            conv_cfg=dict(type="SubMConv3d")):
        assert block_type in ["conv_module", "basicblock"]
        self.encoder_layers = spconv.SparseSequential()
        n_stage = len(self.encoder_channels)
        for i, blocks in enumerate(self.encoder_channels):
            blocks = tuple(blocks)
            blocks_list = []
            for j, out_channels in enumerate(blocks):
                padding = tuple(self.encoder_paddings[i])[j]
                strided_first = i != 0 and j == 0 and block_type == "conv_module"
                strided_last = block_type == "basicblock" and j == len(blocks) - 1 and i != n_stage - 1
                if strided_first or strided_last:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, stride=2,
                                                  padding=padding, indice_key="spconv%d" % (i + 1),
                                                  conv_type="SparseConv3d"))
                elif block_type == "basicblock":
                    blocks_list.append(SparseBasicBlock(out_channels, out_channels, norm_cfg=norm_cfg,
                                                        conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  padding=padding, indice_key="subm%d" % (i + 1),
                                                  conv_type="SubMConv3d"))
                in_channels = out_channels
            self.encoder_layers.add_module("encoder_layer%d" % (i + 1), spconv.SparseSequential(*blocks_list))
        return out_channels


@MIDDLE_ENCODERS.register_module()
class SparseEncoder(SparseEncoderFusion):
    """LiDAR-only variant (sparse_encoder.py:11-205): same stack without the fusion hook."""

    def __init__(self, in_channels, sparse_shape, **kwargs):
        kwargs.pop("fusion_layer", None)
        super().__init__(in_channels, sparse_shape, **kwargs)

    def forward(self, voxel_features, coors, batch_size):
        return super().forward(voxel_features, coors, batch_size)
