"""Voxel-RCNN (OpenPCDet) flavour of the hot path: ``BACKBONE_3D.NAME: VoxelBackBone8xFusion``
(VoxelRCNN/pcdet/models/backbones_3d/spconv_backbone.py:436-929; ``post_act_block`` :25-77).

Single camera, EVERY stride-8 voxel is a query (out-of-image ones sample zeros and carry a zero image
feature, :724-742), queries are padded to the largest sample of the batch, the fusion encoder runs
with the Voxel-RCNN layer ordering (gate before the FFNs) and, for ``*ACTRv2*``, with the 3D local
self-attention in front of every layer.  ``FUSION_POS`` 1 is the cheap "MVX" gather-and-add of a
reduced image feature at stride 1 (:860-875).

What is new:
  * the per-voxel image feature is read by evaluating the bilinear interpolation formula at the
    voxel's pixel only; the reference up-samples the whole 256-channel map to image resolution
    (``F.interpolate(x_rgb[0], (h, w))``, 477 MB per KITTI image, :680-681) to pick one pixel per
    voxel;
  * projection runs on the GPU when ``batch_dict['lidar2img']`` (B, 3|4, 4) is given; KITTI ``calib``
    objects with ``lidar_to_img`` are still honoured (host round trip, as in the reference :717-719).
The image branch (DeepLabV3 ``semseg``), IFAT gate and auxiliary losses are outside the hot path: the
backbone reads image features from ``batch_dict['img_dict']`` (or from a ``semseg`` callable the host
attaches).
"""
from functools import partial

import numpy as np
import torch
from torch import nn

from ..ops import spconv
from .actr import build as build_actr


def _matvec(m, v):
    """Per-row m[n] @ v[n] (or v[n] @ m for a shared 2-D m via ``_rowmat``) as explicit multiply-adds: geometry must
    never be routed to a library GEMM, which runs tf32 under ``allow_tf32`` and moves projected pixels by up to a pixel."""
    return (m * v[:, None, :]).sum(-1)


def _rowmat(p, m):
    """p (n, k) @ m (k, j) in exact fp32 multiply-adds."""
    return (p[:, :, None] * m[None]).sum(1)



def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0,
                   conv_type="subm", norm_fn=None):
    if conv_type == "subm":
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == "spconv":
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                   bias=False, indice_key=indice_key)
    elif conv_type == "inverseconv":
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
    else:
        raise NotImplementedError
    return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


def _get(cfg, key, default=None):
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def sample_upsampled_pixels(feat, u, v, h, w):
    """Value of ``F.interpolate(feat[None], (h, w), mode='bilinear')[0][:, v, u]`` (align_corners
    False) without materialising the up-sampled map. feat (C, Hf, Wf); u, v long pixel indices."""
    C, Hf, Wf = feat.shape
    sy = ((v.float() + 0.5) * (Hf / h) - 0.5).clamp(min=0)
    sx = ((u.float() + 0.5) * (Wf / w) - 0.5).clamp(min=0)
    y0 = sy.floor().long().clamp(max=Hf - 1)
    x0 = sx.floor().long().clamp(max=Wf - 1)
    y1 = (y0 + 1).clamp(max=Hf - 1)
    x1 = (x0 + 1).clamp(max=Wf - 1)
    ly = sy - y0.float()
    lx = sx - x0.float()
    f = feat
    top = f[:, y0, x0] * (1 - lx) + f[:, y0, x1] * lx
    bot = f[:, y1, x0] * (1 - lx) + f[:, y1, x1] * lx
    return (top * (1 - ly) + bot * ly).permute(1, 0)


class VoxelBackBone8xFusion(nn.Module):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = (np.asarray(grid_size)[::-1] + [1, 0, 0]).tolist()
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1"), norm_fn(16), nn.ReLU())
        block = post_act_block
        self.fusion_pos = _get(model_cfg, "FUSION_POS", [1])
        self.fusion_method = _get(model_cfg, "FUSION_METHOD", "MVX")
        self.feature_levels = _get(model_cfg, "FEATURE_LEVELS", [0])
        # (z, y, x) voxel size and range minimum, hard-coded in the reference (:468-470)
        self.register_buffer("voxel_size", torch.tensor([0.1, 0.05, 0.05]), persistent=False)
        self.register_buffer("point_cloud_range", torch.tensor([-3.0, -40.0, 0.0, 1.0, 40.0, 70.4]), persistent=False)
        self.inv_idx = [2, 1, 0]
        self.img_out_channel = 16 if 1 in self.fusion_pos else 64
        self.semseg = None  # image branch: attached by the host, or features come in batch_dict['img_dict']
        if "ACTR" in self.fusion_method:
            model_name = self.fusion_method if "MVX+" not in self.fusion_method else self.fusion_method[4:]
            actr_cfg = _get(model_cfg, "ACTR_CFG", None)
            assert actr_cfg is not None
            self.actr = build_actr(actr_cfg, model_name=model_name, lt_cfg=_get(model_cfg, "LT_CFG", None),
                                   hybrid_cfg=_get(model_cfg, "HYBRID_CFG", None), gate_first=True)
            self.max_num_nev = _get(actr_cfg, "max_num_ne_voxel", 26000)
        if _get(model_cfg, "I_FUSION_METHOD", False):
            raise NotImplementedError("IFAT image gate is a 'next' row (SURVEY.md 8f-2); set I_FUSION_METHOD: False")
        self.conv1 = spconv.SparseSequential(block(16, 16, 3, norm_fn=norm_fn, padding=1, indice_key="subm1"))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key="spconv2", conv_type="spconv"),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key="subm2"),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key="subm2"))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key="spconv3", conv_type="spconv"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm3"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm3"))
        self.conv4 = spconv.SparseSequential(
            block(64, 64, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key="spconv4", conv_type="spconv"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm4"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm4"))
        last_pad = _get(model_cfg, "last_pad", 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key="spconv_down2"), norm_fn(128), nn.ReLU())
        self.num_point_features = 128
        self.backbone_channels = {"x_conv1": 16, "x_conv2": 32, "x_conv3": 64, "x_conv4": 64}

    # ---- projection -------------------------------------------------------------------------------
    def _voxel_xyz(self, x, voxel_stride, batch_dict):
        """(N, 3) LiDAR (x, y, z) of the voxel corners with the training augmentations undone."""
        idx = x.indices
        zyx = (idx[:, 1:] * voxel_stride).float() * self.voxel_size + self.point_cloud_range[:3]
        b_idx = idx[:, 0].long()
        if "noise_scale" in batch_dict:
            zyx = zyx / torch.as_tensor(batch_dict["noise_scale"], dtype=zyx.dtype, device=zyx.device)[b_idx, None]
        xyz = zyx[:, self.inv_idx]
        if "noise_rot" in batch_dict:
            ang = -torch.as_tensor(batch_dict["noise_rot"], dtype=xyz.dtype, device=xyz.device)[b_idx]
            c, s = torch.cos(ang), torch.sin(ang)
            xyz = torch.stack([xyz[:, 0] * c - xyz[:, 1] * s, xyz[:, 0] * s + xyz[:, 1] * c, xyz[:, 2]], 1)
        if "flip_x" in batch_dict:
            f = torch.as_tensor(batch_dict["flip_x"], device=xyz.device)[b_idx]
            xyz = torch.stack([xyz[:, 0], torch.where(f.bool(), -xyz[:, 1], xyz[:, 1]), xyz[:, 2]], 1)
        if "flip_y" in batch_dict:
            f = torch.as_tensor(batch_dict["flip_y"], device=xyz.device)[b_idx]
            xyz = torch.stack([torch.where(f.bool(), -xyz[:, 0], xyz[:, 0]), xyz[:, 1], xyz[:, 2]], 1)
        return xyz, b_idx

    def _project(self, xyz, b_idx, batch_dict):
        """(N, 2) float pixel coordinates (u, v) in the input image."""
        if "lidar2img" in batch_dict:
            P = torch.as_tensor(batch_dict["lidar2img"], dtype=xyz.dtype, device=xyz.device)[:, :3, :]
            homo = torch.cat([xyz, xyz.new_ones(xyz.shape[0], 1)], 1)
            cam = _matvec(P[b_idx], homo)
            return cam[:, :2] / cam[:, 2:3]
        uv = xyz.new_zeros((xyz.shape[0], 2))
        for b, calib in enumerate(batch_dict["calib"]):   # reference path: KITTI calib object on the host
            sel = b_idx == b
            pts_img, _ = calib.lidar_to_img(xyz[sel].detach().cpu().numpy())
            uv[sel] = torch.as_tensor(pts_img, dtype=xyz.dtype, device=xyz.device)
        return uv

    def point_fusion(self, x_list, batch_dict, img_dict, fusion_method, voxel_stride=1):
        x = x_list[-1]
        x_rgb = [img_dict[k] for k in img_dict]
        batch_size = batch_dict["batch_size"]
        h, w = batch_dict["images"].shape[2:] if "images" in batch_dict else batch_dict["image_hw"]
        xyz, b_idx = self._voxel_xyz(x, voxel_stride, batch_dict)
        uv = self._project(xyz, b_idx, batch_dict)
        uv_int = uv.long()
        inside = (uv_int[:, 1] >= 0) & (uv_int[:, 1] < h) & (uv_int[:, 0] >= 0) & (uv_int[:, 0] < w)
        img_feat = x.features.new_zeros((x.features.shape[0], x_rgb[0].shape[1]))
        for b in range(batch_size):
            sel = (b_idx == b) & inside
            img_feat[sel] = sample_upsampled_pixels(x_rgb[0][b], uv_int[sel, 0], uv_int[sel, 1], h, w)
        if "ACTR" not in fusion_method:                      # MVX: gather-and-add
            return x.replace_feature(img_feat + x.features)
        counts = torch.bincount(b_idx, minlength=batch_size)
        starts = torch.cumsum(counts, 0) - counts
        col = torch.arange(b_idx.numel(), device=b_idx.device) - starts[b_idx]
        n_max = int(counts.max().item())

        def pad(t):
            out = t.new_zeros((batch_size, n_max) + tuple(t.shape[1:]))
            out[b_idx, col] = t
            return out

        grid = uv / uv.new_tensor([w, h])                     # not masked: out-of-image queries sample zeros
        enh = self.actr(v_feat=pad(x.features), v_i_feat=pad(img_feat), grid=pad(grid), i_feats=x_rgb,
                        lidar_grid=pad(xyz))
        return x.replace_feature(enh[b_idx, col] + x.features)

    def forward(self, batch_dict):
        voxel_features, voxel_coords = batch_dict["voxel_features"], batch_dict["voxel_coords"]
        batch_size = batch_dict["batch_size"]
        input_sp_tensor = spconv.SparseConvTensor(features=voxel_features, indices=voxel_coords.int(),
                                                  spatial_shape=self.sparse_shape, batch_size=batch_size)
        img_dict = dict(self.semseg(batch_dict["images"]) if self.semseg is not None else batch_dict["img_dict"])
        x = self.conv_input(input_sp_tensor)
        x_conv1 = self.conv1(x)
        if 1 in self.fusion_pos:
            if "mvx_layer1_feat2d" in img_dict:
                x_conv1 = self.point_fusion([x_conv1], batch_dict, {"layer2_feat2d": img_dict.pop("mvx_layer1_feat2d")},
                                            "MVX", voxel_stride=1)
            else:
                x_conv1 = self.point_fusion([x_conv1], batch_dict, img_dict, "MVX", voxel_stride=1)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        if 4 in self.fusion_pos:
            if 0 not in self.feature_levels and "layer1_feat2d" in img_dict:
                img_dict.pop("layer1_feat2d")
            x_conv4 = self.point_fusion([x_conv2, x_conv3, x_conv4], batch_dict, img_dict, "ACTR", voxel_stride=8)
        out = self.conv_out(x_conv4)
        batch_dict.update({"encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8})
        batch_dict.update({"multi_scale_3d_features": {"x_conv1": x_conv1, "x_conv2": x_conv2, "x_conv3": x_conv3,
                                                        "x_conv4": x_conv4}})
        batch_dict.update({"multi_scale_3d_strides": {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8}})
        return batch_dict
