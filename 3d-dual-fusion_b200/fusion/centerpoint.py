"""CenterPoint (Det3D) flavour of the hot path: registry entries ``BACKBONES['SpMiddleResNetFHD' |
'SpMiddleResNetFHDFusion']`` (CenterPoint/det3d/models/backbones/scn.py:52-236) and
``FUSION['VoxelWithPointProjection']`` (CenterPoint/det3d/models/fusion/
voxel_with_point_projection.py:14-385 with Point2ImageProjection, point_to_image_projection.py:17-231).

Differences from the TransFusion flavour that are kept: every camera builds
its OWN query list from the voxels it sees (inside the image and deeper than the per-camera
``depth_thres``), so a voxel in two fields of view is a query twice and receives two additive
updates, voxels seen by no camera are not queries; reference points are integer FEATURE-MAP pixels
divided by (W_f, H_f); voxel positions are voxel CORNERS (index * size + range_min); queries are the
``x_conv4`` voxels.  What is new: the B x 6 Python loops over boolean masks become one vectorised
pass (projection of all voxels to all cameras, one stable sort by (sample, camera)).
The IFAT image gate (``ifat_cfg``, fusion/ifat.py) gates the camera planes with the voxel features of
the configured backbone scales before the encoder samples them; the 2-D segmentation auxiliary
(``seg_cfg``) is outside the hot path and raises NotImplementedError.
"""
import numpy as np
import torch
from torch import nn

from ..ops import spconv
from ..ops.spconv import SparseConv3d, SubMConv3d
from ..registry import BACKBONES, FUSION
from .actr import build as build_actr
from ..ops import fused as _fused
from ..ops import sparse_norm
from .sparse_block import build_norm_layer


def _matvec(m, v):
    """Per-row m[n] @ v[n] (or v[n] @ m for a shared 2-D m via ``_rowmat``) as explicit multiply-adds: geometry must
    never be routed to a library GEMM, which runs tf32 under ``allow_tf32`` and moves projected pixels by up to a pixel."""
    return (m * v[:, None, :]).sum(-1)


def _rowmat(p, m):
    """p (n, k) @ m (k, j) in exact fp32 multiply-adds."""
    return (p[:, :, None] * m[None]).sum(1)



def replace_feature(out, new_features):
    return out.replace_feature(new_features)


def conv3x3(in_planes, out_planes, stride=1, indice_key=None, bias=True):
    return SubMConv3d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                      indice_key=indice_key)


class SparseBasicBlock(spconv.SparseModule):
    """scn.py:52-99: SubM convs WITH bias and a shared indice_key per stage."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_cfg=None, downsample=None, indice_key=None):
        super(SparseBasicBlock, self).__init__()
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        bias = norm_cfg is not None
        self.conv1 = conv3x3(inplanes, planes, stride, indice_key=indice_key, bias=bias)
        self.bn1 = build_norm_layer(norm_cfg, planes)[1]
        self.relu = nn.ReLU()
        self.conv2 = conv3x3(planes, planes, indice_key=indice_key, bias=bias)
        self.bn2 = build_norm_layer(norm_cfg, planes)[1]
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.conv1(x)
        out = replace_feature(out, sparse_norm.batch_norm_act(self.bn1, out.features, None, True))
        out = self.conv2(out)
        if self.downsample is not None:
            identity = self.downsample(x)
        return replace_feature(out, sparse_norm.batch_norm_act(self.bn2, out.features, identity.features, True))


@BACKBONES.register_module
class SpMiddleResNetFHD(nn.Module):
    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleResNetFHD", **kwargs):
        super(SpMiddleResNetFHD, self).__init__()
        self.name = name
        self.dcn = None
        self.zero_init_residual = False
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        bn = lambda c: build_norm_layer(norm_cfg, c)[1]
        self.conv_input = spconv.SparseSequential(
            SubMConv3d(num_input_features, 16, 3, bias=False, indice_key="res0"), bn(16), nn.ReLU(inplace=True))
        self.conv1 = spconv.SparseSequential(
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"),
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"))
        self.conv2 = spconv.SparseSequential(
            SparseConv3d(16, 32, 3, 2, padding=1, bias=False), bn(32), nn.ReLU(inplace=True),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"))
        self.conv3 = spconv.SparseSequential(
            SparseConv3d(32, 64, 3, 2, padding=1, bias=False), bn(64), nn.ReLU(inplace=True),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"))
        self.conv4 = spconv.SparseSequential(
            SparseConv3d(64, 128, 3, 2, padding=[0, 1, 1], bias=False), bn(128), nn.ReLU(inplace=True),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"))
        self.extra_conv = spconv.SparseSequential(
            SparseConv3d(128, 128, (3, 1, 1), (2, 1, 1), bias=False), bn(128), nn.ReLU())

    def _stem(self, voxel_features, coors, batch_size, input_shape):
        sparse_shape = (np.array(input_shape[::-1]) + [1, 0, 0]).tolist()
        ret = spconv.SparseConvTensor(voxel_features, coors.int(), sparse_shape, batch_size)
        x = self.conv_input(ret)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        return x_conv1, x_conv2, x_conv3, x_conv4

    def _head(self, x_conv1, x_conv2, x_conv3, x_conv4):
        ret = self.extra_conv(x_conv4).dense()
        N, C, D, H, W = ret.shape
        return ret.view(N, C * D, H, W), {"conv1": x_conv1, "conv2": x_conv2, "conv3": x_conv3, "conv4": x_conv4}

    def forward(self, voxel_features, coors, batch_size, input_shape):
        return self._head(*self._stem(voxel_features, coors, batch_size, input_shape))


@BACKBONES.register_module
class SpMiddleResNetFHDFusion(SpMiddleResNetFHD):
    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleResNetFHDFusion", **kwargs):
        super(SpMiddleResNetFHDFusion, self).__init__(num_input_features, norm_cfg, name, **kwargs)

    def forward(self, voxel_features, batch_dict, coors, batch_size, input_shape, example, fuse_func=None):
        x_conv1, x_conv2, x_conv3, x_conv4 = self._stem(voxel_features, coors, batch_size, input_shape)
        if fuse_func.fuse_mode == "pfat":
            x_conv4 = fuse_func(batch_dict, example, encoded_voxel_list=[x_conv2, x_conv3, x_conv4],
                                layer_name="layer1_ori", fuse_mode="pfat", d_factor_list=[2, 4, 8])
        return self._head(x_conv1, x_conv2, x_conv3, x_conv4)


class Point2ImageProjection(nn.Module):
    """Voxel index -> LiDAR corner -> camera -> integer image pixel, all cameras at once."""

    def __init__(self, voxel_size, pc_range, depth_thres={}, double_flip=False, device="cuda"):
        super().__init__()
        if double_flip:
            raise NotImplementedError("double-flip test-time augmentation is outside the hot path")
        self.voxel_size = [float(v) for v in voxel_size]
        self.pc_range = pc_range
        self.depth_thres = depth_thres

    def lidar_points(self, indices, d_factor, batch_dict):
        """indices (N, 4) int [b, z, y, x] -> (N, 3) xyz of the voxel corner, reverse-augmented."""
        dev = indices.device
        size = torch.tensor([v * d_factor for v in self.voxel_size], dtype=torch.float32, device=dev)
        pc_min = torch.tensor([float(v) for v in self.pc_range[:3]], dtype=torch.float32, device=dev)
        pts = indices[:, [3, 2, 1]].float() * size + pc_min
        if "aug_matrix_inv" in batch_dict:
            b_idx = indices[:, 0].long()
            out = pts.clone()
            for b, aug in enumerate(batch_dict["aug_matrix_inv"]):
                sel = b_idx == b
                p = pts[sel]
                for aug_type in ["translate", "rescale", "rotate", "flip"]:
                    if aug_type in aug:
                        m = torch.as_tensor(np.asarray(aug[aug_type]), dtype=torch.float32, device=dev)
                        p = p + m if aug_type == "translate" else _rowmat(p, m)
                out[sel] = p
            pts = out
        return pts

    def project_all(self, indices, pts, image_scale, batch_dict, cam_keys, Hf=1, Wf=1):
        """All cameras in ONE kernel (csrc/projection.cu: ddf_project_cameras): image_grid (n_cam, N, 2) long, depth,
        point_mask as ``forward``, plus the pixel of the (Hf, Wf) feature map under every (camera, voxel)."""
        from .. import lib as _lib
        dev = indices.device
        n, n_cam = indices.shape[0], len(cam_keys)
        keys = [k.lower().lstrip("cam_") for k in cam_keys]          # the reference's (quirky) key derivation
        l2c = torch.stack([batch_dict["calib"]["lidar2cam_" + k].to(dev) for k in keys]).float().contiguous()
        intr = torch.stack([batch_dict["calib"]["cam_intrinsic_" + k].to(dev) for k in keys]).float().contiguous()
        shape = torch.stack([batch_dict["image_shape"][k.lower()].to(dev) for k in cam_keys]).float().contiguous()
        thres = torch.tensor([float(self.depth_thres[k.upper()]) for k in cam_keys], dtype=torch.float32, device=dev)
        grid = torch.empty((n_cam, n, 2), dtype=torch.long, device=dev)
        depth = torch.empty((n_cam, n), dtype=torch.float32, device=dev)
        mask = torch.empty((n_cam, n), dtype=torch.bool, device=dev)
        fx = torch.empty((n_cam, n), dtype=torch.long, device=dev)
        fy = torch.empty((n_cam, n), dtype=torch.long, device=dev)
        idx = indices.contiguous()
        p = pts.contiguous().float()
        with _lib.on_device(dev):
            rc = _lib.get_lib().ddf_project_cameras(
                _lib.ptr(idx), _lib.ptr(p), _lib.ptr(l2c), _lib.ptr(intr), _lib.ptr(shape), _lib.ptr(thres), n, n_cam,
                l2c.shape[1], float(image_scale), int(Hf), int(Wf), _lib.ptr(grid), _lib.ptr(depth), _lib.ptr(mask),
                _lib.ptr(fx), _lib.ptr(fy), _lib.current_stream())
        _lib.check(rc, "project_cameras")
        return grid, depth, mask, fx, fy

    def forward(self, indices, pts, image_scale, batch_dict, cam_keys):
        """Returns image_grid (n_cam, N, 2) long (x, y) in scaled-image pixels, depth (n_cam, N) and
        point_mask (n_cam, N) following transform_grid / forward of the reference projector."""
        if indices.is_cuda and indices.dtype == torch.int32:
            return self.project_all(indices, pts, image_scale, batch_dict, cam_keys)[:3]
        return self.forward_eager(indices, pts, image_scale, batch_dict, cam_keys)

    def forward_eager(self, indices, pts, image_scale, batch_dict, cam_keys):
        """The tensor-op form (host tensors; the oracle-driven CPU path)."""
        b_idx = indices[:, 0].long()
        grids, depths, masks = [], [], []
        homo = torch.cat([pts, pts.new_ones(pts.shape[0], 1)], 1)
        for cam_key in cam_keys:
            calib_key = cam_key.lower().lstrip("cam_")  # the reference's (quirky) key derivation
            l2c = batch_dict["calib"]["lidar2cam_" + calib_key].float()[b_idx]          # (N, 4, 4)
            K = batch_dict["calib"]["cam_intrinsic_" + calib_key].float()[b_idx]        # (N, 3, 3)
            cam = _matvec(l2c, homo)
            cam = cam[:, :3] / cam[:, 3:4]                                               # transform_points
            depth = cam[:, 2].clone()
            img = _matvec(K, cam)
            img = img / img[:, 2:3]                                                      # camera_to_image
            grid = img[:, :2].long()
            grid = (image_scale * grid.float()).long()
            shape = batch_dict["image_shape"][cam_key.lower()].to(grid.device)[b_idx]    # (N, 2) H, W
            mask = ((grid[:, 0] > 0) & (grid[:, 0] < shape[:, 1]) & (grid[:, 1] > 0) & (grid[:, 1] < shape[:, 0])
                    & (depth > self.depth_thres[cam_key.upper()]))
            grids.append(torch.where(mask[:, None], grid, torch.zeros_like(grid)))
            depths.append(torch.where(mask, depth, torch.zeros_like(depth)))
            masks.append(mask)
        return torch.stack(grids), torch.stack(depths), torch.stack(masks)


@FUSION.register_module
class VoxelWithPointProjection(nn.Module):
    def __init__(self, fuse_mode, interpolate, voxel_size, pc_range, image_list, image_scale=1, depth_thres=0,
                 double_flip=False, layer_channel=None, pfat_cfg=None, lt_cfg=None, ifat_cfg=None, seg_cfg=None,
                 model_name="ACTR"):
        super().__init__()
        if seg_cfg:
            raise NotImplementedError("the 2-D segmentation auxiliary head (seg_cfg) is outside the hot path")
        if interpolate:
            raise NotImplementedError("interpolate=True (full-resolution image features) is not configured by 3D-DF")
        self.voxel_size = voxel_size
        self.pc_range = pc_range
        self.point_projector = Point2ImageProjection(voxel_size=voxel_size, pc_range=pc_range,
                                                     depth_thres=depth_thres, double_flip=double_flip)
        self.fuse_mode = fuse_mode
        self.image_interp = interpolate
        self.image_list = image_list
        self.image_scale = image_scale
        self.double_flip = double_flip
        if self.fuse_mode == "pfat":
            self.pfat = build_actr(pfat_cfg, lt_cfg=lt_cfg, model_name=model_name)
        elif self.fuse_mode not in ("sum", "mean"):
            raise NotImplementedError("fuse_mode %r" % (fuse_mode,))
        self.ifat_cfg = None
        if ifat_cfg:
            from . import ifat
            self.ifat_cfg = ifat_cfg
            cfg = dict(ifat_cfg)
            self.ifat = ifat.__all__[cfg.pop("fusion_method")](**cfg)
        self.seg_cfg = None

    def _feature_pixels(self, indices, pts, batch_dict, cams, Hf, Wf):
        """Feature-map pixel (x, y) under every (camera, voxel) and the visibility mask, each (n_cam, N):
        projection + scaled-image pixels -> feature-map pixels (voxel_with_point_projection.py:253-257)."""
        if indices.is_cuda and indices.dtype == torch.int32:
            _, _, mask, gx, gy = self.point_projector.project_all(indices, pts, self.image_scale, batch_dict,
                                                                  self.image_list, Hf, Wf)
            return gx, gy, mask
        grid, _, mask = self.point_projector(indices, pts, self.image_scale, batch_dict, self.image_list)
        b_idx = indices[:, 0].long()
        raw = torch.stack([batch_dict["image_shape"][c].to(grid.device)[b_idx] for c in cams]).float()  # (n_cam,N,2)
        gf = grid.float()
        return (gf[..., 0] * (Wf / raw[..., 1])).long(), (gf[..., 1] * (Hf / raw[..., 0])).long(), mask

    def _queries(self, sp_tensor, d_factor, batch_dict, cams, Hf, Wf):
        """(camera, voxel) queries of one backbone scale in (sample, camera)-major, voxel-minor order:
        group id, feature-map pixel (x, y), voxel row, reverse-augmented xyz."""
        indices = sp_tensor.indices
        n_cam = len(cams)
        pts = self.point_projector.lidar_points(indices, d_factor, batch_dict)
        b_idx = indices[:, 0].long()
        gx, gy, mask = self._feature_pixels(indices, pts, batch_dict, cams, Hf, Wf)
        cam_id, vox = mask.nonzero(as_tuple=True)
        group = b_idx[vox] * n_cam + cam_id
        order = torch.sort(group, stable=True)[1]
        cam_id, vox, group = cam_id[order], vox[order], group[order]
        return group, gx[cam_id, vox], gy[cam_id, vox], vox, pts

    def _gate_image_features(self, img, encoded_voxel_list, d_factor_list, batch_dict, cams, batch_size):
        """IFAT (voxel_with_point_projection.py:277-293): gate every camera plane with the voxel features
        of the configured backbone scales before the 3D-DF encoder samples it."""
        Hf, Wf = img.shape[-2:]
        feats, cells, coords = {}, {}, {}
        for s in self.ifat.voxel_idx:
            group, qx, qy, vox, pts = self._queries(encoded_voxel_list[s], d_factor_list[s], batch_dict, cams, Hf, Wf)
            feats[s] = encoded_voxel_list[s].features[vox]
            cells[s] = (group * Hf + qy) * Wf + qx
            coords[s] = pts[vox]
        return self.ifat(img, feats, cells, coords)

    def forward(self, batch_dict, example, encoded_voxel_list=None, layer_name=None, img_conv_func=None,
                fuse_mode=None, d_factor_list=None):
        encoded_voxel = encoded_voxel_list[-1]
        fuse_mode = fuse_mode or self.fuse_mode
        cams = [c.lower() for c in self.image_list]
        n_cam = len(cams)
        batch_size = len(batch_dict["image_shape"][cams[0]])
        indices = encoded_voxel.indices
        feats = encoded_voxel.features
        n = indices.shape[0]
        # image features: (B * n_cam, C, Hf, Wf), sample-major like the reference's reshape (:300-304)
        img = torch.stack([batch_dict["img_feat"][layer_name + "_feat2d"][c] for c in cams], 1)
        if img_conv_func:
            img = img_conv_func(img.flatten(0, 1)).unflatten(0, (batch_size, n_cam))
        img = img.flatten(0, 1)
        Hf, Wf = img.shape[-2:]

        pts = self.point_projector.lidar_points(indices, d_factor_list[-1], batch_dict)
        b_idx = indices[:, 0].long()
        gx, gy, mask = self._feature_pixels(indices, pts, batch_dict, cams, Hf, Wf)

        if self.ifat_cfg is not None and fuse_mode == "pfat":
            img = self._gate_image_features(img, encoded_voxel_list, d_factor_list, batch_dict, cams, batch_size)

        cam_id, vox = mask.nonzero(as_tuple=True)            # every (camera, voxel) query
        group = b_idx[vox] * n_cam + cam_id
        order = torch.sort(group, stable=True)[1]            # keeps voxel order inside a (sample, camera)
        cam_id, vox, group = cam_id[order], vox[order], group[order]
        qx, qy = gx[cam_id, vox], gy[cam_id, vox]
        if img.is_cuda and img.dtype == torch.float32 and fuse_mode == "pfat":
            # camera features (frozen, or the output of the IFAT gate: the conversion carries the gradient): token-major
            # once - the feature under a query is one contiguous row, and ACTR's input projection runs on the same rows
            img = _fused.nchw_to_rows(img.contiguous())
            v_i = img.rows.view(-1, img.rows.shape[-1]).index_select(0, (group * Hf + qy.remainder(Hf)) * Wf + qx.remainder(Wf))
        else:
            v_i = img[group, :, qy, qx]                       # camera feature under each query

        if fuse_mode in ("sum", "mean"):
            upd = torch.zeros_like(feats).index_add_(0, vox, v_i)
            new = feats + upd if fuse_mode == "sum" else None
            if new is None:  # 'mean' is applied camera after camera in the reference; keep that order
                new = feats
                for c in range(n_cam):
                    sel = cam_id == c
                    cur = new.clone()
                    cur[vox[sel]] = (new[vox[sel]] + v_i[sel]) / 2
                    new = cur
            return encoded_voxel.replace_feature(new)

        n_groups = batch_size * n_cam
        counts = torch.bincount(group, minlength=n_groups)
        starts = torch.cumsum(counts, 0) - counts
        col = torch.arange(group.numel(), device=group.device) - starts[group]
        max_ne = int(counts.max().item()) if group.numel() else 0

        def pad(x):
            out = x.new_zeros((n_groups, max_ne) + tuple(x.shape[1:]))
            out[group, col] = x
            return out

        ref = torch.stack([qx, qy], -1).float() / torch.tensor([Wf, Hf], dtype=torch.float32, device=feats.device)
        enh = self.pfat(v_feat=pad(feats[vox]), grid=pad(ref), i_feats=[img], lidar_grid=pad(pts[vox]),
                        v_i_feat=pad(v_i),
                        valid_index=(group * max_ne + col) if group.numel() < 0.75 * n_groups * max_ne else None)
        new = feats.index_add(0, vox, enh[group, col])       # one additive update per (voxel, camera)
        return encoded_voxel.replace_feature(new)
