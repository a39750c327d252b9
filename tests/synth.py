"""Seeded synthetic inputs shared by tests and bench."""
import numpy as np

NUSC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
NUSC_VOXEL = [0.075, 0.075, 0.2]
KITTI_RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
KITTI_VOXEL = [0.05, 0.05, 0.1]


def lidar_points(n, seed=0, nfeat=5, rng_m=54.0, beams=32, forward_only=False):
    """Ring-pattern LiDAR sweep: `beams` elevation rings, uniform azimuth, log-normal range,
    Gaussian jitter; features [x, y, z, intensity, dt]; shuffled like the reference's PointShuffle."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-30.7, 10.7, beams))[rng.integers(0, beams, n)]
    az = rng.uniform(-np.pi / 2 if forward_only else -np.pi, np.pi / 2 if forward_only else np.pi, n)
    r = np.clip(rng.lognormal(2.7, 0.7, n), 1.0, rng_m)
    x = r * np.cos(elev) * np.cos(az)
    y = r * np.cos(elev) * np.sin(az)
    z = r * np.sin(elev) - 1.8 + 1.8  # sensor ~1.8 m above ground; keep lidar frame
    pts = np.stack([x, y, z], 1) + rng.normal(0, 0.02, (n, 3))
    feats = [pts, rng.random((n, 1))]
    if nfeat >= 5:
        feats.append(rng.integers(0, 10, (n, 1)) * 0.05)
    out = np.concatenate(feats, 1).astype(np.float32)
    rng.shuffle(out)
    return np.ascontiguousarray(out[:, :nfeat])


def uniform_points(n, pc_range, seed=0, nfeat=5, margin=1.0):
    """Uniform points in (and slightly outside) the range: worst case for hashing, exercises the
    out-of-range rejection."""
    rng = np.random.default_rng(seed)
    lo = np.array(pc_range[:3]) - margin
    hi = np.array(pc_range[3:]) + margin
    xyz = rng.uniform(lo, hi, (n, 3))
    rest = rng.random((n, nfeat - 3))
    return np.ascontiguousarray(np.concatenate([xyz, rest], 1).astype(np.float32))


def camera_rig(n_cam=6, ori_hw=(900, 1600), focal=1266.0, cam_z=-0.3):
    """Pinhole rig: yaw {0, +-55, 180, +-110} deg, principal point at the centre; returns lidar2img
    (n_cam, 4, 4) float64 mapping LiDAR xyz1 -> (u*d, v*d, d, 1) in ORIGINAL image pixels."""
    yaws = np.deg2rad([0.0, 55.0, -55.0, 180.0, 110.0, -110.0])[:n_cam]
    H, W = ori_hw
    K = np.array([[focal, 0, W / 2.0, 0], [0, focal, H / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    out = []
    for th in yaws:
        fwd = np.array([np.cos(th), np.sin(th), 0.0])
        right = np.array([np.sin(th), -np.cos(th), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ np.array([0.0, 0.0, cam_z])
        out.append(K @ T)
    return np.stack(out)


def nusc_img_meta(n_cam=6, ori_hw=(900, 1600), input_hw=(448, 800)):
    """img_metas entry the TransFusion fusion layer reads (SURVEY.md section 8(b) metadata contract)."""
    scale = min(input_hw[1] / ori_hw[1], input_hw[0] / ori_hw[0])
    img_hw = (int(ori_hw[0] * scale + 0.5), int(ori_hw[1] * scale + 0.5))
    return dict(lidar2img=camera_rig(n_cam, ori_hw), ori_shape=(ori_hw[0], ori_hw[1], 3),
                img_shape=(img_hw[0], img_hw[1], 3), input_shape=input_hw,
                scale_factor=np.array([scale, scale, scale, scale], np.float32), flip=False,
                filename=["CAM_%d" % i for i in range(n_cam)], sample_idx="synthetic")


def camera_features(batch, n_cam, hw, channels=256, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((batch * n_cam, channels, hw[0], hw[1]), dtype=np.float32)


CP_CAMS = ["CAM_FRONT", "CAM_FRONT_LEFT", "CAM_FRONT_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_BACK_RIGHT"]


def centerpoint_batch(batch, feat_hw=(150, 267), img_hw=(600, 1066), ori_hw=(900, 1600), channels=256, seed=0):
    """Det3D batch_dict pieces VoxelWithPointProjection reads (SURVEY.md 8(b) metadata contract)."""
    import torch
    rig = camera_rig(6, ori_hw)             # K @ [R|t]; split back into extrinsic / intrinsic
    focal = 1266.0
    K = np.array([[focal, 0, ori_hw[1] / 2.0], [0, focal, ori_hw[0] / 2.0], [0, 0, 1]], np.float64)
    rng = np.random.default_rng(seed)
    calib, image_shape, img_feat = {}, {}, {}
    for i, cam in enumerate(CP_CAMS):
        key = cam.lower()[4:]
        Kinv = np.linalg.inv(K)
        ext = np.eye(4)
        ext[:3, :] = Kinv @ rig[i][:3, :]
        calib["lidar2cam_" + key] = torch.from_numpy(np.repeat(ext[None], batch, 0).astype(np.float32))
        calib["cam_intrinsic_" + key] = torch.from_numpy(np.repeat(K[None], batch, 0).astype(np.float32))
        image_shape[cam.lower()] = torch.tensor([list(img_hw)] * batch)
        img_feat[cam.lower()] = torch.from_numpy(rng.standard_normal((batch, channels, *feat_hw), dtype=np.float32))
    return dict(calib=calib, image_shape=image_shape, img_feat={"layer1_ori_feat2d": img_feat})


def kitti_lidar2img(img_hw=(375, 1242), focal=721.5):
    """Single forward camera (x forward) for the KITTI-shaped config: (3, 4) LiDAR -> pixel matrix."""
    H, W = img_hw
    K = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]], np.float64)
    R = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])  # cam x = -y_l, cam y = -z_l, cam z = x_l
    t = -R @ np.array([0.27, 0.0, -0.08])
    return (K @ np.concatenate([R, t[:, None]], 1)).astype(np.float32)
