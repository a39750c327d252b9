"""Input recipes shared by the golden generators (reference side, build container) and the parity tests
(our side, any box). Everything is a pure function of names/shapes through detfill.uniform."""
import numpy as np
import torch

import detfill


def actr_inputs(name, cfg, dims, valid):
    """Padded fusion-encoder inputs the way the wrappers hand them over: rows past ``valid[b]`` are zero in
    every tensor (point_fusion.py:342-382). Returns v_feat, grid, i_feat, v_i_feat, lidar_grid."""
    B, L, H, W = dims
    C, Ci = cfg["query_num_feat"], cfg["num_channels"][0]
    v_feat = detfill.uniform(name + "/v_feat", (B, L, C))
    grid = detfill.uniform(name + "/grid", (B, L, 2), -0.05, 1.05)
    i_feat = detfill.uniform(name + "/i_feat", (B, Ci, H, W))
    v_i_feat = detfill.uniform(name + "/v_i_feat", (B, L, Ci))
    lidar = np.concatenate([detfill.uniform(name + "/depth", (B, L, 1), 0.5, 58.0),
                            detfill.uniform(name + "/yz", (B, L, 2), -3.0, 3.0)], -1)
    # voxel-centre-like coordinates: snap to a 0.6 m lattice so that ball queries see exact ties / duplicates
    lidar = (np.round(lidar / 0.6) * 0.6).astype(np.float32)
    for b, n in enumerate(valid):
        for t in (v_feat, grid, v_i_feat, lidar):
            t[b, n:] = 0
    return [torch.from_numpy(np.ascontiguousarray(t)) for t in (v_feat, grid, i_feat, v_i_feat, lidar)]


# ---- TransFusion wrapper: a synthetic nuScenes "database" -----------------------------------------------
CAMS = ["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_BACK_RIGHT"]
CAM_YAW_DEG = [0.0, -55.0, 55.0, 180.0, 110.0, -110.0]


def quat_from_matrix(R):
    """(w, x, y, z) of a rotation matrix (trace branch only; the rigs below keep w well away from 0... or not:
    fall back to the largest-diagonal branches)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = np.zeros(4)
    q[0] = (R[k, j] - R[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (R[j, i] + R[i, j]) / s
    q[1 + k] = (R[k, i] + R[i, k]) / s
    return q


def matrix_from_quat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _rot_z(deg):
    a = np.deg2rad(deg)
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])


def _small_rot(name):
    r = detfill.uniform(name, (3,), -0.03, 0.03).astype(np.float64)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    q, _ = np.linalg.qr(np.eye(3) + K)
    return q * np.sign(np.linalg.det(q))


def nusc_database(n_samples, ori_hw, focal, tag="db"):
    """Records in the shape the reference's projection() reads (point_fusion.py:576-611): sample ->
    sample_data -> calibrated_sensor / ego_pose, rotations as (w,x,y,z) quaternions. Returns (tables, tokens,
    lidar2img[n_samples, 6, 4, 4] float64 = the same chain composed into one matrix per camera)."""
    H, W = ori_hw
    tables = {"sample": {}, "sample_data": {}, "calibrated_sensor": {}, "ego_pose": {}}
    tokens, l2i = [], np.zeros((n_samples, len(CAMS), 4, 4))
    cam_axes = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])   # cam(x right,y down,z fwd) <- ego
    for s in range(n_samples):
        tok = "%s_sample%d" % (tag, s)
        tokens.append(tok)
        data = {}

        def rigid(name, R, t):
            tables["calibrated_sensor" if "cs" in name else "ego_pose"][name] = dict(
                rotation=list(quat_from_matrix(R)), translation=list(t))
            return name
        # lidar sensor -> ego, ego -> global at the sweep time
        R1 = _small_rot(tok + "/lidar_cs") @ _rot_z(-90.0 + 90.0)
        t1 = np.array([0.94, 0.0, 1.84]) + detfill.uniform(tok + "/lidar_t", (3,), -0.05, 0.05)
        R2 = _rot_z(float(detfill.uniform(tok + "/ego_yaw", (1,), -180, 180)[0])) @ _small_rot(tok + "/ego_r")
        t2 = detfill.uniform(tok + "/ego_t", (3,), -200, 200).astype(np.float64)
        cs_l = rigid(tok + "/cs_lidar", R1, t1)
        ep_l = rigid(tok + "/ep_lidar", R2, t2)
        tables["sample_data"][tok + "/sd_lidar"] = dict(calibrated_sensor_token=cs_l, ego_pose_token=ep_l)
        data["LIDAR_TOP"] = tok + "/sd_lidar"
        for c, (cam, yaw) in enumerate(zip(CAMS, CAM_YAW_DEG)):
            # ego pose at the image time: a slightly moved vehicle
            R3 = R2 @ _small_rot(tok + cam + "/dr")
            t3 = t2 + R2 @ detfill.uniform(tok + cam + "/dt", (3,), -0.3, 0.3).astype(np.float64)
            R4 = _rot_z(yaw) @ cam_axes.T @ _small_rot(tok + cam + "/cs")       # camera -> ego
            t4 = _rot_z(yaw) @ np.array([1.5, 0.0, 1.5])
            K = np.array([[focal, 0, W / 2.0 + 3.0], [0, focal * 1.01, H / 2.0 - 2.0], [0, 0, 1.0]])
            cs_c = rigid(tok + cam + "/cs_cam", R4, t4)
            tables["calibrated_sensor"][cs_c]["camera_intrinsic"] = K.tolist()
            ep_c = rigid(tok + cam + "/ep_cam", R3, t3)
            tables["sample_data"][tok + cam + "/sd"] = dict(calibrated_sensor_token=cs_c, ego_pose_token=ep_c)
            data[cam] = tok + cam + "/sd"
            # rebuild from the STORED quaternions, exactly what projection() will see
            q = lambda tbl, n: matrix_from_quat(np.array(tables[tbl][n]["rotation"]))
            A1, A2, A3, A4 = q("calibrated_sensor", cs_l), q("ego_pose", ep_l), q("ego_pose", ep_c), q("calibrated_sensor", cs_c)
            Rm = A4.T @ A3.T @ A2 @ A1
            tm = A4.T @ (A3.T @ (A2 @ t1 + t2 - t3) - t4)
            M = np.eye(4)
            M[:3, :3], M[:3, 3] = K @ Rm, K @ tm
            l2i[s, c] = M
        tables["sample"][tok] = dict(data=data)
    return tables, tokens, l2i


def tf_wrapper_case(name, n_pts=(260, 340), ori_hw=(180, 320), scale=0.5, pad_hw=(96, 160), c_img=32, c_pts=64,
                    flip=(False, True), crop=((2.0, 1.0), None)):
    """Inputs of the TransFusion ``FUSION_LAYERS['ACTR']`` wrapper for B = len(n_pts) samples."""
    B = len(n_pts)
    tables, tokens, l2i = nusc_database(B, ori_hw, focal=0.8 * ori_hw[1], tag=name)
    pts, metas = [], []
    for b, n in enumerate(n_pts):
        ang = detfill.uniform("%s/ang%d" % (name, b), (n,), -np.pi, np.pi)
        rad = detfill.uniform("%s/rad%d" % (name, b), (n,), 2.0, 50.0)
        z = detfill.uniform("%s/z%d" % (name, b), (n,), -3.0, 2.0)
        pts.append(torch.from_numpy(np.stack([rad * np.cos(ang), rad * np.sin(ang), z], 1).astype(np.float32)))
        img_hw = (int(ori_hw[0] * scale), int(ori_hw[1] * scale))
        meta = dict(sample_idx=tokens[b], filename=["samples/%s/x__%s__%d.jpg" % (c, c, b) for c in CAMS],
                    ori_shape=(ori_hw[0], ori_hw[1], 3), img_shape=(img_hw[0], img_hw[1], 3),
                    input_shape=pad_hw, scale_factor=np.array([scale, scale, scale, scale], np.float32),
                    flip=flip[b], lidar2img=l2i[b].astype(np.float32))
        if crop[b] is not None:
            meta["img_crop_offset"] = list(crop[b])
        metas.append(meta)
    feats = torch.from_numpy(detfill.uniform(name + "/pts_feats", (sum(n_pts), c_pts)))
    img = torch.from_numpy(detfill.uniform(name + "/img", (B * len(CAMS), c_img, pad_hw[0] // 4, pad_hw[1] // 4)))
    return dict(pts=pts, pts_feats=feats, img_feats=[img], img_metas=metas, tables=tables)


# ---- CenterPoint wrapper: Det3D batch pieces ---------------------------------------------------------------
CP_CAMS = ["CAM_FRONT", "CAM_FRONT_LEFT", "CAM_FRONT_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_BACK_RIGHT"]
CP_YAW_DEG = [0.0, 55.0, -55.0, 180.0, 110.0, -110.0]
CP_VOXEL = [0.075, 0.075, 0.2]
CP_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]


def cp_wrapper_case(name, batch=2, n_vox=(220, 160, 120), chans=(8, 16, 64), c_img=32, img_hw=(60, 107), feat_hw=(15, 27),
                    ori_hw=(90, 160)):
    """Inputs of ``FUSION['VoxelWithPointProjection'].forward`` (fuse_mode 'pfat'): three sparse tensors (x_conv2..4,
    d_factor 2/4/8) as (indices, features) pairs, per-camera calibration / image shapes / feature maps."""
    focal = 0.8 * ori_hw[1]
    K = np.array([[focal, 0, ori_hw[1] / 2.0], [0, focal, ori_hw[0] / 2.0], [0, 0, 1]], np.float64)
    calib, image_shape, img_feat = {}, {}, {}
    for cam, yaw in zip(CP_CAMS, CP_YAW_DEG):
        key = cam.lower()[4:]
        th = np.deg2rad(yaw)
        R = np.stack([[np.sin(th), -np.cos(th), 0.0], [0.0, 0.0, -1.0], [np.cos(th), np.sin(th), 0.0]])
        ext = np.eye(4)
        ext[:3, :3] = R @ _small_rot(name + cam)
        ext[:3, 3] = -ext[:3, :3] @ np.array([0.3 * np.cos(th), 0.3 * np.sin(th), -0.3])
        calib["lidar2cam_" + key] = torch.from_numpy(np.repeat(ext[None], batch, 0).astype(np.float32))
        calib["cam_intrinsic_" + key] = torch.from_numpy(np.repeat(K[None], batch, 0).astype(np.float32))
        image_shape[cam.lower()] = torch.tensor([list(img_hw)] * batch)
        img_feat[cam.lower()] = torch.from_numpy(detfill.uniform(name + "/img/" + cam, (batch, c_img, *feat_hw)))
    tensors = []
    for s, (n, c, d) in enumerate(zip(n_vox, chans, (2, 4, 8))):
        dims = (40 // d + 1, 1440 // d, 1440 // d)
        idx = []
        for b in range(batch):
            nb = n - 17 * b
            z = (detfill.uniform("%s/z%d%d" % (name, s, b), (nb,), 0, 1) * dims[0] * 0.6).astype(np.int64)
            ang = detfill.uniform("%s/a%d%d" % (name, s, b), (nb,), -np.pi, np.pi)
            rad = detfill.uniform("%s/r%d%d" % (name, s, b), (nb,), 2.0, 50.0)
            x = ((rad * np.cos(ang) - CP_RANGE[0]) / (CP_VOXEL[0] * d)).astype(np.int64)
            y = ((rad * np.sin(ang) - CP_RANGE[1]) / (CP_VOXEL[1] * d)).astype(np.int64)
            cells = np.unique(np.stack([np.full(nb, b), z, y, x], 1), axis=0)      # sorted by (b, z, y, x), no duplicates
            idx.append(cells)
        idx = np.concatenate(idx)
        feat = detfill.uniform("%s/feat%d" % (name, s), (len(idx), c))
        tensors.append((torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(feat)))
    return dict(tensors=tensors, calib=calib, image_shape=image_shape, img_feat={"layer1_ori_feat2d": img_feat})
