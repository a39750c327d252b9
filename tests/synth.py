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
