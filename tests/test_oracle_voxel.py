"""Oracle pinning (CPU): oracle/voxel_ref.c against (i) the reference's known-answer vector
(TransFusion/tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22) and (ii) the
reference's own voxel_layer extension built unmodified into oracle/_ref (when present)."""
import numpy as np
import pytest

import synth
from oracle import ref_build, voxel


def test_known_answer_vector_of_reference_test():
    np.random.seed(0)
    pts = np.random.rand(1000, 4)
    v, c, n = voxel.hard_voxelize(pts, [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    expected = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1],
                         [7, 80, 0], [6, 80, 0]])
    assert np.array_equal(c, expected)
    assert np.array_equal(n, [120, 121, 127, 134, 115, 127, 125, 131])
    assert v.shape == (8, 1000, 4)
    # points inside each voxel keep input order
    p32 = pts.astype(np.float32)
    first = p32[(np.floor((p32[:, 2] + 3) / 0.5) == 7) & (np.floor((p32[:, 1] + 40) / 0.5) == 81)
                & (np.floor(p32[:, 0] / 0.5) == 1)]
    assert np.array_equal(v[0, :120], first)


CASES = [
    dict(n=20000, kind="lidar", max_points=10, max_voxels=120000),
    dict(n=20000, kind="uniform", max_points=10, max_voxels=120000),
    dict(n=30000, kind="lidar", max_points=3, max_voxels=5000),     # both caps bite
    dict(n=5000, kind="uniform", max_points=1, max_voxels=100),
    dict(n=1, kind="lidar", max_points=10, max_voxels=10),
]


def make_points(c):
    if c["kind"] == "lidar":
        return synth.lidar_points(c["n"], seed=1)
    return synth.uniform_points(c["n"], synth.NUSC_RANGE, seed=2)


@pytest.mark.parametrize("c", CASES)
def test_oracle_equals_reference_extension(c):
    ext = ref_build.load("voxel_layer")
    if ext is None:
        pytest.skip("oracle/_ref/voxel_layer.so not built (needs /root/reference)")
    import torch
    pts = make_points(c)
    v, co, n = voxel.hard_voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, c["max_points"], c["max_voxels"])
    tp = torch.from_numpy(pts)
    rv = torch.zeros(c["max_voxels"], c["max_points"], pts.shape[1])
    rc = torch.zeros(c["max_voxels"], 3, dtype=torch.int32)
    rn = torch.zeros(c["max_voxels"], dtype=torch.int32)
    m = ext.hard_voxelize(tp, rv, rc, rn, synth.NUSC_VOXEL, synth.NUSC_RANGE, c["max_points"], c["max_voxels"], 3)
    assert m == len(co)
    assert np.array_equal(rc[:m].numpy(), co)
    assert np.array_equal(rn[:m].numpy(), n)
    assert np.array_equal(rv[:m].numpy(), v)
    dc = torch.zeros(len(pts), 3, dtype=torch.int32)
    ext.dynamic_voxelize(tp, dc, synth.NUSC_VOXEL, synth.NUSC_RANGE, 3)
    assert np.array_equal(dc.numpy(), voxel.dynamic_voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE))


def test_max_voxels_break_drops_later_points_of_open_voxels():
    # 3 points: A (voxel 0), B (voxel 1 -> would be number max_voxels=1 -> break), A' (voxel 0 again)
    pts = np.array([[0.1, 0.1, 0.1, 1], [5.1, 0.1, 0.1, 2], [0.2, 0.1, 0.1, 3]], np.float32)
    v, c, n = voxel.hard_voxelize(pts, [1, 1, 1], [0, 0, 0, 10, 10, 10], 5, 1)
    assert len(c) == 1 and n[0] == 1 and v[0, 0, 3] == 1 and v[0, 1, 3] == 0
