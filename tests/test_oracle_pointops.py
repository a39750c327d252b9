"""Oracle pinning (CPU): oracle/pointops.py against the known-answer vectors captured from the
reference's own test file (tests/golden/make_pointops_golden.py executes
TransFusion/tests/test_models/test_common_modules/test_pointnet_ops.py with the oracle plugged in)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import pointops as op

G = np.load(os.path.join(GOLDEN, "pointops_golden.npz"))


def test_fps_known_answer():
    idx = op.furthest_point_sample(G["fps0/xyz"], int(G["fps0/npoint"]))
    assert np.array_equal(idx, G["fps0/idx"]) and np.array_equal(idx, [[0, 2, 4], [0, 2, 1]])


def test_ball_query_known_answers():
    for i in range(2):
        p = "ball_query%d/" % i
        idx = op.ball_query(float(G[p + "min_r"]), float(G[p + "max_r"]), int(G[p + "nsample"]), G[p + "xyz"], G[p + "new_xyz"])
        assert np.array_equal(idx, G[p + "idx"])
    assert G["ball_query1/idx"][0, 0].tolist() == [0, 5, 7, 0, 0]  # dilated query, padded with the first hit


def test_group_gather_known_answers_and_grads():
    out = op.grouping_operation(G["group0/features"], G["group0/idx"])
    assert np.array_equal(out, G["group0/out"])
    out = op.gather_points(G["gather0/features"], G["gather0/idx"])
    assert np.array_equal(out, G["gather0/out"])
    # grads are the adjoint scatter-adds
    rng = np.random.default_rng(0)
    f = rng.standard_normal((2, 3, 10)).astype(np.float32)
    idx = rng.integers(0, 10, (2, 4, 5)).astype(np.int32)
    g = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    lhs = (op.grouping_operation(f, idx) * g).sum()
    rhs = (f * op.grouping_operation_grad(g, idx, 10)).sum()
    assert abs(lhs - rhs) < 1e-4


def test_fps_tie_break_follows_reference_thread_layout():
    # 6 identical points + 2 distinct: block size B = 8; the reference's pairwise tree resolves ties by the low bits
    # of the thread index first = smallest bit-reversed (k mod B), then lowest k (checked against the compiled
    # reference kernel on the GPU box: tests/test_reference_cuda_gpu.py::test_fps_tiny_tie_case)
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, 5] = [1, 0, 0]
    xyz[0, 6] = [1, 0, 0]
    idx = op.furthest_point_sample(xyz, 3)
    # after picking 0, points 5 (0b101 -> reversed 0b101) and 6 (0b110 -> reversed 0b011) tie at distance 1 -> 6;
    # then everything is at distance 0 from the picked set -> reversed index 0 (k = 0)
    assert idx.tolist() == [[0, 6, 0]]
