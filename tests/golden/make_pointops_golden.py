"""Capture the known-answer vectors of the reference's own point-op tests into
tests/golden/pointops_golden.npz.  Run in the build container only:

    python tests/golden/make_pointops_golden.py

The reference test file (TransFusion/tests/test_models/test_common_modules/test_pointnet_ops.py) is
executed unmodified with `mmdet3d.ops` replaced by the numpy ORACLE (oracle/pointops.py) and
`.cuda()` made a no-op: its own `assert torch.all(idx == expected_idx)` statements therefore check
the oracle against the reference's expected values; every (inputs, outputs) pair that passed is
recorded as a fixture.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointops as op  # noqa: E402

REF_TEST = "/root/reference/TransFusion/tests/test_models/test_common_modules/test_pointnet_ops.py"
OUT = os.path.join(ROOT, "tests", "golden", "pointops_golden.npz")
REC = {}
COUNT = {}


def rec(name, **arrs):
    i = COUNT.get(name, 0)
    COUNT[name] = i + 1
    for k, v in arrs.items():
        REC["%s%d/%s" % (name, i, k)] = np.asarray(v)


def fps(xyz, n):
    out = op.furthest_point_sample(xyz.numpy(), n)
    rec("fps", xyz=xyz.numpy(), npoint=n, idx=out)
    return torch.from_numpy(out)


def ball_query(min_r, max_r, ns, xyz, new_xyz):
    out = op.ball_query(min_r, max_r, ns, xyz.numpy(), new_xyz.numpy())
    rec("ball_query", xyz=xyz.numpy(), new_xyz=new_xyz.numpy(), min_r=min_r, max_r=max_r, nsample=ns, idx=out)
    return torch.from_numpy(out)


def grouping_operation(features, idx):
    out = op.grouping_operation(features.numpy(), idx.numpy())
    rec("group", features=features.numpy(), idx=idx.numpy(), out=out)
    return torch.from_numpy(out)


def gather_points(features, idx):
    out = op.gather_points(features.numpy(), idx.numpy())
    rec("gather", features=features.numpy(), idx=idx.numpy(), out=out)
    return torch.from_numpy(out)


def main():
    stub = types.ModuleType("mmdet3d.ops")
    for n, f in dict(ball_query=ball_query, furthest_point_sample=fps, gather_points=gather_points,
                     grouping_operation=grouping_operation).items():
        setattr(stub, n, f)
    for n in ("furthest_point_sample_with_dist", "knn", "three_interpolate", "three_nn"):
        setattr(stub, n, None)  # off-path ops
    sys.modules["mmdet3d"] = types.ModuleType("mmdet3d")
    sys.modules["mmdet3d.ops"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    ns = {}
    exec(compile(open(REF_TEST).read(), REF_TEST, "exec"), ns)
    for t in ("test_fps", "test_ball_query", "test_grouping_points", "test_gather_points"):
        ns[t]()   # raises AssertionError if the oracle disagrees with the reference's expected values
        print("reference", t, "passed against the oracle")
    np.savez_compressed(OUT, **REC)
    print("wrote", OUT, sorted(COUNT.items()))


if __name__ == "__main__":
    main()
