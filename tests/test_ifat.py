"""IFAT image gate (fusion/ifat.py) against golden vectors produced by the reference's own class
(tests/golden/make_ifat_golden.py -> ifat_golden.npz): same state dict, same inputs, fp32 tolerance 1e-5;
pure PyTorch module, runs on the CPU. The CenterPoint wrapper wiring is exercised on the GPU."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

CFG = dict(img_num_channel=8, pts_num_channel=8, voxel_feat_channel=[4, 6, 8], voxel_idx=[0, 2])


def _gate(g):
    from ddf_b200.fusion.ifat import Basicgate_patch_iv_multivoxel
    gate = Basicgate_patch_iv_multivoxel(**CFG)
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    assert sorted(sd) == sorted(gate.state_dict())          # reference checkpoints load unchanged
    gate.load_state_dict(sd)
    return gate.eval()


@pytest.mark.parametrize("case", ["unique", "dup"])
def test_gate_matches_reference_class(case):
    g = np.load(os.path.join(GOLDEN, "ifat_golden.npz"))
    gate = _gate(g)
    img = torch.from_numpy(g[case + "_img"])[None]
    H, W = img.shape[-2:]
    feats = {s: torch.from_numpy(g["%s_feat%d" % (case, s)]) for s in range(3)}
    coords = {s: torch.from_numpy(g["%s_coord%d" % (case, s)]) for s in range(3)}
    cells = {s: torch.from_numpy(g["%s_grid%d" % (case, s)]).long() for s in range(3)}
    cells = {s: c[:, 1] * W + c[:, 0] for s, c in cells.items()}
    with torch.no_grad():
        out = gate(img, feats, cells, coords)[0]
    ref = torch.from_numpy(g[case + "_out"])
    assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max())


def test_groups_are_independent_and_gradients_flow():
    g = np.load(os.path.join(GOLDEN, "ifat_golden.npz"))
    gate = _gate(g)
    torch.manual_seed(1)
    H, W, G = 6, 9, 3
    img = torch.randn(G, 8, H, W)
    feats, cells, coords = {}, {}, {}
    for s, c in ((0, 4), (2, 8)):
        n = 40
        grp = torch.randint(0, G, (n,)).sort().values
        pix = torch.randint(0, H * W, (n,))
        feats[s] = torch.randn(n, c, requires_grad=True)
        coords[s] = torch.randn(n, 3)
        cells[s] = grp * H * W + pix
    out = gate(img, feats, cells, coords)
    for k in range(G):      # a batch of planes == the planes one by one (the reference loops over them)
        sel = {s: (cells[s] // (H * W)) == k for s in cells}
        one = gate(img[k:k + 1], {s: feats[s][sel[s]] for s in feats}, {s: cells[s][sel[s]] - k * H * W for s in cells},
                   {s: coords[s][sel[s]] for s in coords})
        assert torch.allclose(out[k], one[0], atol=1e-6)
    out.sum().backward()
    assert all(float(f.grad.abs().sum()) > 0 for f in feats.values())
