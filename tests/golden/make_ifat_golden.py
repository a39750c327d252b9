"""Golden vectors for the IFAT image gate from the REFERENCE class (run in the build container, where
/root/reference exists): CenterPoint/det3d/models/model_utils/attention.py Basicgate_patch_iv_multivoxel.
The reference file is imported with its package neighbours stubbed (losses.auxseg_loss) and
``Tensor.cuda`` made a no-op (pts2img calls .cuda() on a shape tensor); it runs on the CPU."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/CenterPoint/det3d/models/model_utils/attention.py"


def load_reference():
    for name in ("det3d", "det3d.models", "det3d.models.model_utils", "det3d.models.losses"):
        sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules[name].__path__ = []
    seg = types.ModuleType("det3d.models.losses.auxseg_loss")
    seg.SEGLOSS = object
    sys.modules["det3d.models.losses.auxseg_loss"] = seg
    spec = importlib.util.spec_from_file_location("det3d.models.model_utils.attention", REF)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "det3d.models.model_utils"
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    mod = load_reference()
    torch.manual_seed(0)
    cfg = dict(img_num_channel=8, pts_num_channel=8, voxel_feat_channel=[4, 6, 8], voxel_idx=[0, 2])
    gate = mod.Basicgate_patch_iv_multivoxel(**cfg)
    H, W = 12, 20
    out = {}
    for case, dup in (("unique", False), ("dup", True)):
        img = torch.randn(8, H, W)
        feats, grids, coords = [], [], []
        for s, (c, n) in enumerate(zip(cfg["voxel_feat_channel"], (60, 40, 30))):
            cells = torch.randperm(H * W)[:n]
            if dup:
                cells[n // 2:] = cells[:n - n // 2]      # second half lands on the pixels of the first half
            grids.append(torch.stack([cells % W, cells // W], 1))      # (x, y)
            feats.append(torch.randn(n, c))
            coords.append(torch.randn(n, 3))
        with torch.no_grad():
            res = gate(img, feats, grids, coords, None, None, 0, None)
        out[case + "_img"] = img.numpy()
        out[case + "_out"] = res.numpy()
        for s in range(3):
            out["%s_feat%d" % (case, s)] = feats[s].numpy()
            out["%s_grid%d" % (case, s)] = grids[s].numpy()
            out["%s_coord%d" % (case, s)] = coords[s].numpy()
    for k, v in gate.state_dict().items():
        out["w:" + k] = v.numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ifat_golden.npz"), **out)
    print("wrote ifat_golden.npz", sorted(k for k in out if k.startswith("w:")))


if __name__ == "__main__":
    main()
