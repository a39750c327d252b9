"""Generate tests/golden/msda_golden.npz from the REFERENCE's own pure-PyTorch MSDA
(`ms_deform_attn_core_pytorch`, <proj>/models/model_utils/ops/functions/ms_deform_attn_func.py:41-61)
imported from /root/reference (read-only). Run in the build container only:

    python tests/golden/make_msda_golden.py

The first case reproduces the reference's own test script (ops/test.py:21-36: N,M,D=1,2,2;
Lq,L,P=2,2,2; shapes (6,4),(3,2); torch.manual_seed(3); value=rand*0.01; weights normalised).
Gradients come from autograd through the reference function with a seeded grad_output.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/TransFusion/mmdet3d/models/model_utils/ops/functions/ms_deform_attn_func.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msda_golden.npz")


def load_ref():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("ref_msda_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def make_case(core, name, N, M, D, Lq, P, shapes, seed, dtype, loc_lo=0.0, loc_hi=1.0, out=None):
    torch.manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = len(shapes)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    S = int(shapes_t.prod(1).sum())
    value = (torch.rand(N, S, M, D) * 0.01).to(dtype)
    loc = (torch.rand(N, Lq, M, L, P, 2) * (loc_hi - loc_lo) + loc_lo).to(dtype)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dtype)
    gout = torch.randn(N, Lq, M * D).to(dtype)
    value.requires_grad_(True)
    loc.requires_grad_(True)
    attn.requires_grad_(True)
    o = core(value, shapes_t, loc, attn)
    o.backward(gout)
    out.update({
        name + "/value": value.detach().numpy(), name + "/shapes": shapes_t.numpy(),
        name + "/lsi": lsi.numpy(), name + "/loc": loc.detach().numpy(),
        name + "/attn": attn.detach().numpy(), name + "/gout": gout.numpy(),
        name + "/out": o.detach().numpy(), name + "/gvalue": value.grad.numpy(),
        name + "/gloc": loc.grad.numpy(), name + "/gattn": attn.grad.numpy(),
    })


def main():
    core = load_ref()
    out = {}
    # reference's own test shapes (ops/test.py)
    make_case(core, "reftest_f64", 1, 2, 2, 2, 2, [(6, 4), (3, 2)], 3, torch.float64, out=out)
    make_case(core, "reftest_f32", 1, 2, 2, 2, 2, [(6, 4), (3, 2)], 3, torch.float32, out=out)
    # hot-path head layout (M=8, D=16, L=1, P=4) incl. out-of-map sampling points
    make_case(core, "hot_d16_f32", 2, 8, 16, 37, 4, [(8, 12)], 11, torch.float32, -0.2, 1.2, out=out)
    # Voxel-RCNN head layout (d_model 64 -> D=8), multi-level, odd P
    make_case(core, "kitti_d8_f32", 1, 8, 8, 19, 3, [(9, 14), (5, 7), (3, 4)], 5, torch.float32, -0.1, 1.1, out=out)
    # generic (non multiple-of-4) channel count from the reference gradcheck list (D=30)
    make_case(core, "generic_d30_f64", 1, 2, 30, 5, 2, [(6, 4), (3, 2)], 7, torch.float64, out=out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
