"""Quick numerics + timing check of the tcgen05 sparse-conv path against the C oracle (GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from test_oracle_spconv import random_voxels
from ddf_b200.ops.spconv import ops, functional as Fsp
from oracle import spconv as osp

def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())

shape = [11, 80, 80]
idx = random_voxels(20000, 2, shape, seed=4)
for subm, ks, st, pad in [(True, [3]*3, [1]*3, [1]*3), (False, [3]*3, [2]*3, [1]*3)]:
    o_out, o_pairs, o_num, _ = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, [1]*3, subm, order="gpu")
    rb = ops.build_rulebook(torch.from_numpy(idx).cuda(), 2, shape, ks, st, pad, 1, 0, subm, False)
    for cin, cout in [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128), (8, 16), (24, 48)]:
        rng = np.random.default_rng(1)
        feat = rng.standard_normal((len(idx), cin)).astype(np.float32)
        w = (rng.standard_normal((*ks, cin, cout)) / np.sqrt(cin * 9)).astype(np.float32)
        go = rng.standard_normal((len(o_out), cout)).astype(np.float32)
        ref = osp.indice_conv(feat, w, o_pairs, o_num, len(o_out))
        ref_gi, ref_gw = osp.indice_conv_backward(feat, w, go, o_pairs, o_num)
        f = torch.from_numpy(feat).cuda().requires_grad_(); wt = torch.from_numpy(w).cuda().requires_grad_()
        out = Fsp.table_conv(f, wt, None, rb, len(o_out)); out.backward(torch.from_numpy(go).cuda())
        torch.cuda.synchronize()
        print("subm=%d %3d->%3d  fwd %.2e  dgrad %.2e  wgrad %.2e" % (subm, cin, cout, rel(out.detach().cpu().numpy(), ref),
              rel(f.grad.cpu().numpy(), ref_gi), rel(wt.grad.cpu().numpy(), ref_gw)), flush=True)
