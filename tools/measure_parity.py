"""Print whole-path parity numbers (GPU product vs the oracle-driven CPU path) per conv precision mode and cuBLAS
tf32 setting; used to calibrate the tolerances asserted in tests/test_hotpath_gpu.py."""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import test_hotpath_gpu as T  # noqa: E402
from ddf_b200 import lib  # noqa: E402
from oracle import cpu_path  # noqa: E402


def run(mode, allow_tf32, n_points, train):
    lib.get_lib().ddf_set_tensor_cores(mode)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    m_cpu = T.build(seed=1)
    m_cpu.train(train)
    for mod in m_cpu.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m_gpu = copy.deepcopy(m_cpu).cuda()
    pts, feats, metas = T.inputs(1, n_points)
    with cpu_path.reference_cpu_ops():
        ref = m_cpu(pts, [feats], metas)
        if train:
            ref.square().mean().backward()
    out = m_gpu([p.cuda() for p in pts], [feats.cuda()], metas)
    res = dict(mode=mode, allow_tf32=allow_tf32, n_points=n_points, train=train, out=T.rel(out.detach().cpu(), ref.detach()))
    if train:
        out.square().mean().backward()
        g_cpu = dict(m_cpu.named_parameters())
        gmax = max(float(p.grad.abs().max()) for p in m_cpu.parameters() if p.grad is not None)
        worst, num, den, mincos = 0.0, 0.0, 0.0, 1.0
        for name, p in m_gpu.named_parameters():
            gc = g_cpu[name].grad
            if gc is None:
                continue
            gd, gc = p.grad.cpu().double(), gc.double()
            num += float((gd - gc).square().sum())
            den += float(gc.square().sum())
            worst = max(worst, float((gd - gc).abs().max() / max(float(gc.abs().max()), 1e-3 * gmax)))
            if float(gc.norm()) > 1e-3 * gmax:
                mincos = min(mincos, float((gd * gc).sum() / (gd.norm() * gc.norm()).clamp_min(1e-300)))
        res.update(grad_worst_max=worst, grad_l2=(num / den) ** 0.5, grad_min_cos=mincos)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    run(4, False, n, True)
    run(4, True, n, True)
    run(4, True, n, False)
    run(1, True, n, True)
