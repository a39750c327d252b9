"""Fused BatchNorm(+ReLU) forward / backward per stage of the bench workload's sparse encoder: ms and fraction of the
HBM rate (algorithmic bytes: forward 2 reads + 1 write, backward 5 reads + 1 write of N*C*4)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
from bench_ops import time_cuda, peaks
from ddf_b200.ops import sparse_norm
hbm, how = peaks()
for n, C in [(240000, 16), (647219, 32), (427653, 64), (135502, 128)]:
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    x = torch.randn(n, C, device="cuda", requires_grad=True)
    go = torch.randn(n, C, device="cuda")
    y = sparse_norm.batch_norm_act(bn, x, relu=True)
    med, _ = time_cuda(lambda: sparse_norm.batch_norm_act(bn, x, relu=True), 10)
    fb = 3.0 * 4 * n * C
    print(json.dumps(dict(kernel="bn+relu fwd", n=n, C=C, ms=round(med, 4), GBs=round(fb / med / 1e6, 1), hbm_frac=round(fb / med / 1e6 / hbm, 3))))
    def bwd():
        x.grad = None
        y.backward(go, retain_graph=True)
    med, _ = time_cuda(bwd, 10)
    bb = 6.0 * 4 * n * C
    print(json.dumps(dict(kernel="bn+relu bwd", n=n, C=C, ms=round(med, 4), GBs=round(bb / med / 1e6, 1), hbm_frac=round(bb / med / 1e6 / hbm, 3))))
