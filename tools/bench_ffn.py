"""Fused FFN forward (ddf_ffn_forward) against the chain it replaces (library GEMM, in-place bias / ReLU / dropout
kernel, library GEMM) on the bench step's shape: T = 146016 tokens, d_model 128, d_ffn 1024."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
from torch import nn
from bench_ops import time_cuda, peaks
from ddf_b200.ops import fused
torch.backends.cuda.matmul.allow_tf32 = True
hbm, how = peaks()
T, D, F_ = 146016, 128, 1024
l1, l2, drop = nn.Linear(D, F_).cuda(), nn.Linear(F_, D).cuda(), nn.Dropout(0.1).train()
x = torch.randn(T, D, device="cuda")
with torch.no_grad():
    for name, fn in (("fused ffn forward", lambda: fused.ffn(l1, drop, l2, x)),
                     ("chain: GEMM + bias/ReLU/dropout + GEMM", lambda: fused.linear(l2, fused.ffn_hidden(l1, drop, x)))):
        med, best = time_cuda(fn, 10)
        wr = 4.0 * T * (F_ + 2 * D)
        print(json.dumps(dict(kernel=name, ms_median=round(med, 4), write_GBs=round(wr / med / 1e6, 1),
                              hbm_frac=round(wr / med / 1e6 / hbm, 3), TFLOPs=round(4.0 * T * D * F_ / med / 1e9, 1))))
