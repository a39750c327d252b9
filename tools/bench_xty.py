"""Times ddf_xty_tf32 against the library product (a.t() @ b, tf32 allowed) on the Linear weight-gradient shapes of the
bench step (T = 146016 tokens)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
from bench_ops import time_cuda, peaks
from ddf_b200.ops import fused
torch.backends.cuda.matmul.allow_tf32 = True
hbm, how = peaks()
for K, M, N in [(146016, 1024, 128), (146016, 128, 1024), (146016, 128, 128), (146016, 256, 128), (146016, 128, 256)]:
    a = torch.randn(K, M, device="cuda"); b = torch.randn(K, N, device="cuda")
    byts = 4.0 * K * (M + N)
    for name, fn in (("ddf_xty_tf32", lambda: fused.xty(a, b)), ("library a.t() @ b", lambda: a.t() @ b)):
        med, best = time_cuda(fn, 10)
        print(json.dumps(dict(kernel=name, K=K, M=M, N=N, ms_median=round(med, 4), GBs=round(byts / med / 1e6, 1),
                              hbm_frac=round(byts / med / 1e6 / hbm, 3), TFLOPs=round(2.0 * K * M * N / med / 1e9, 1))))
