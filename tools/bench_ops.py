"""Per-kernel micro-benchmarks with roofline fractions (GPU box). Not the headline bench (bench.py).

    python tools/bench_ops.py msda [--shape ctf|c1|ccp|kitti] [--iters 20]

Timing: CUDA events on the launching stream, >=3 warm-ups, L2 flushed between iterations by
writing a 256 MiB buffer. Algorithmic bytes per SURVEY.md section 8(d).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_cuda(fn, iters, flush=True):
    fbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            fbuf.fill_(1)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


MSDA_SHAPES = {
    # name: N, H, W, M, D, Lq, P   
    "c1": (6, 64, 176, 8, 16, 3500, 4),
    "ctf": (12, 112, 200, 8, 16, 8000, 4),
    "ccp": (24, 150, 267, 8, 16, 6000, 4),
    "kitti": (2, 94, 311, 8, 8, 20000, 4),
}


def bench_msda(args):
    from ddf_b200.ops import msda
    N, H, W, M, D, Lq, P = MSDA_SHAPES[args.shape]
    dev = "cuda"
    torch.manual_seed(0)
    S = H * W
    value = torch.randn(N, S, M, D, device=dev)
    shapes = torch.tensor([[H, W]], dtype=torch.long, device=dev)
    lsi = torch.zeros(1, dtype=torch.long, device=dev)
    if args.loc == "uniform":
        loc = torch.rand(N, Lq, M, 1, P, 2, device=dev)
    else:  # clustered: sorted reference points + N(0, 2px) offsets, like projected voxel centres
        ref = torch.rand(N, Lq, 1, 1, 1, 2, device=dev)
        ref, _ = torch.sort(ref, dim=1)
        loc = ref + torch.randn(N, Lq, M, 1, P, 2, device=dev) * torch.tensor([2.0 / W, 2.0 / H], device=dev)
        loc = loc.contiguous()
    attn = torch.softmax(torch.randn(N, Lq, M, P, device=dev), -1).view(N, Lq, M, 1, P).contiguous()
    gout = torch.randn(N, Lq, M * D, device=dev)
    hbm, how = peaks()
    fwd_bytes = 4 * (N * S * M * D + 3 * N * Lq * M * P + N * Lq * M * D)
    bwd_bytes = 4 * (2 * N * S * M * D + 2 * N * Lq * M * D + 2 * 3 * N * Lq * M * P)
    med, best = time_cuda(lambda: msda.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), args.iters)
    print(json.dumps(dict(kernel="msda_fwd", shape=args.shape, loc=args.loc, ms_median=med, ms_best=best,
                          alg_MB=fwd_bytes / 1e6, GBs=fwd_bytes / med / 1e6, frac=fwd_bytes / med / 1e6 / hbm, peak=how)))
    med, best = time_cuda(lambda: msda.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64), args.iters)
    print(json.dumps(dict(kernel="msda_bwd", shape=args.shape, loc=args.loc, ms_median=med, ms_best=best,
                          alg_MB=bwd_bytes / 1e6, GBs=bwd_bytes / med / 1e6, frac=bwd_bytes / med / 1e6 / hbm, peak=how)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["msda"])
    ap.add_argument("--shape", default="ctf")
    ap.add_argument("--loc", default="clustered", choices=["uniform", "clustered"])
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dict(msda=bench_msda)[a.op](a)
