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


WARM = 3


def time_cuda(fn, iters, flush=True):
    fbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(WARM):
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


    if P == 4 and msda.tile_supported(M, D, 1, P):
        # tile-staged dual-query form (the hot path): raw offsets + logits + reference points
        refp = torch.rand(N, Lq, 2, device=dev)
        off = torch.randn(N, Lq, M, 1, P, 2, device=dev) * 2.0
        logit = torch.randn(N, Lq, M, P, device=dev)
        med, best = time_cuda(lambda: msda.TilePlan(refp, H, W), args.iters)
        print(json.dumps(dict(kernel="msda_tile_plan", shape=args.shape, ms_median=med, ms_best=best)))
        plan = msda.TilePlan(refp, H, W)
        med, best = time_cuda(lambda: msda.msda_tile_forward(value, plan, off, logit), args.iters)
        print(json.dumps(dict(kernel="msda_tile_fwd", shape=args.shape, ms_median=med, ms_best=best,
                              alg_MB=fwd_bytes / 1e6, GBs=fwd_bytes / med / 1e6, frac=fwd_bytes / med / 1e6 / hbm, peak=how)))
        med, best = time_cuda(lambda: msda.msda_tile_backward(value, plan, off, logit, gout), args.iters)
        print(json.dumps(dict(kernel="msda_tile_bwd", shape=args.shape, ms_median=med, ms_best=best,
                              alg_MB=bwd_bytes / 1e6, GBs=bwd_bytes / med / 1e6, frac=bwd_bytes / med / 1e6 / hbm, peak=how)))


def _tf_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), "measured bf16 burst"
    return 1590.0, "fallback bf16"


def _scene(args):
    """Voxelised synthetic nuScenes sweep(s): int32 [N,4] (b,z,y,x) indices + the points."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from ddf_b200.ops.voxel import Voxelization
    vox = Voxelization(synth.NUSC_VOXEL, synth.NUSC_RANGE, 10, (120000, 160000)).cuda().train()
    idx, pts_all = [], []
    for b in range(args.batch):
        pts = torch.from_numpy(synth.lidar_points(args.points, seed=b)).cuda()
        _, coors, _ = vox(pts)
        idx.append(torch.nn.functional.pad(coors, (1, 0), value=b))
        pts_all.append(pts)
    return torch.cat(idx).contiguous(), pts_all


def bench_voxel(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from ddf_b200.ops import voxel
    pts = torch.from_numpy(synth.lidar_points(args.points, seed=0)).cuda()
    n, f = pts.shape
    T, cap = 10, 120000
    voxels = pts.new_empty((cap, T, f)); coors = pts.new_empty((cap, 3), dtype=torch.int)
    npv = pts.new_empty((cap,), dtype=torch.int)
    m = voxel.hard_voxelize(pts, voxels, coors, npv, synth.NUSC_VOXEL, synth.NUSC_RANGE, T, cap)
    hbm, how = peaks()
    byts = 4 * n * f + m * (4 * T * f + 16)
    med, best = time_cuda(lambda: voxel.hard_voxelize_device(pts, voxels, coors, npv, synth.NUSC_VOXEL,
                                                             synth.NUSC_RANGE, T, cap), args.iters)
    print(json.dumps(dict(kernel="hard_voxelize (all launches)", points=n, voxels=m, ms_median=med, ms_best=best,
                          alg_MB=byts / 1e6, GBs=byts / med / 1e6, frac=byts / med / 1e6 / hbm, peak=how)))


STAGES = [  # (name, subm, ksize, stride, padding, cin, cout) along the TransFusion sparse encoder
    ("subm1 16->16", True, [3] * 3, [1] * 3, [1] * 3, 16, 16),
    ("down1 16->32", False, [3] * 3, [2] * 3, [1] * 3, 16, 32),
    ("subm2 32->32", True, [3] * 3, [1] * 3, [1] * 3, 32, 32),
    ("down2 32->64", False, [3] * 3, [2] * 3, [1] * 3, 32, 64),
    ("subm3 64->64", True, [3] * 3, [1] * 3, [1] * 3, 64, 64),
    ("down3 64->128", False, [3] * 3, [2] * 3, [0, 1, 1], 64, 128),
    ("subm4 128->128", True, [3] * 3, [1] * 3, [1] * 3, 128, 128),
]


def bench_spconv(args):
    """Rulebook build + conv fwd / dgrad / wgrad per stage of the encoder on the real active sets."""
    from ddf_b200.ops.spconv import ops
    idx, _ = _scene(args)
    shape = [41, 1440, 1440]
    hbm, how = peaks()
    tfp, tfhow = _tf_peak()
    for name, subm, ks, st, pad, cin, cout in STAGES:
        if args.stages and subm and not any(t in name for t in args.stages.split(",")):
            continue
        n = idx.shape[0]
        build = lambda: ops.build_rulebook(idx, args.batch, shape, ks, st, pad, 1, 0, subm, False)
        rb = build()
        n_out = rb.outids.shape[0]
        pairs = int(rb.indice_pair_num.sum())
        med, best = time_cuda(build, args.iters)
        rb_bytes = 16 * n + 8 * pairs + 16 * n_out
        print(json.dumps(dict(kernel="rulebook " + name, n_in=n, n_out=n_out, pairs=pairs, ms_median=med,
                              alg_MB=rb_bytes / 1e6, GBs=rb_bytes / med / 1e6, frac=rb_bytes / med / 1e6 / hbm,
                              peak=how)))
        feat = ops.round_tf32(torch.randn(n, cin, device="cuda"))
        w = torch.randn(*ks, cin, cout, device="cuda") / (cin * 27) ** 0.5
        go = ops.round_tf32(torch.randn(n_out, cout, device="cuda"))
        flops = 2.0 * pairs * cin * cout
        byts = 4.0 * (n * cin + n_out * cout + 27 * cin * cout) + 8.0 * pairs
        mode = ops.tc_mode(27, cin, cout)
        # bf16x3 layers take pre-split operands (the split pass is timed separately as "split")
        f_in, f_fmt = (ops.split_bf16x3(feat)[0], 1) if mode & 16 else (feat, 0)
        g_in, g_fmt = (ops.split_bf16x3(go)[0], 1) if mode & 32 else (go, 0)
        if mode & 16:
            med, best = time_cuda(lambda: ops.split_bf16x3(feat, want_rounded=True), args.iters)
            print(json.dumps(dict(kernel="split_bf16x3 %s" % name, ms_median=med, alg_MB=12.0 * n * cin / 1e6,
                                  GBs=12.0 * n * cin / med / 1e6, hbm_frac=12.0 * n * cin / med / 1e6 / hbm)))
        for kind, fn in (("fwd", lambda: ops.sparse_conv_forward(f_in, w, rb.gather_table, None, n_out, f_fmt)),
                         ("dgrad", lambda: ops.sparse_conv_dgrad(w, g_in, rb.scatter_table, n, g_fmt)),
                         ("wgrad", lambda: ops.sparse_conv_wgrad(feat, w, go, rb.indice_pairs, rb.indice_pair_num)),
                         ("wgrad_table", (lambda: ops.sparse_conv_wgrad_table(feat, w, go, rb.gather_table))
                          if (subm and ops.tc_mode(27, cin, cout) & 8) else None)):
            if fn is None:
                continue
            med, best = time_cuda(fn, args.iters)
            print(json.dumps(dict(kernel="spconv %s %s" % (kind, name), ms_median=med, ms_best=best,
                                  GFLOP=flops / 1e9, TFLOPs=flops / med / 1e9, tensor_frac_bf16=flops / med / 1e9 / tfp,
                                  tensor_frac_tf32=flops / med / 1e9 / (tfp / 2), alg_MB=byts / 1e6,
                                  GBs=byts / med / 1e6, hbm_frac=byts / med / 1e6 / hbm, peak="%s; %s" % (how, tfhow))))
        if not subm:
            idx = rb.outids
            shape = rb.out_spatial_shape


def bench_dense(args):
    from ddf_b200.ops.spconv.structure import SparseConvTensor
    n, C, B, D, H, W = 60000 * args.batch, 128, args.batch, 2, 180, 180
    g = torch.Generator(device="cuda").manual_seed(0)
    flat = torch.randperm(B * D * H * W, device="cuda", generator=g)[:n].sort().values
    idx = torch.stack([flat // (D * H * W), flat // (H * W) % D, flat // W % H, flat % W], 1).int().contiguous()
    feat = torch.randn(n, C, device="cuda")
    t = SparseConvTensor(feat, idx, [D, H, W], B)
    hbm, how = peaks()
    byts = 4 * n * C + 16 * n + 4 * B * C * D * H * W
    med, best = time_cuda(lambda: t.dense(), args.iters)
    print(json.dumps(dict(kernel="sparse_to_dense (memset + scatter)", n=n, ms_median=med, ms_best=best,
                          alg_MB=byts / 1e6, GBs=byts / med / 1e6, frac=byts / med / 1e6 / hbm, peak=how)))


def bench_pointops(args):
    from ddf_b200.ops import pointops
    Bp, N, m, ns, C = 12, 8000, 2048, 32, 128
    torch.manual_seed(0)
    xyz = (torch.rand(Bp, N, 3, device="cuda") * torch.tensor([108.0, 108.0, 8.0], device="cuda")).contiguous()
    feats = torch.randn(Bp, C, N, device="cuda")
    hbm, how = peaks()
    med, best = time_cuda(lambda: pointops.furthest_point_sample(xyz, m), args.iters)
    idx = pointops.furthest_point_sample(xyz, m)
    byts = 12 * Bp * N
    print(json.dumps(dict(kernel="furthest_point_sample", rows=Bp, N=N, m=m, ms_median=med, us_per_row=med * 1e3 / Bp,
                          alg_MB=byts / 1e6, GBs=byts / med / 1e6, frac=byts / med / 1e6 / hbm, note="latency-bound")))
    centres = pointops.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    med, best = time_cuda(lambda: pointops.ball_query(0.0, 2.0, ns, xyz, centres), args.iters)
    bq = pointops.ball_query(0.0, 2.0, ns, xyz, centres)
    byts = 12 * Bp * (N + m) + 4 * Bp * m * ns
    print(json.dumps(dict(kernel="ball_query", ms_median=med, alg_MB=byts / 1e6, GBs=byts / med / 1e6,
                          frac=byts / med / 1e6 / hbm, peak=how)))
    med, best = time_cuda(lambda: pointops.grouping_operation(feats, bq), args.iters)
    byts = 4 * Bp * C * m * ns * 2
    print(json.dumps(dict(kernel="group_points", ms_median=med, alg_MB=byts / 1e6, GBs=byts / med / 1e6,
                          frac=byts / med / 1e6 / hbm, peak=how)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["msda", "voxel", "spconv", "dense", "pointops", "all"])
    ap.add_argument("--shape", default="ctf")
    ap.add_argument("--loc", default="clustered", choices=["uniform", "clustered"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--points", type=int, default=260000)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--stages", default="", help="spconv: comma list of SubM stage names to run (strided convs always run)")
    ap.add_argument("--warm", type=int, default=3, help="untimed warm-up launches (0 for ncu captures)")
    ap.add_argument("--tc-mode", type=int, default=4, help="conv kernels: 4 bf16x3 fwd/dgrad on TMA+tcgen05 (default), 1 tf32 TMA+tcgen05, 2 cp.async+tcgen05, 0 fp32 SIMT")
    a = ap.parse_args()
    WARM = a.warm
    from ddf_b200 import lib as _l
    _l.get_lib().ddf_set_tensor_cores(a.tc_mode)
    table = dict(msda=bench_msda, voxel=bench_voxel, spconv=bench_spconv, dense=bench_dense, pointops=bench_pointops)
    for name in (table if a.op == "all" else [a.op]):
        table[name](a)
