"""The "existing kernel to beat": the reference's own CUDA kernels, rebuilt for sm_100 (oracle/ref_build.py ->
oracle/_ref/*.so: spconv + voxel unmodified; MSDA / point ops with the documented two-line torch-2 fixes), timed on
the same B200 and the same inputs as this library's kernels, with the results compared.

    python tools/bench_reference_kernels.py [--iters 10] > gpurun_out/ref_kernels.jsonl

CUDA events on the launching stream, 2 warm-ups, L2 flushed between iterations. Only tests / tools may load oracle/.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import ref_build  # noqa: E402


def timeit(fn, iters, warm=2):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def emit(op, shape, ref_ms, our_ms, check):
    print(json.dumps(dict(op=op, shape=shape, reference_cuda_ms=ref_ms, ours_ms=our_ms,
                          speedup=(ref_ms / our_ms) if (ref_ms and our_ms) else None, check=check)), flush=True)


def bench_msda(args):
    ext = ref_build.load("MultiScaleDeformableAttention")
    if ext is None:
        return
    from ddf_b200.ops import msda
    for name, (N, H, W, M, D, Lq) in dict(ctf=(12, 112, 200, 8, 16, 8000), ccp=(24, 150, 267, 8, 16, 6000),
                                          kitti=(2, 94, 311, 8, 8, 20000)).items():
        torch.manual_seed(0)
        value = torch.randn(N, H * W, M, D, device="cuda")
        shapes = torch.tensor([[H, W]], device="cuda")
        lsi = torch.zeros(1, dtype=torch.long, device="cuda")
        ref = torch.rand(N, Lq, 2, device="cuda")
        off = torch.randn(N, Lq, M, 1, 4, 2, device="cuda") * 2
        logit = torch.randn(N, Lq, M, 4, device="cuda")
        loc = (ref[:, :, None, None, None, :] + off / torch.tensor([W, H], device="cuda", dtype=torch.float32)).contiguous()
        attn = torch.softmax(logit, -1).view(N, Lq, M, 1, 4).contiguous()
        gout = torch.randn(N, Lq, M * D, device="cuda")
        plan = msda.TilePlan(ref, H, W)
        r_out = ext.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
        o_out = msda.msda_tile_forward(value, plan, off, logit)
        g_out = msda.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
        t_ref = timeit(lambda: ext.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), args.iters)
        t_tile = timeit(lambda: msda.msda_tile_forward(value, plan, off, logit), args.iters)
        t_gen = timeit(lambda: msda.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), args.iters)
        emit("msda forward (tile-staged, fused softmax/loc)", name, t_ref, t_tile, "rel err %.1e" % rel(o_out, r_out))
        emit("msda forward (reference signature)", name, t_ref, t_gen, "rel err %.1e" % rel(g_out, r_out))
        r_gv, r_gl, r_ga = ext.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
        o_gv, o_go, o_glog = msda.msda_tile_backward(value, plan, off, logit, gout)
        g_gv, g_gl, g_ga = msda.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
        t_ref = timeit(lambda: ext.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64), args.iters)
        t_tile = timeit(lambda: msda.msda_tile_backward(value, plan, off, logit, gout), args.iters)
        t_gen = timeit(lambda: msda.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64), args.iters)
        norm = torch.tensor([W, H], device="cuda", dtype=torch.float32)
        emit("msda backward (tile-staged)", name, t_ref, t_tile,
             "grad_value %.1e, grad_offsets vs grad_loc/(W,H) %.1e" % (rel(o_gv, r_gv), rel(o_go, r_gl / norm)))
        emit("msda backward (reference signature)", name, t_ref, t_gen,
             "grad_value %.1e grad_loc %.1e grad_attn %.1e" % (rel(g_gv, r_gv), rel(g_gl, r_gl), rel(g_ga, r_ga)))


def bench_voxel(args):
    ext = ref_build.load("voxel_layer")
    if ext is None:
        return
    import synth
    from ddf_b200.ops import voxel
    for n in (20000, 65536):
        pts = torch.from_numpy(synth.lidar_points(n, seed=0)).cuda()
        T, cap = 10, 120000
        def run_ref():
            v = pts.new_zeros((cap, T, 5)); c = pts.new_zeros((cap, 3), dtype=torch.int); k = pts.new_zeros((cap,), dtype=torch.int)
            m = ext.hard_voxelize(pts, v, c, k, list(synth.NUSC_VOXEL), list(synth.NUSC_RANGE), T, cap, 3)
            return v[:m], c[:m], k[:m]
        def run_ours():
            v = pts.new_empty((cap, T, 5)); c = pts.new_empty((cap, 3), dtype=torch.int); k = pts.new_empty((cap,), dtype=torch.int)
            m = voxel.hard_voxelize(pts, v, c, k, synth.NUSC_VOXEL, synth.NUSC_RANGE, T, cap)
            return v[:m], c[:m], k[:m]
        rv, rc, rk = run_ref()
        ov, oc, ok = run_ours()
        same = bool(torch.equal(rc, oc) and torch.equal(rk, ok) and torch.equal(rv, ov))
        emit("hard_voxelize (incl. the voxel-count read)", "%d points" % n, timeit(run_ref, min(args.iters, 3), 1),
             timeit(run_ours, args.iters), "bit-exact: %s" % same)


def bench_spconv(args):
    ext = ref_build.load("sparse_conv_ext")
    if ext is None:
        return
    import synth
    from ddf_b200.ops.spconv import functional as Fsp, ops
    from ddf_b200.ops.voxel import Voxelization
    vox = Voxelization(synth.NUSC_VOXEL, synth.NUSC_RANGE, 10, (120000, 160000)).cuda().train()
    idx = []
    for b in range(2):
        _, c, _ = vox(torch.from_numpy(synth.lidar_points(260000, seed=b)).cuda())
        idx.append(torch.nn.functional.pad(c, (1, 0), value=b))
    idx = torch.cat(idx).contiguous()
    shape = [41, 1440, 1440]
    for name, subm, st, pad, cin, cout in (("subm 16->16 @stride1", True, 1, 1, 16, 16), ("down 16->32", False, 2, 1, 16, 32)):
        k3, s3, p3 = [3] * 3, [st] * 3, [pad] * 3
        out_shape = shape if subm else [(s + 2 * pad - 3) // st + 1 for s in shape]
        def ref_rb():
            return ext.get_indice_pairs_3d(idx, 2, out_shape, shape, k3, s3, p3, [1] * 3, [0] * 3, int(subm), 0)
        def our_rb():
            return ops.build_rulebook(idx, 2, shape, k3, s3, p3, 1, 0, subm, False)
        r_out, r_pairs, r_num = ref_rb()
        rb = our_rb()
        chk = "pair counts equal: %s, outputs equal: %s" % (bool(torch.equal(r_num.cpu(), rb.indice_pair_num.cpu())),
                                                           bool(torch.equal(r_out, rb.outids)) if not subm else True)
        emit("rulebook build " + name, "%d voxels" % idx.shape[0], timeit(ref_rb, args.iters), timeit(our_rb, args.iters), chk)
        n_out = r_out.shape[0]
        feat = torch.randn(idx.shape[0], cin, device="cuda")
        w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
        go = torch.randn(n_out, cout, device="cuda")
        r_y = ext.indice_conv_fp32(feat, w, r_pairs, r_num, n_out, 0, int(subm))
        f = feat.clone().requires_grad_(); wt = w.clone().requires_grad_()
        o_y = Fsp.table_conv(f, wt, None, rb, n_out)
        # the reference GPU rulebook orders strided outputs by sorted flat index like ours: rows comparable directly
        emit("conv forward " + name, "%d -> %d rows" % (idx.shape[0], n_out),
             timeit(lambda: ext.indice_conv_fp32(feat, w, r_pairs, r_num, n_out, 0, int(subm)), args.iters),
             timeit(lambda: Fsp.table_conv(feat, w, None, rb, n_out), args.iters), "rel err %.1e" % rel(o_y, r_y))
        r_gi, r_gw = ext.indice_conv_backward_fp32(feat, w, go, r_pairs, r_num, 0, int(subm))
        o_y.backward(go)
        def our_bwd():
            f2 = feat.detach().requires_grad_(); w2 = w.detach().requires_grad_()
            Fsp.table_conv(f2, w2, None, rb, n_out).backward(go)
        t_fb = timeit(our_bwd, args.iters)
        t_f = timeit(lambda: Fsp.table_conv(feat, w, None, rb, n_out), args.iters)
        emit("conv backward (dgrad + wgrad) " + name, "%d rows" % n_out,
             timeit(lambda: ext.indice_conv_backward_fp32(feat, w, go, r_pairs, r_num, 0, int(subm)), args.iters),
             max(t_fb - t_f, 1e-6), "grad_in %.1e grad_w %.1e" % (rel(f.grad, r_gi), rel(wt.grad, r_gw)))
        if not subm:
            break
    # the wide stages: rulebook at stride 4 / 8 comes from our builder (same format), reference conv on it
    cur, cur_shape = idx, shape
    for st_name, cin, cout in (("subm 64->64 @stride4", 64, 64), ("subm 128->128 @stride8", 128, 128)):
        steps = 2 if cin == 64 else 3
        cur, cur_shape = idx, shape
        for i in range(steps):
            pad = [0, 1, 1] if i == 2 else [1] * 3
            rbd = ops.build_rulebook(cur, 2, cur_shape, 3, 2, pad, 1, 0, False, False)
            cur, cur_shape = rbd.outids, rbd.out_spatial_shape
        r_out, r_pairs, r_num = ext.get_indice_pairs_3d(cur, 2, cur_shape, cur_shape, [3] * 3, [1] * 3, [1] * 3, [1] * 3,
                                                        [0] * 3, 1, 0)
        rb = ops.build_rulebook(cur, 2, cur_shape, 3, 1, 1, 1, 0, True, False)
        n = cur.shape[0]
        feat = torch.randn(n, cin, device="cuda")
        w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
        go = torch.randn(n, cout, device="cuda")
        r_y = ext.indice_conv_fp32(feat, w, r_pairs, r_num, n, 0, 1)
        f = feat.clone().requires_grad_(); wt = w.clone().requires_grad_()
        o_y = Fsp.table_conv(f, wt, None, rb, n)
        o_y.backward(go)
        r_gi, r_gw = ext.indice_conv_backward_fp32(feat, w, go, r_pairs, r_num, 0, 1)
        emit("conv forward " + st_name, "%d rows" % n,
             timeit(lambda: ext.indice_conv_fp32(feat, w, r_pairs, r_num, n, 0, 1), args.iters),
             timeit(lambda: Fsp.table_conv(feat, w, None, rb, n), args.iters), "rel err %.1e" % rel(o_y, r_y))
        def our_bwd():
            f2 = feat.detach().requires_grad_(); w2 = w.detach().requires_grad_()
            Fsp.table_conv(f2, w2, None, rb, n).backward(go)
        t_fb = timeit(our_bwd, args.iters)
        t_f = timeit(lambda: Fsp.table_conv(feat, w, None, rb, n), args.iters)
        emit("conv backward (dgrad + wgrad) " + st_name, "%d rows" % n,
             timeit(lambda: ext.indice_conv_backward_fp32(feat, w, go, r_pairs, r_num, 0, 1), args.iters),
             max(t_fb - t_f, 1e-6), "grad_in %.1e grad_w %.1e" % (rel(f.grad, r_gi), rel(wt.grad, r_gw)))


def bench_pointops(args):
    from ddf_b200.ops import pointops
    fps = ref_build.load("furthest_point_sample_ext")
    bq = ref_build.load("ball_query_ext")
    gp = ref_build.load("group_points_ext")
    Bp, N, m, ns, C = 12, 8000, 2048, 32, 128
    torch.manual_seed(0)
    xyz = (torch.rand(Bp, N, 3, device="cuda") * torch.tensor([108.0, 108.0, 8.0], device="cuda")).contiguous()
    feats = torch.randn(Bp, C, N, device="cuda")
    ours_idx = pointops.furthest_point_sample(xyz, m)
    if fps is not None:
        def run_ref():
            out = torch.zeros(Bp, m, dtype=torch.int32, device="cuda")
            temp = torch.full((Bp, N), 1e10, device="cuda")
            fps.furthest_point_sampling_wrapper(Bp, N, m, xyz, temp, out)
            return out
        r_idx = run_ref()
        emit("furthest_point_sample", "%d rows x %d pts -> %d" % (Bp, N, m), timeit(run_ref, args.iters),
             timeit(lambda: pointops.furthest_point_sample(xyz, m), args.iters),
             "indices equal: %s" % bool(torch.equal(r_idx, ours_idx)))
    centres = pointops.gather_points(xyz.transpose(1, 2).contiguous(), ours_idx).transpose(1, 2).contiguous()
    o_bq = pointops.ball_query(0.0, 2.0, ns, xyz, centres)
    if bq is not None:
        def run_ref():
            out = torch.zeros(Bp, m, ns, dtype=torch.int32, device="cuda")
            bq.ball_query_wrapper(Bp, N, m, 0.0, 2.0, ns, centres, xyz, out)
            return out
        r_bq = run_ref()
        emit("ball_query", "%d x %d centres, %d pts" % (Bp, m, N), timeit(run_ref, args.iters),
             timeit(lambda: pointops.ball_query(0.0, 2.0, ns, xyz, centres), args.iters),
             "indices equal: %s" % bool(torch.equal(r_bq, o_bq)))
    if gp is not None:
        def run_ref():
            out = torch.empty(Bp, C, m, ns, device="cuda")
            gp.forward(Bp, C, N, m, ns, feats, o_bq, out)
            return out
        r_g = run_ref()
        o_g = pointops.grouping_operation(feats, o_bq)
        emit("group_points", "%d x %d x %d x %d" % (Bp, C, m, ns), timeit(run_ref, args.iters),
             timeit(lambda: pointops.grouping_operation(feats, o_bq), args.iters), "equal: %s" % bool(torch.equal(r_g, o_g)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    for name, fn in (("msda", bench_msda), ("voxel", bench_voxel), ("spconv", bench_spconv), ("pointops", bench_pointops)):
        if a.only and name not in a.only.split(","):
            continue
        try:
            fn(a)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(op=name, error=repr(e))), flush=True)
