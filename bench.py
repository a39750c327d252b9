"""Headline benchmark: TransFusion-L + 3D-DF hot path, forward + backward + optimizer step, samples/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic nuScenes-shaped input per rank
(BASELINE.json configs[2]: TransFusion-L+3D-DF, bs=2/GPU): hard voxelization -> HardSimpleVFE ->
SparseEncoderFusion (21 sparse convs + BN/ReLU, 3D-DF fusion hook: GPU projection, per-camera
split, ACTR encoder with dual-query MSDA, FFNs, bi-gate) -> dense BEV map; loss = mean(out^2);
backward; grad-clip 0.1; AdamW step (the reference's optimizer_config / optimizer,
TransFusion/configs/transfusion_nusc_voxel_F.py:302-303).  Camera backbone, BEV backbone and head
are outside the path: camera features are synthetic N(0,1) maps.

Multi-GPU: pure data parallel (DDP over NCCL, gradients only), weak scaling, launched by
torch.distributed.run. `--impl reference` times the reference's own CPU implementation of the
same path (oracle/_ref extensions + PyTorch CPU) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "TransFusion-L+3D-DF hot path fwd+bwd samples/sec"
UNIT = "samples/s"
BATCH_PER_GPU = 2
POINTS_PER_SAMPLE = 260000   # nuScenes 10-sweep cloud
N_CAM = 6
FEAT_HW = (112, 200)         # FPN level 0 of a 448x800 input


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32-gemm", action="store_true",
                    help="keep the library GEMMs (nn.Linear / 1x1 conv) in full fp32 instead of tf32")
    return ap.parse_args()


def workload_config(n_gpus):
    return {
        "workload": "TransFusion-L+3D-DF (transfusion_nusc_voxel_F) hot path: voxelize+VFE+SparseEncoderFusion"
                    "(ACTR hybrid, 2 enc layers)+dense BEV, fwd+bwd+clip+AdamW",
        "per_gpu_batch": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus,
        "points_per_sample": POINTS_PER_SAMPLE, "cams": N_CAM, "cam_feat": [256, *FEAT_HW],
        "sparse_shape": [41, 1440, 1440], "parallelism": "dp%d" % n_gpus,
        "l2": "inputs larger than L2 (285 MB of points + camera features per step)",
    }


def build_model(device):
    import configs
    import ddf_b200.fusion.point_fusion  # noqa: F401  (registers FUSION_LAYERS['ACTR'])
    import ddf_b200.fusion.sparse_encoder  # noqa: F401
    import ddf_b200.fusion.voxel_encoder  # noqa: F401
    from ddf_b200.fusion import structurally_unused_parameters
    from ddf_b200.fusion.detector import TransFusionPtsBranch
    torch.manual_seed(0)
    model = TransFusionPtsBranch(**configs.transfusion_f()).to(device).train()
    frozen = set(structurally_unused_parameters(model))
    for n, p in model.named_parameters():
        if n in frozen:
            p.requires_grad_(False)
    return model


def host_batch(rank, batch):
    import synth
    pts = [torch.from_numpy(synth.lidar_points(POINTS_PER_SAMPLE, seed=1000 * rank + b)) for b in range(batch)]
    feats = torch.from_numpy(synth.camera_features(batch, N_CAM, FEAT_HW, seed=rank))
    metas = [synth.nusc_img_meta(N_CAM) for _ in range(batch)]
    return pts, feats, metas


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons every 100 ms while the timed region runs. NVML in-process
    (a forked nvidia-smi every 200 ms steals the launch thread's core); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n, h = self.nvml, self.handle
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            return [float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                    n.nvmlDeviceGetPowerUsage(h) / 1e3] + [bool(mask & b) for _, b in self.BITS]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        r = [x.strip() for x in out.strip().split(",")]
        return [float(r[0]), float(r[1]), float(r[2])] + [v.lower().startswith("active") for v in r[3:7]]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            self.stop_flag.wait(0.1 if self.nvml is not None else 0.5)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted({name for r in self.rows for (name, _), v in zip(self.BITS, r[3:7]) if v})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(r[2] for r in self.rows) if self.rows else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_reference_run(steps, warmup, batch=1):
    """The reference's own CPU implementation of the path (oracle/cpu_path.py) on the host cores."""
    from oracle import cpu_path
    torch.set_num_threads(os.cpu_count())
    model = build_model("cpu")
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01)
    pts, feats, metas = host_batch(0, batch)
    times = []
    with cpu_path.reference_cpu_ops():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = model(pts, [feats], metas)
            loss = out.square().mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            opt.step()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    t = float(np.mean(times))
    return {"value": batch / t, "unit": UNIT, "cores": os.cpu_count(), "kind": cpu_path.kind(),
            "sample": "%d frame(s) x %d points + %d cam maps, fwd+bwd+clip+AdamW, %d timed step(s) after %d warm-up; "
                      "%.1f s/step" % (batch, POINTS_PER_SAMPLE, N_CAM, steps, warmup, t)}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    base, t = cpu_reference_run(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["config"]["reference_sample"] = "1 frame per step on the host cores (bounded sample of the bs=2/GPU workload)"
    print(json.dumps(line))


class KernelTimer(object):
    """Per-launch CUDA-event timing of this library's kernel families on the launching stream (one
    extra step after the timed region), with their algorithmic FLOPs / bytes."""

    def __init__(self):
        self.records = {"conv": [], "wgrad": [], "msda_fwd": [], "msda_bwd": []}
        self.saved = []

    def _wrap(self, mod, name, kind, meta):
        orig = getattr(mod, name)
        timer = self

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(*a, **k)
            e1.record()
            timer.records[kind].append((e0, e1, meta(*a, **k)))
            return out

        self.saved.append((mod, name, orig))
        setattr(mod, name, wrapped)

    def install(self):
        from ddf_b200.ops import msda
        from ddf_b200.ops.spconv import ops

        def conv_meta(table, n_src, n_dst, cin, cout):
            return ("table", table, n_src, n_dst, cin, cout)

        self._wrap(ops, "sparse_conv_forward", "conv",
                   lambda f, w, t, b, n_out, *fmt: conv_meta(t, f.shape[0], n_out, w.shape[-2], w.shape[-1]))
        # dgrad timing includes the small filter-rounding launch that precedes the conv kernel
        self._wrap(ops, "sparse_conv_dgrad", "conv",
                   lambda w, g, t, n_in, *fmt: conv_meta(t, g.shape[0], n_in, w.shape[-1], w.shape[-2]))
        self._wrap(ops, "sparse_conv_wgrad", "wgrad",
                   lambda f, w, g, pairs, num: ("pairs", num, f.shape[0], g.shape[0], w.shape[-2], w.shape[-1]))
        self._wrap(ops, "sparse_conv_wgrad_table", "wgrad",
                   lambda f, w, g, t: conv_meta(t, f.shape[0], g.shape[0], w.shape[-2], w.shape[-1]))

        def msda_meta(value, shapes, lsi, loc, *rest):
            N, S, M, D = value.shape
            Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
            return (N, S, M, D, Lq, L, P)

        self._wrap(msda, "ms_deform_attn_forward", "msda_fwd", msda_meta)
        self._wrap(msda, "ms_deform_attn_backward", "msda_bwd", msda_meta)

        def tile_meta(value, plan, *rest):
            N, S, M, D = value.shape
            return (N, S, M, D, plan.Lq, 1, 4)

        # the tile-staged dual-query kernels (what the encoder runs): same algorithmic bytes, the softmax / location
        # arithmetic of the module is inside the kernel
        self._wrap(msda, "msda_tile_forward", "msda_fwd", tile_meta)
        self._wrap(msda, "msda_tile_backward", "msda_bwd", tile_meta)

    def remove(self):
        for mod, name, orig in self.saved:
            setattr(mod, name, orig)

    def conv_summary(self, kind, by_width=None):
        ms = flops = byts = 0.0
        for a, b, (how, t, n_src, n_dst, cin, cout) in self.records[kind]:
            pairs = int((t >= 0).sum().item()) if how == "table" else int(t.sum().item())
            kvol = t.shape[1] if how == "table" else t.shape[0]
            dt, fl = a.elapsed_time(b), 2.0 * pairs * cin * cout
            ms += dt
            flops += fl
            byts += 4.0 * (n_src * cin + n_dst * cout + kvol * cin * cout) + 8.0 * pairs
            if by_width is not None:
                w = by_width.setdefault("%d->%d" % (cin, cout), [0, 0.0, 0.0])
                w[0] += 1
                w[1] += dt
                w[2] += fl
        return len(self.records[kind]), ms, flops, byts

    def msda_summary(self, kind):
        ms = byts = 0.0
        for a, b, (N, S, M, D, Lq, L, P) in self.records[kind]:
            ms += a.elapsed_time(b)
            if kind == "msda_fwd":
                byts += 4.0 * (N * S * M * D + 3 * N * Lq * M * L * P + N * Lq * M * D)
            else:
                byts += 4.0 * (2 * N * S * M * D + 2 * N * Lq * M * D + 6 * N * Lq * M * L * P)
        return len(self.records[kind]), ms, byts


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: everything native libraries write to file descriptor 1 meanwhile
    # (NCCL prints its version banner there at NCCL_DEBUG=VERSION, which this image sets) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ddf_b200 import lib
    L = lib.get_lib()
    # the sparse convs already run tf32 tensor-core inputs with fp32 accumulation; give the library
    # GEMMs of the fusion encoder (value_proj, FFNs, 1x1 input_proj) the same arithmetic
    torch.backends.cuda.matmul.allow_tf32 = not args.fp32_gemm
    torch.backends.cudnn.allow_tf32 = not args.fp32_gemm
    model = build_model(dev)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    # same optimizer as the reference config (AdamW lr 1e-4 wd 0.01), PyTorch's single-kernel variant
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01, fused=True)

    h_pts, h_feats, metas = host_batch(rank, BATCH_PER_GPU)
    h_pts = [p.pin_memory() for p in h_pts]
    h_feats = h_feats.pin_memory()
    h_loss = torch.zeros(1).pin_memory()
    d_pts = [p.to(dev) for p in h_pts]
    d_feats = h_feats.to(dev)

    def step(pts, feats):
        out = net(pts, [feats], metas)
        loss = out.square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(d_pts, d_feats)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (1) device-resident inputs
    L.ddf_launch_count(1)
    ms_dev = timed(lambda: step(d_pts, d_feats), args.steps)
    launches = int(L.ddf_launch_count(1))

    # (2) end to end through the public module call: every step copies ITS inputs from pinned host
    # memory and reads the loss back. Like a data loader with pinned memory, the copy of step i+1 is
    # enqueued on a side stream while step i computes; each step waits for its own copy.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            pts = [p.to(dev, non_blocking=True) for p in h_pts]
            feats = h_feats.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["next"] = (pts, feats, ev)

    def e2e_step():
        pts, feats, ev = staged.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        for t in pts + [feats]:
            t.record_stream(torch.cuda.current_stream())
        stage_inputs()                      # next step's host->device copy overlaps this step
        loss = step(pts, feats)
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    stage_inputs()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    staged.clear()
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()

    # (3) kernel-family rooflines, timed live with CUDA events on the launching stream
    ct = KernelTimer()
    ct.install()
    step(d_pts, d_feats)
    torch.cuda.synchronize()
    by_width = {}
    n_launch, conv_ms, conv_flops, conv_bytes = ct.conv_summary("conv", by_width)
    n_wg, wg_ms, wg_flops, _ = ct.conv_summary("wgrad")
    n_mf, mf_ms, mf_bytes = ct.msda_summary("msda_fwd")
    n_mb, mb_ms, mb_bytes = ct.msda_summary("msda_bwd")
    ct.remove()
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks, how = measured_peaks()
        global_batch = BATCH_PER_GPU * world
        value = global_batch * args.steps / (ms_dev / 1e3)
        e2e = global_batch * args.steps / (ms_e2e / 1e3)
        h2d = sum(p.numel() * 4 for p in h_pts) + h_feats.numel() * 4
        tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        traffic = None   # per-launch DRAM bytes of the conv family from the committed ncu capture of this workload
        tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 storage/accumulate; sparse convs bf16x3 (hi/lo split operands, 16-bit significand products) on "
                     "tcgen05, wgrad tf32; library GEMMs " + ("f32" if args.fp32_gemm else "tf32"),
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {
                "kernel": "sparse-conv implicit GEMM, tcgen05 (forward + dgrad launches of one step)",
                "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": tf / peak_tf if peak_tf else None, "traffic": traffic,
                "peak_source": "%s bf16 dense sustained (MEASURED_PEAKS.json); kernel computes in tf32/fp32" % how,
                "launches_per_step": n_launch, "avg_launch_ms": conv_ms / max(n_launch, 1),
                "share_of_step": conv_ms / (ms_dev / args.steps),
                "algorithmic_GFLOP_per_step": conv_flops / 1e9, "algorithmic_MB_per_step": conv_bytes / 1e6,
                # contraction->output width of the launch (dgrad swaps them): launches, ms, TFLOP/s
                "by_layer_width": {k: {"launches": v[0], "ms": v[1], "TFLOPs": v[2] / (v[1] / 1e3) / 1e12 if v[1] else None}
                                   for k, v in sorted(by_width.items())},
            },
            # the other kernel families of the step, same method (achieved = algorithmic work / event time)
            "kernels": {
                "sparse_conv_wgrad": {"bound": "tensor", "launches_per_step": n_wg, "ms_per_step": wg_ms,
                                      "achieved_TFLOPs": wg_flops / (wg_ms / 1e3) / 1e12 if wg_ms else None,
                                      "frac": wg_flops / (wg_ms / 1e3) / 1e12 / peak_tf if wg_ms else None},
                "deform_attn_fwd": {"bound": "hbm", "launches_per_step": n_mf, "ms_per_step": mf_ms,
                                    "achieved_GBs": mf_bytes / (mf_ms / 1e3) / 1e9 if mf_ms else None,
                                    "peak_GBs": peaks.get("hbm_gbs"),
                                    "frac": mf_bytes / (mf_ms / 1e3) / 1e9 / peaks["hbm_gbs"] if mf_ms else None},
                "deform_attn_bwd": {"bound": "hbm", "launches_per_step": n_mb, "ms_per_step": mb_ms,
                                    "achieved_GBs": mb_bytes / (mb_ms / 1e3) / 1e9 if mb_ms else None,
                                    "peak_GBs": peaks.get("hbm_gbs"),
                                    "frac": mb_bytes / (mb_ms / 1e3) / 1e9 / peaks["hbm_gbs"] if mb_ms else None},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_reference_run(1, 0)
            except Exception as e:  # the oracle is a checker; its absence must not fake a number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
