"""End-to-end parity of the assembled hot path (voxelize -> VFE -> SparseEncoderFusion with the
3D-DF fusion hook -> dense BEV) on the GPU against the SAME module graph driven by the reference's
own CPU code (oracle/cpu_path.py: reference voxelization / spconv extensions from oracle/_ref when
present, else the C restatements; pure-PyTorch MSDA).

Tolerances (max |a - b| / max |b|), written per mode:
  * default mode = what bench.py times (``ddf_set_tensor_cores(4)``: forward / dgrad of the wide convs as bf16x3 on
    tcgen05 - 16-bit significand per product, fp32 accumulation -, narrow convs fp32, wgrad tf32; library GEMMs of the
    fusion encoder under ``allow_tf32 = True`` exactly as bench.py sets it): BEV output within 1e-3 (north_star bar;
    measured 1.7e-4 train / 4.9e-4 eval on B200), whole-model gradient within 5e-3 relative L2 (2.1e-3 measured),
    every parameter gradient within 5e-2 of its max (3.3e-2: tf32 wgrad of small-gradient layers), cosine > 0.999.
  * fp32 mode (``ddf_set_tensor_cores(0)``, library GEMMs in fp32): 1e-3 outputs, 1e-2 every parameter gradient.
  * single-pass tf32 convs (mode 1, not the default any more): 1.3e-3 at the BEV output in train mode - above the bar,
    which is why bf16x3 is the default; kept as a looser regression check (3e-3).
  * one full-size frame (260k points, the bench workload's shape) against the CPU oracle in the default mode."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def build(seed=0):
    import configs
    import ddf_b200.fusion.point_fusion  # noqa: F401
    import ddf_b200.fusion.sparse_encoder  # noqa: F401
    import ddf_b200.fusion.voxel_encoder  # noqa: F401
    from ddf_b200.fusion.detector import TransFusionPtsBranch
    torch.manual_seed(seed)
    m = TransFusionPtsBranch(**configs.transfusion_f())
    # non-trivial attention: the reference init zeroes the offset / weight matrices
    for mod in m.modules():
        if mod.__class__.__name__ == "MSDeformAttn":
            torch.nn.init.normal_(mod.sampling_offsets.weight, std=0.02)
            torch.nn.init.normal_(mod.attention_weights.weight, std=0.05)
    return m


def inputs(batch, n_points):
    pts = [torch.from_numpy(synth.lidar_points(n_points, seed=40 + b)) for b in range(batch)]
    feats = torch.from_numpy(synth.camera_features(batch, 6, (112, 200), seed=3))
    metas = [synth.nusc_img_meta(6) for _ in range(batch)]
    return pts, feats, metas


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _bench_arithmetic(on=True):
    # exactly what bench.py:run_ours sets for the library GEMMs
    torch.backends.cuda.matmul.allow_tf32 = on
    torch.backends.cudnn.allow_tf32 = on


def _forward_eval(n_points=8000, batch=2, tol=1e-3):
    from oracle import cpu_path
    m_cpu = build().eval()
    m_gpu = copy.deepcopy(m_cpu).cuda().eval()
    pts, feats, metas = inputs(batch, n_points)
    with torch.no_grad():
        with cpu_path.reference_cpu_ops():
            ref = m_cpu(pts, [feats], metas)
        out = m_gpu([p.cuda() for p in pts], [feats.cuda()], metas).cpu()
    assert out.shape == ref.shape == (batch, 256, 180, 180)
    # same active BEV cells, up to ReLU outputs that sit at +-0 within rounding
    differ = (out != 0) != (ref != 0)
    assert float(torch.maximum(out.abs(), ref.abs())[differ].max() if differ.any() else 0.0) < tol * float(ref.abs().max())
    assert rel(out, ref) < tol


def test_forward_eval_matches_reference_cpu_path():
    """Default conv mode + the benchmark's library-GEMM arithmetic."""
    _bench_arithmetic(True)
    try:
        _forward_eval()
    finally:
        _bench_arithmetic(False)


def test_full_size_frame_matches_reference_cpu_path():
    """One frame of the bench workload's size (260k points, 6 x 256 x 112 x 200 camera features) against the CPU
    oracle, in the benchmarked arithmetic (about 20 s of host time)."""
    _bench_arithmetic(True)
    try:
        _forward_eval(n_points=260000, batch=1)
    finally:
        _bench_arithmetic(False)


@pytest.fixture
def fp32_convs():
    from ddf_b200 import lib
    prev = lib.get_lib().ddf_set_tensor_cores(0)
    yield
    lib.get_lib().ddf_set_tensor_cores(prev)


@pytest.fixture
def tf32_convs():
    from ddf_b200 import lib
    prev = lib.get_lib().ddf_set_tensor_cores(1)
    yield
    lib.get_lib().ddf_set_tensor_cores(prev)


def test_forward_eval_fp32_mode_matches_reference_cpu_path(fp32_convs):
    _bench_arithmetic(False)
    _forward_eval()


def test_train_step_fp32_mode_gradients_match_reference_cpu_path(fp32_convs):
    _train_step_parity(out_tol=1e-3, grad_max_tol=1e-2, grad_l2_tol=1e-2, min_cos=0.9999)


def test_train_step_gradients_match_reference_cpu_path():
    """The benchmarked arithmetic: bf16x3 convs (default mode) + tf32 library GEMMs."""
    _train_step_parity(out_tol=1e-3, grad_max_tol=5e-2, grad_l2_tol=5e-3, min_cos=0.999, bench_gemms=True)


def test_train_step_single_pass_tf32_convs(tf32_convs):
    # tf32 products through 21 convs and their batch-statistics BatchNorm backward compound to 1.3e-3 at the output
    _train_step_parity(out_tol=3e-3, grad_max_tol=None, grad_l2_tol=0.05, min_cos=0.9)


def _train_step_parity(out_tol, grad_max_tol, grad_l2_tol, min_cos, bench_gemms=False):
    from oracle import cpu_path
    _bench_arithmetic(bench_gemms)   # the 1x1 Conv2d input_proj runs through cuDNN, the Linears through cuBLAS
    try:
        _train_step_parity_body(out_tol, grad_max_tol, grad_l2_tol, min_cos)
    finally:
        _bench_arithmetic(False)


def _train_step_parity_body(out_tol, grad_max_tol, grad_l2_tol, min_cos):
    from oracle import cpu_path
    m_cpu = build(seed=1).train()
    for mod in m_cpu.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0  # dropout draws differ between devices; everything else is deterministic
    m_gpu = copy.deepcopy(m_cpu).cuda().train()
    pts, feats, metas = inputs(1, 6000)
    with cpu_path.reference_cpu_ops():
        ref = m_cpu(pts, [feats], metas)
        ref.square().mean().backward()
    out = m_gpu([p.cuda() for p in pts], [feats.cuda()], metas)
    out.square().mean().backward()
    assert rel(out.detach().cpu(), ref.detach()) < out_tol
    g_cpu = dict(m_cpu.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in m_cpu.parameters() if p.grad is not None)
    checked, bad, num, den = 0, [], 0.0, 0.0
    for name, p in m_gpu.named_parameters():
        gc = g_cpu[name].grad
        if gc is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        gd, gc = p.grad.cpu().double(), gc.double()
        num += float((gd - gc).square().sum())
        den += float(gc.square().sum())
        checked += 1
        if grad_max_tol is not None:
            # gradients below 1e-3 of the largest one (cancellation noise, e.g. the gate biases) are
            # compared on that absolute scale
            r_max = float((gd - gc).abs().max() / max(float(gc.abs().max()), 1e-3 * gmax))
            if not r_max < grad_max_tol:
                bad.append((name, r_max, float(gc.abs().max())))
        if float(gc.norm()) > 1e-3 * gmax:
            cos = float((gd * gc).sum() / (gd.norm() * gc.norm()).clamp_min(1e-300))
            if not cos > min_cos:
                bad.append((name, cos, float(gc.abs().max())))
    assert not bad, sorted(bad, key=lambda t: -t[1])[:8]
    assert (num / den) ** 0.5 < grad_l2_tol, (num / den) ** 0.5   # whole-model gradient, relative L2
    assert checked > 100
    # parameters that can never get a gradient
    from ddf_b200.fusion import structurally_unused_parameters
    for n in structurally_unused_parameters(m_gpu):
        assert dict(m_gpu.named_parameters())[n].grad is None
