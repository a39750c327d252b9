"""Tile-staged dual-query MSDA (csrc/msda_tile.cu) against the reference module's own arithmetic on the CPU:
loc = ref + offsets / (W, H), weights = softmax(logits) (ops/modules/ms_deform_attn.py:149-166), then the reference's
pure-PyTorch ``ms_deform_attn_core_pytorch`` (ops/functions/ms_deform_attn_func.py:41-61; restated in
oracle/cpu_path.py and pinned by tests/golden/msda_golden.npz) with autograd for the backward.  fp32, 1e-3 relative
(north_star); measured ~1e-6 forward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference(value, ref, off, logit, H, W):
    """The sampling locations are formed in fp32 exactly as the module does (ref + off / (W, H)): which pixel a
    sample falls in - and with it the location gradient, discontinuous at pixel borders - is decided there. Everything
    downstream runs in float64."""
    from oracle import cpu_path
    value, logit = (t.detach().cpu().double().requires_grad_() for t in (value, logit))
    off = off.detach().cpu().float().requires_grad_()
    N, Lq, M = off.shape[:3]
    loc = (ref.cpu().float()[:, :, None, None, None, :] + off / torch.tensor([W, H], dtype=torch.float32)).double()
    attn = torch.softmax(logit.view(N, Lq, M, 4), -1).view(N, Lq, M, 1, 4)
    out = cpu_path.ms_deform_attn_core_pytorch(value, torch.tensor([[H, W]]), loc, attn)
    return out, (value, off, logit)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CASES = [
    # N, H, W, M, D, Lq, offset scale (pixels), ref spread
    dict(N=2, H=23, W=37, M=8, D=16, Lq=300, osc=3.0),                 # ragged tiles at the right / bottom edge
    dict(N=3, H=32, W=48, M=8, D=8, Lq=257, osc=3.0),                  # Voxel-RCNN head width
    dict(N=1, H=40, W=40, M=8, D=16, Lq=700, osc=2.0, cluster=True),   # > 128 queries in one tile: chunked
    dict(N=2, H=30, W=30, M=8, D=16, Lq=200, osc=25.0),                # offsets far beyond the halo: global fallback
    dict(N=2, H=20, W=28, M=4, D=16, Lq=5, osc=1.0),                   # tiny
]


@pytest.mark.parametrize("case", CASES)
def test_tile_msda_matches_module_arithmetic(case):
    from ddf_b200.ops import msda
    N, H, W, M, D, Lq = (case[k] for k in ("N", "H", "W", "M", "D", "Lq"))
    g = torch.Generator().manual_seed(7)
    value = torch.randn(N, H * W, M, D, generator=g)
    ref = torch.rand(N, Lq, 2, generator=g) * 1.2 - 0.1                      # some outside [0, 1]
    if case.get("cluster"):
        ref = 0.5 + 0.05 * torch.randn(N, Lq, 2, generator=g)
    ref[:, -Lq // 5:] = 0.0                                                   # padded rows: (0, 0)
    off = case["osc"] * torch.randn(N, Lq, M, 1, 4, 2, generator=g)
    logit = torch.randn(N, Lq, M, 4, generator=g)
    gout = torch.randn(N, Lq, M * D, generator=g)
    assert msda.tile_supported(M, D, 1, 4)

    plan = msda.TilePlan(ref.cuda(), H, W)
    # plan invariants: perm is a permutation grouped by tile, work items tile the query list in chunks <= 128
    buf = plan.buf.cpu().numpy()
    TX, TY = (W + 15) // 16, (H + 15) // 16
    NT, NQ = N * TX * TY, N * Lq
    n_work = int(buf[0])
    counts, tile_start = buf[1:1 + NT], buf[1 + NT:2 + 2 * NT]
    perm = buf[2 + 2 * NT + 2 * NQ:2 + 2 * NT + 3 * NQ]
    work = buf[2 + 2 * NT + 3 * NQ:][:3 * n_work].reshape(-1, 3)
    assert counts.sum() == NQ and tile_start[-1] == NQ and np.array_equal(np.sort(perm), np.arange(NQ))
    assert work[:, 2].sum() == NQ and work[:, 2].max() <= 128 and work[:, 2].min() >= 1
    px = np.clip(np.floor(ref.numpy()[..., 0] * np.float32(W)), 0, W - 1).astype(int) // 16
    py = np.clip(np.floor(ref.numpy()[..., 1] * np.float32(H)), 0, H - 1).astype(int) // 16
    tile = ((np.arange(N)[:, None] * TY + py) * TX + px).reshape(-1)
    for t, q0, n in work:
        assert np.all(tile[perm[q0:q0 + n]] == t)

    v, o, l = (t.cuda().requires_grad_() for t in (value, off, logit))
    out = msda.MSDeformAttnTileFunction.apply(v, o, l, plan)
    out.backward(gout.cuda())
    r_out, (rv, ro, rl) = reference(value, ref, off, logit, H, W)
    r_out.backward(gout.double())
    assert rel(out, r_out) < 1e-5
    assert rel(v.grad, rv.grad) < 1e-4
    assert rel(o.grad, ro.grad) < 1e-4
    assert rel(l.grad, rl.grad) < 1e-4

    # the generic op (reference signature) on the same problem gives the same answer
    norm = torch.tensor([W, H], dtype=torch.float32)
    loc = (ref[:, :, None, None, None, :] + off / norm).cuda().contiguous()
    attn = torch.softmax(logit, -1).view(N, Lq, M, 1, 4).cuda().contiguous()
    out2 = msda.MSDeformAttnFunction.apply(value.cuda(), torch.tensor([[H, W]]).cuda(), torch.tensor([0]).cuda(),
                                           loc, attn, 64)
    assert rel(out, out2) < 1e-5


def test_tile_msda_module_uses_plan_and_matches_generic_path():
    """MSDeformAttn with a TilePlan == the same module through the reference-signature op."""
    from ddf_b200.fusion.ms_deform_attn import MSDeformAttn
    from ddf_b200.ops import msda
    torch.manual_seed(0)
    attn = MSDeformAttn(d_model=128, q_model=128, n_levels=1, n_heads=8, n_points=4, q_method="sum",
                        q_rep_place=["weight"]).cuda()
    torch.nn.init.normal_(attn.sampling_offsets.weight, std=0.05)
    torch.nn.init.normal_(attn.attention_weights.weight, std=0.1)
    N, Lq, H, W = 3, 400, 28, 50
    q, iq = torch.randn(N, Lq, 128, device="cuda"), torch.randn(N, Lq, 128, device="cuda")
    src = torch.randn(N, H * W, 128, device="cuda", requires_grad=True)
    ref = torch.rand(N, Lq, 1, 2, device="cuda")
    shapes, lsi = torch.tensor([[H, W]], device="cuda"), torch.tensor([0], device="cuda")
    a = attn(q, ref, src, shapes, lsi, None, i_query=iq, plan=msda.TilePlan(ref, H, W))
    ga, = torch.autograd.grad(a.square().sum(), src)
    b = attn(q, ref, src, shapes, lsi, None, i_query=iq)
    gb, = torch.autograd.grad(b.square().sum(), src)
    assert rel(a, b) < 1e-5 and rel(ga, gb) < 1e-4


def test_tile_msda_cp_async_staging_variant():
    """DDF_MSDA_STAGE=cp: the window is staged by LDGSTS instead of the 4-D TMA box load (selected once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DDF_MSDA_STAGE="cp")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.join(root, "tests", "test_msda_tile_gpu.py"),
                        "-k", "matches_module_arithmetic", "-m", "gpu"], env=env, capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_compact_encoder_equals_padded_encoder_on_real_rows():
    """The encoder with ``valid_index`` (layers visit only the real queries; ragged MSDA plan) returns, on those rows,
    what the padded pass returns - forward and parameter gradients (loss over the real rows only, as the wrappers use it)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    import detfill
    from ddf_b200.fusion import actr
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = dict(num_channels=[64], query_num_feat=128, num_enc_layers=2, max_num_ne_voxel=26000, pos_encode_method="depth",
               feature_modal="hybrid", hybrid_cfg=dict(attn_layer="BiGateSum1D_2", q_method="sum", q_rep_place=["weight"]))
    net = actr.build(cfg).cuda().train()
    detfill.fill_state_dict(net)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    Bp, Lq, H, W = 6, 300, 28, 50
    g = torch.Generator().manual_seed(3)
    counts = [300, 120, 0, 77, 250, 1]
    valid = torch.cat([b * Lq + torch.arange(c) for b, c in enumerate(counts)]).cuda()
    mask = torch.zeros(Bp * Lq, 1, device="cuda")
    mask[valid] = 1
    pad = lambda t: (t.reshape(Bp * Lq, -1) * mask).reshape(t.shape)
    v_feat = pad(torch.randn(Bp, Lq, 128, generator=g).cuda())
    grid = pad(torch.rand(Bp, Lq, 2, generator=g).cuda())
    v_i = pad(torch.randn(Bp, Lq, 64, generator=g).cuda())
    lidar = pad((torch.rand(Bp, Lq, 3, generator=g) * 50).cuda())
    img = torch.randn(Bp, 64, H, W, generator=g).cuda()
    outs, grads = [], []
    for vi in (None, valid):
        net.zero_grad()
        out = net(v_feat, grid, [img], v_i, lidar, valid_index=vi)
        sel = out.reshape(Bp * Lq, -1)[valid]
        sel.square().sum().backward()
        outs.append(sel.detach())
        grads.append({n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
    assert rel(outs[1], outs[0]) < 1e-5
    assert grads[0].keys() == grads[1].keys()
    for n in grads[0]:
        assert rel(grads[1][n], grads[0][n]) < 1e-4, n
    # padded rows of the compact result are zero
    full = net(v_feat, grid, [img], v_i, lidar, valid_index=valid).reshape(Bp * Lq, -1)
    assert float(full[mask[:, 0] == 0].abs().max()) == 0.0
