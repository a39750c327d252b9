"""GPU-vs-GPU parity against the REFERENCE's own CUDA kernels, rebuilt for sm_100 from the reference sources by
oracle/ref_build.py (oracle/_ref/*.so: sparse_conv_ext and voxel_layer unmodified; MultiScaleDeformableAttention and
the point ops with the documented two-line torch-2 fixes). These have no CPU path, so they complement the CPU oracle:
integer / index results must be identical, floating point within 1e-3 relative (north_star)."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def ref(name):
    from oracle import ref_build
    ext = ref_build.load(name)
    if ext is None:
        pytest.skip("oracle/_ref/%s.so was not built (needs /root/reference at build time)" % name)
    return ext


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def voxel_centres(n_points, seed, stride=8):
    """stride-8 voxel centres of a synthetic sweep, the kind of cloud the LocalTransformer samples from."""
    pts = synth.lidar_points(n_points, seed=seed)[:, :3]
    vs = np.array(synth.NUSC_VOXEL, np.float32) * stride
    cells = np.unique(np.floor((pts - np.array(synth.NUSC_RANGE[:3], np.float32)) / vs).astype(np.int32), axis=0)
    return (cells.astype(np.float32) + 0.5) * vs + np.array(synth.NUSC_RANGE[:3], np.float32)


@pytest.mark.parametrize("n,m", [(3000, 512), (8000, 2048), (20000, 2048)])
def test_fps_indices_identical_to_reference_cuda_kernel(n, m):
    from ddf_b200.ops import pointops
    ext = ref("furthest_point_sample_ext")
    rows = []
    for b in range(3):
        c = voxel_centres(120000, seed=b)
        c = c[np.random.default_rng(b).permutation(len(c))][:n]
        row = np.zeros((n, 3), np.float32)          # zero-padded tail, as the fusion wrappers produce it
        row[:len(c) - 200 * b] = c[:len(c) - 200 * b]
        rows.append(row)
    xyz = torch.from_numpy(np.stack(rows)).cuda()
    out = torch.zeros(3, m, dtype=torch.int32, device="cuda")
    temp = torch.full((3, n), 1e10, device="cuda")
    ext.furthest_point_sampling_wrapper(3, n, m, xyz, temp, out)
    ours = pointops.furthest_point_sample(xyz, m)
    assert torch.equal(ours, out)


def test_fps_tiny_tie_case():
    """The tie case of tests/test_oracle_pointops.py on the compiled reference kernel: pins the oracle's tie rule."""
    from ddf_b200.ops import pointops
    from oracle import pointops as opo
    ext = ref("furthest_point_sample_ext")
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, 5] = [1, 0, 0]
    xyz[0, 6] = [1, 0, 0]
    t = torch.from_numpy(xyz).cuda()
    out = torch.zeros(1, 3, dtype=torch.int32, device="cuda")
    ext.furthest_point_sampling_wrapper(1, 8, 3, t, torch.full((1, 8), 1e10, device="cuda"), out)
    assert out.cpu().tolist() == opo.furthest_point_sample(xyz, 3).tolist() == pointops.furthest_point_sample(t, 3).cpu().tolist()


def test_ball_query_and_group_identical_to_reference_cuda_kernels():
    from ddf_b200.ops import pointops
    bq, gp, ga = ref("ball_query_ext"), ref("group_points_ext"), ref("gather_points_ext")
    B, n, m, ns, C = 2, 6000, 1024, 32, 64
    xyz = torch.from_numpy(np.stack([voxel_centres(100000, seed=5 + b)[:n] for b in range(B)])).cuda()
    idx = pointops.furthest_point_sample(xyz, m)
    centres = pointops.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    r_c = torch.empty(B, 3, m, device="cuda")
    ga.gather_points_wrapper(B, 3, n, m, xyz.transpose(1, 2).contiguous(), idx, r_c)
    assert torch.equal(r_c.transpose(1, 2), centres)
    r_idx = torch.zeros(B, m, ns, dtype=torch.int32, device="cuda")
    bq.ball_query_wrapper(B, n, m, 0.0, 2.0, ns, centres, xyz, r_idx)
    o_idx = pointops.ball_query(0.0, 2.0, ns, xyz, centres)
    assert torch.equal(o_idx, r_idx)
    feats = torch.randn(B, C, n, device="cuda")
    r_g = torch.empty(B, C, m, ns, device="cuda")
    gp.forward(B, C, n, m, ns, feats, r_idx, r_g)
    assert torch.equal(pointops.grouping_operation(feats, o_idx), r_g)


def test_msda_matches_reference_cuda_kernels():
    from ddf_b200.ops import msda
    ext = ref("MultiScaleDeformableAttention")
    N, H, W, M, D, Lq = 3, 40, 66, 8, 16, 900
    torch.manual_seed(1)
    value = torch.randn(N, H * W, M, D, device="cuda")
    shapes = torch.tensor([[H, W]], device="cuda")
    lsi = torch.zeros(1, dtype=torch.long, device="cuda")
    refp = torch.rand(N, Lq, 2, device="cuda") * 1.1 - 0.05
    off = torch.randn(N, Lq, M, 1, 4, 2, device="cuda") * 3
    logit = torch.randn(N, Lq, M, 4, device="cuda")
    norm = torch.tensor([W, H], device="cuda", dtype=torch.float32)
    loc = (refp[:, :, None, None, None, :] + off / norm).contiguous()
    attn = torch.softmax(logit, -1).view(N, Lq, M, 1, 4).contiguous()
    gout = torch.randn(N, Lq, M * D, device="cuda")
    r_out = ext.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
    r_gv, r_gl, r_ga = ext.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
    # reference-signature op
    assert rel(msda.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), r_out) < 1e-5
    gv, gl, ga = msda.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
    assert rel(gv, r_gv) < 1e-4 and rel(gl, r_gl) < 1e-4 and rel(ga, r_ga) < 1e-4
    # tile-staged dual-query op (grad wrt raw offsets = grad_loc / (W, H))
    plan = msda.TilePlan(refp, H, W)
    assert rel(msda.msda_tile_forward(value, plan, off, logit), r_out) < 1e-5
    tv, to, tl = msda.msda_tile_backward(value, plan, off, logit, gout)
    assert rel(tv, r_gv) < 1e-4 and rel(to, r_gl / norm) < 1e-4


def test_voxelization_identical_to_reference_cuda_kernel():
    from ddf_b200.ops import voxel
    ext = ref("voxel_layer")
    pts = torch.from_numpy(synth.lidar_points(30000, seed=3)).cuda()
    T, cap = 10, 20000     # cap below the voxel count: the max_voxels cut-off is part of the contract
    rv = pts.new_zeros((cap, T, 5)); rc = pts.new_zeros((cap, 3), dtype=torch.int); rk = pts.new_zeros((cap,), dtype=torch.int)
    m = ext.hard_voxelize(pts, rv, rc, rk, list(synth.NUSC_VOXEL), list(synth.NUSC_RANGE), T, cap, 3)
    ov = pts.new_empty((cap, T, 5)); oc = pts.new_empty((cap, 3), dtype=torch.int); ok = pts.new_empty((cap,), dtype=torch.int)
    m2 = voxel.hard_voxelize(pts, ov, oc, ok, synth.NUSC_VOXEL, synth.NUSC_RANGE, T, cap)
    assert m == m2 == cap
    assert torch.equal(rc[:m], oc[:m]) and torch.equal(rk[:m], ok[:m]) and torch.equal(rv[:m], ov[:m])


def test_sparse_conv_matches_reference_cuda_path():
    """Rulebook (GPU order of the reference: outputs sorted by flat index) and conv forward / backward."""
    from ddf_b200.ops.spconv import functional as Fsp, ops
    from test_oracle_spconv import random_voxels
    ext = ref("sparse_conv_ext")
    shape = [11, 60, 60]
    idx = torch.from_numpy(random_voxels(4000, 2, shape, seed=2)).cuda()
    for subm, st, pad, cin, cout in ((True, 1, 1, 64, 64), (False, 2, 1, 32, 64)):
        out_shape = shape if subm else [(s + 2 * pad - 3) // st + 1 for s in shape]
        r_out, r_pairs, r_num = ext.get_indice_pairs_3d(idx, 2, out_shape, shape, [3] * 3, [st] * 3, [pad] * 3, [1] * 3,
                                                        [0] * 3, int(subm), 0)
        rb = ops.build_rulebook(idx, 2, shape, 3, st, pad, 1, 0, subm, False)
        assert torch.equal(r_num.cpu(), rb.indice_pair_num.cpu())
        if not subm:
            assert torch.equal(r_out, rb.outids)
        n_out = r_out.shape[0]
        feat = torch.randn(idx.shape[0], cin, device="cuda")
        w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
        go = torch.randn(n_out, cout, device="cuda")
        r_y = ext.indice_conv_fp32(feat, w, r_pairs, r_num, n_out, 0, int(subm))
        r_gi, r_gw = ext.indice_conv_backward_fp32(feat, w, go, r_pairs, r_num, 0, int(subm))
        f, wt = feat.clone().requires_grad_(), w.clone().requires_grad_()
        y = Fsp.table_conv(f, wt, None, rb, n_out)
        y.backward(go)
        assert rel(y, r_y) < 1e-4 and rel(f.grad, r_gi) < 1e-4 and rel(wt.grad, r_gw) < 1e-3
