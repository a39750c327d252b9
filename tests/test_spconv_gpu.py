"""GPU parity for the sparse-conv path through the C-ABI: rulebooks bit-exact vs the oracle
(canonical order: outputs sorted by flat index = reference GPU order, pair slots ascending in the
input row = reference CPU order), conv fwd / dgrad / wgrad and dense() within 1e-3 rel (fp32),
plus size-independent properties at nuScenes scale."""
import numpy as np
import pytest
import torch

import synth
from test_oracle_spconv import GEOMS, random_voxels

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("geom", GEOMS)
def test_rulebook_bit_exact(geom):
    from ddf_b200.ops.spconv import ops
    from oracle import spconv as osp
    shape, ks, st, pad, dil, subm = geom
    idx = random_voxels(900, 2, shape, seed=3)
    o_out, o_pairs, o_num, o_shape = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm, order="gpu")
    rb = ops.build_rulebook(torch.from_numpy(idx).cuda(), 2, shape, ks, st, pad, dil, 0, subm, False)
    assert rb.out_spatial_shape == o_shape
    assert np.array_equal(rb.indice_pair_num.cpu().numpy(), o_num)
    assert np.array_equal(rb.outids.cpu().numpy(), o_out)
    assert np.array_equal(rb.indice_pairs.cpu().numpy(), o_pairs)
    # the reference-named wrapper returns the same three tensors
    outids, pairs, num = ops.get_indice_pairs(torch.from_numpy(idx).cuda(), 2, shape, ks, st, pad, dil, 0, subm)
    assert np.array_equal(pairs.cpu().numpy(), o_pairs) and np.array_equal(num.cpu().numpy(), o_num)
    # row-major tables agree with the pair lists
    G = rb.gather_table.cpu().numpy()
    GT = rb.scatter_table.cpu().numpy()
    G2 = np.full_like(G, -1)
    GT2 = np.full_like(GT, -1)
    for k in range(o_pairs.shape[0]):
        i, o = o_pairs[k, 0, :o_num[k]], o_pairs[k, 1, :o_num[k]]
        G2[o, k] = i
        GT2[i, k] = o
    assert np.array_equal(G, G2) and np.array_equal(GT, GT2)


@pytest.mark.parametrize("geom", GEOMS[:6])
@pytest.mark.parametrize("chan", [(5, 16), (16, 32), (64, 64), (128, 128), (20, 136)])
def test_conv_fwd_bwd_vs_oracle(geom, chan):
    from ddf_b200.ops.spconv import functional as Fsp, ops
    from oracle import spconv as osp
    shape, ks, st, pad, dil, subm = geom
    cin, cout = chan
    rng = np.random.default_rng(11)
    idx = random_voxels(1500, 2, shape, seed=4)
    o_out, o_pairs, o_num, _ = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm, order="gpu")
    feat = rng.standard_normal((len(idx), cin)).astype(np.float32)
    w = (rng.standard_normal((*ks, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    go = rng.standard_normal((len(o_out), cout)).astype(np.float32)
    ref = osp.indice_conv(feat, w, o_pairs, o_num, len(o_out))
    ref_gi, ref_gw = osp.indice_conv_backward(feat, w, go, o_pairs, o_num)

    rb = ops.build_rulebook(torch.from_numpy(idx).cuda(), 2, shape, ks, st, pad, dil, 0, subm, False)
    # fp32 SIMT kernels: 1e-4.  Layers that run as tcgen05 implicit GEMMs: bf16x3 forward / dgrad (default mode,
    # 16-bit significand per product, fp32 accumulation) also hold 1e-4; single-pass tf32 (wgrad; forward / dgrad in
    # mode 1) multiplies 10-bit mantissas rounded to nearest: 1e-3 relative, the north_star bar.
    kv = int(np.prod(ks))
    mode = ops.tc_mode(kv, ops.padded_cin(kv, cin, cout), cout)
    tol = 1e-3 if mode else 1e-4
    tol_f = 1e-4 if (mode & 16 or not mode & 1) else 1e-3
    tol_d = 1e-4 if (mode & 32 or not mode & 2) else 1e-3
    # (a) hot path: table-driven Function
    f = torch.from_numpy(feat).cuda().requires_grad_()
    wt = torch.from_numpy(w).cuda().requires_grad_()
    out = Fsp.table_conv(f, wt, None, rb, len(o_out))
    out.backward(torch.from_numpy(go).cuda())
    assert rel_err(out.detach().cpu().numpy(), ref) < tol_f
    assert rel_err(f.grad.cpu().numpy(), ref_gi) < tol_d
    assert rel_err(wt.grad.cpu().numpy(), ref_gw) < tol
    # (b) drop-in path: reference-named Functions on the reference-format rulebook
    f2 = torch.from_numpy(feat).cuda().requires_grad_()
    w2 = torch.from_numpy(w).cuda().requires_grad_()
    fn = Fsp.indice_subm_conv if subm else Fsp.indice_conv
    out2 = fn(f2, w2, rb.indice_pairs, rb.indice_pair_num, len(o_out))
    out2.backward(torch.from_numpy(go).cuda())
    # the reference-ABI entry points are the indice_conv*_fp32 contract: fp32 SIMT forward and backward
    assert rel_err(out2.detach().cpu().numpy(), ref) < 1e-4
    assert rel_err(f2.grad.cpu().numpy(), ref_gi) < 1e-4
    assert rel_err(w2.grad.cpu().numpy(), ref_gw) < 1e-4


@pytest.mark.parametrize("tc", [1, 4])
@pytest.mark.parametrize("chan", [(32, 32), (64, 128), (128, 64), (32, 64)])
def test_conv_precision_modes(tc, chan):
    """tf32 (mode 1) vs bf16x3 (mode 4) forward / dgrad on the multi-tile kernel against the oracle."""
    from ddf_b200 import lib
    from ddf_b200.ops.spconv import functional as Fsp, ops
    from oracle import spconv as osp
    shape, ks, st, pad, dil, subm = GEOMS[0]
    cin, cout = chan
    rng = np.random.default_rng(5)
    idx = random_voxels(2500, 2, shape, seed=6)
    o_out, o_pairs, o_num, _ = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm, order="gpu")
    feat = (rng.standard_normal((len(idx), cin)) * np.exp(rng.uniform(-4, 4, (len(idx), 1)))).astype(np.float32)
    w = (rng.standard_normal((*ks, cin, cout)) / np.sqrt(cin)).astype(np.float32)
    go = rng.standard_normal((len(o_out), cout)).astype(np.float32)
    ref = osp.indice_conv(feat, w, o_pairs, o_num, len(o_out))
    ref_gi, ref_gw = osp.indice_conv_backward(feat, w, go, o_pairs, o_num)
    prev = lib.get_lib().ddf_set_tensor_cores(tc)
    try:
        mode = ops.tc_mode(int(np.prod(ks)), cin, cout)
        assert bool(mode & 16) == (tc == 4) and bool(mode & 32) == (tc == 4)
        rb = ops.build_rulebook(torch.from_numpy(idx).cuda(), 2, shape, ks, st, pad, dil, 0, subm, False)
        f = torch.from_numpy(feat).cuda().requires_grad_()
        wt = torch.from_numpy(w).cuda().requires_grad_()
        out = Fsp.table_conv(f, wt, None, rb, len(o_out))
        out.backward(torch.from_numpy(go).cuda())
    finally:
        lib.get_lib().ddf_set_tensor_cores(prev)
    tol = 5e-5 if tc == 4 else 1e-3
    assert rel_err(out.detach().cpu().numpy(), ref) < tol
    assert rel_err(f.grad.cpu().numpy(), ref_gi) < tol
    assert rel_err(wt.grad.cpu().numpy(), ref_gw) < 1e-3     # wgrad is tf32 in both modes


def test_split_bf16x3_layout():
    """ddf_split_bf16x3: [32 x bf16 hi | 32 x bf16 lo] per 32 channels, hi + lo == x to 2^-16, tf32 copy rounded."""
    from ddf_b200.ops.spconv import ops
    x = torch.randn(257, 96, device="cuda") * torch.exp(torch.empty(257, 1, device="cuda").uniform_(-20, 20))
    split, rounded = ops.split_bf16x3(x, want_rounded=True)
    blocks = split.view(torch.bfloat16).view(257, 3, 2, 32)
    hi, lo = blocks[:, :, 0].float(), blocks[:, :, 1].float()
    xb = x.view(257, 3, 32)
    assert torch.equal(hi, xb.bfloat16().float())
    assert torch.equal(lo, (xb - hi).bfloat16().float())
    assert float(((hi + lo - xb).abs() / xb.abs().clamp_min(1e-30)).max()) < 2.0 ** -16
    assert float(((rounded - x).abs() / x.abs().clamp_min(1e-30)).max()) <= 2.0 ** -11
    assert torch.equal(rounded.view(torch.int32) & 0x1FFF, torch.zeros_like(rounded, dtype=torch.int32))


def test_large_kernel_volume_falls_back_to_fp32_wgrad():
    """5x5x5 SubM kernel (K = 125 > 100): the pair-list tensor-core wgrad keeps per-offset bookkeeping for at most
    100 offsets in shared memory, larger kernel volumes must take the fp32 kernel (and stay correct)."""
    from ddf_b200.ops.spconv import functional as Fsp, ops
    from oracle import spconv as osp
    shape, ks = [9, 20, 20], [5, 5, 5]
    rng = np.random.default_rng(3)
    idx = random_voxels(600, 1, shape, seed=8)
    o_out, o_pairs, o_num, _ = osp.get_indice_pairs(idx, 1, shape, ks, [1] * 3, [2] * 3, [1] * 3, True, order="gpu")
    feat = rng.standard_normal((len(idx), 16)).astype(np.float32)
    w = (rng.standard_normal((*ks, 16, 16)) / 4).astype(np.float32)
    go = rng.standard_normal((len(o_out), 16)).astype(np.float32)
    ref = osp.indice_conv(feat, w, o_pairs, o_num, len(o_out))
    ref_gi, ref_gw = osp.indice_conv_backward(feat, w, go, o_pairs, o_num)
    assert not ops.tc_mode(125, 16, 16) & 4
    rb = ops.build_rulebook(torch.from_numpy(idx).cuda(), 1, shape, ks, [1] * 3, [2] * 3, [1] * 3, 0, True, False)
    f = torch.from_numpy(feat).cuda().requires_grad_()
    wt = torch.from_numpy(w).cuda().requires_grad_()
    out = Fsp.table_conv(f, wt, None, rb, len(o_out))
    out.backward(torch.from_numpy(go).cuda())
    assert rel_err(out.detach().cpu().numpy(), ref) < 1e-3
    assert rel_err(f.grad.cpu().numpy(), ref_gi) < 1e-3
    assert rel_err(wt.grad.cpu().numpy(), ref_gw) < 1e-3


def test_inverse_conv_roundtrip_shapes_and_values():
    """SparseInverseConv3d reuses the saved rulebook with the roles swapped (conv.py:156-163)."""
    import ddf_b200.ops.spconv as sp
    from oracle import spconv as osp
    shape = [11, 40, 40]
    idx = random_voxels(800, 2, shape, seed=9)
    rng = np.random.default_rng(2)
    feat = rng.standard_normal((len(idx), 16)).astype(np.float32)
    x = sp.SparseConvTensor(torch.from_numpy(feat).cuda(), torch.from_numpy(idx).cuda(), shape, 2)
    down = sp.SparseConv3d(16, 32, 3, 2, padding=1, bias=False, indice_key="cp1").cuda()
    up = sp.SparseInverseConv3d(32, 16, 3, indice_key="cp1", bias=False).cuda()
    y = down(x)
    z = up(y)
    assert z.features.shape == (len(idx), 16) and z.spatial_shape == shape
    assert torch.equal(z.indices, x.indices)
    o_out, o_pairs, o_num, _ = osp.get_indice_pairs(idx, 2, shape, [3] * 3, [2] * 3, [1] * 3, [1] * 3, False, order="gpu")
    ref_y = osp.indice_conv(feat, down.weight.detach().cpu().numpy(), o_pairs, o_num, len(o_out))
    ref_z = osp.indice_conv(ref_y, up.weight.detach().cpu().numpy(), o_pairs, o_num, len(idx), inverse=True)
    assert rel_err(z.features.detach().cpu().numpy(), ref_z) < 1e-3


def test_dense_matches_oracle_and_backward():
    import ddf_b200.ops.spconv as sp
    from oracle import spconv as osp
    shape = [2, 18, 18]
    idx = random_voxels(300, 2, shape, seed=5)
    feat = np.random.default_rng(1).standard_normal((len(idx), 128)).astype(np.float32)
    f = torch.from_numpy(feat).cuda().requires_grad_()
    x = sp.SparseConvTensor(f, torch.from_numpy(idx).cuda(), shape, 2)
    d = x.dense()
    assert np.array_equal(d.detach().cpu().numpy(), osp.dense(feat, idx, shape, 2))
    g = torch.randn_like(d)
    d.backward(g)
    want = g.cpu().numpy()[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    assert np.array_equal(f.grad.cpu().numpy(), want)
    # scatter_nd keeps the reference contract too
    s = sp.scatter_nd(torch.from_numpy(idx).cuda().long(), f.detach(), [2, *shape, 128])
    assert torch.equal(s.permute(0, 4, 1, 2, 3), d.detach())


def test_empty_sparse_tensor():
    import ddf_b200.ops.spconv as sp
    x = sp.SparseConvTensor(torch.zeros(0, 16).cuda(), torch.zeros(0, 4, dtype=torch.int32).cuda(), [11, 40, 40], 2)
    y = sp.SubMConv3d(16, 16, 3, bias=False).cuda()(x)
    assert y.features.shape == (0, 16)
    z = sp.SparseConv3d(16, 32, 3, 2, padding=1, bias=False).cuda()(x)
    assert z.features.shape == (0, 32) and z.indices.shape == (0, 4)


def test_full_size_properties_nuscenes_grid():
    """~100k voxels on the [41,1440,1440] grid, batch 2: properties that need no oracle."""
    import ddf_b200.ops.spconv as sp
    from ddf_b200.ops.spconv import ops
    from ddf_b200.ops.voxel import Voxelization
    vox = Voxelization(synth.NUSC_VOXEL, synth.NUSC_RANGE, 10, 120000).eval()
    coors = []
    for b in range(2):
        _, c, _ = vox(torch.from_numpy(synth.lidar_points(250000, seed=20 + b)).cuda())
        coors.append(torch.nn.functional.pad(c, (1, 0), value=b))
    idx = torch.cat(coors).contiguous()
    n = idx.shape[0]
    shape = [41, 1440, 1440]
    rb = ops.build_rulebook(idx, 2, shape, 3, 1, 1, 1, 0, True, False)
    num = rb.indice_pair_num.cpu().numpy()
    assert num[13] == n                       # centre tap = identity (spconv_ops.h:271-303 relies on it)
    assert np.array_equal(num, num[::-1])     # SubM pair sets are mirror-symmetric
    p = rb.indice_pairs
    for k in (0, 5, 12):
        a = p[k, :, :num[k]]
        b = p[26 - k, :, :num[k]].flip(0)     # swapped roles
        sa = a[:, torch.argsort(a[0] * (n + 1) + a[1])]
        sb = b[:, torch.argsort(b[0] * (n + 1) + b[1])]
        assert torch.equal(sa, sb)
        assert bool((a[0][1:] > a[0][:-1]).all())  # ascending input rows
    # identity kernel returns the input; all-ones 1->1 kernel counts neighbours
    x = sp.SparseConvTensor(torch.randn(n, 16, device="cuda"), idx, shape, 2)
    conv = sp.SubMConv3d(16, 16, 3, bias=False).cuda()
    with torch.no_grad():
        conv.weight.zero_()
        conv.weight[1, 1, 1] = torch.eye(16)
    y = conv(x)
    # the 16-channel layers run the fp32 warp-per-row kernel: identity weights return the input bit for bit
    assert torch.equal(y.features, x.features)
    x64 = sp.SparseConvTensor(torch.randn(n, 64, device="cuda"), idx, shape, 2)
    conv64 = sp.SubMConv3d(64, 64, 3, bias=False).cuda()
    with torch.no_grad():
        conv64.weight.zero_()
        conv64.weight[1, 1, 1] = torch.eye(64)
    # tensor-core layers, bf16x3 (default): identity weights return hi + lo = the input to 16 bits, exactly
    # bf16(x) + bf16(x - bf16(x)); single-pass tf32 (mode 1) returns the tf32-rounded input
    hi = x64.features.bfloat16().float()
    assert torch.equal(conv64(x64).features, hi + (x64.features - hi).bfloat16().float())
    from ddf_b200 import lib
    prev = lib.get_lib().ddf_set_tensor_cores(1)
    try:
        assert torch.equal(conv64(x64).features, ops.round_tf32(x64.features))
    finally:
        lib.get_lib().ddf_set_tensor_cores(prev)
    ones = sp.SparseConvTensor(torch.ones(n, 1, device="cuda"), idx, shape, 2)
    c1 = sp.SubMConv3d(1, 1, 3, bias=False).cuda()
    with torch.no_grad():
        c1.weight.fill_(1.0)
    cnt = c1(ones).features[:, 0]
    assert torch.equal(cnt, (rb.gather_table >= 0).sum(1).float())
    # strided conv: outputs sorted & unique, every input reaches >= 1 output, linearity
    down = sp.SparseConv3d(16, 32, 3, 2, padding=1, bias=False).cuda()
    z = down(x)
    o = z.indices.long()
    flat = ((o[:, 0] * 21 + o[:, 1]) * 720 + o[:, 2]) * 720 + o[:, 3]
    assert z.spatial_shape == [21, 720, 720] and bool((flat[1:] > flat[:-1]).all())
    x2 = sp.SparseConvTensor(torch.randn(n, 16, device="cuda"), idx, shape, 2)
    x12 = sp.SparseConvTensor(x.features + 3 * x2.features, idx, shape, 2)
    lhs = down(x12).features
    rhs = z.features + 3 * down(x2).features
    assert (lhs - rhs).abs().max() < 1e-3 * rhs.abs().max()


@pytest.mark.parametrize("n_side,chan", [(3, 64), (9, 128), (40, 32), (71, 128), (80, 128), (97, 64), (120, 32)])
def test_tile_schedules_agree_across_kernels(n_side, chan):
    """The multi-tile tcgen05 kernel picks tiles per CTA (1..4), ring depths and CTAs per SM from the row
    count and width; the table-driven wgrad splits rows and offsets over CTAs. Row counts from 72 to
    115k rows (not multiples of the 128-row tile / 32-row sub-tile) must all give what the single-tile
    cp.async kernel and the pair-list wgrad give: the same tf32 products, so only the fp32 summation
    order may differ (<= 1e-5 of the largest value; 1e-4 for wgrad, a sum over all rows)."""
    from ddf_b200 import lib
    from ddf_b200.ops.spconv import ops
    zz, yy, xx = torch.meshgrid(torch.arange(8), torch.arange(n_side), torch.arange(n_side), indexing="ij")
    keep = ((zz * 7 + yy * 3 + xx) % 5) != 0                       # holes: not every offset in every tile
    idx = torch.stack([torch.zeros_like(zz), zz, yy, xx], -1)[keep].int().cuda().contiguous()
    n = idx.shape[0]
    g = torch.Generator().manual_seed(n_side)
    rb = ops.build_rulebook(idx, 1, [8, n_side, n_side], 3, 1, 1, 1, 0, True, False)
    feat = ops.round_tf32(torch.randn(n, chan, generator=g).cuda())
    go = ops.round_tf32(torch.randn(n, chan, generator=g).cuda())
    w = (torch.randn(3, 3, 3, chan, chan, generator=g) / (27 * chan) ** 0.5).cuda()
    L = lib.get_lib()
    prev = L.ddf_set_tensor_cores(1)
    try:
        out = ops.sparse_conv_forward(feat, w, rb.gather_table, None, n)
        gin = ops.sparse_conv_dgrad(w, go, rb.scatter_table, n)
        gw = ops.sparse_conv_wgrad_table(feat, w, go, rb.gather_table)
        L.ddf_set_tensor_cores(2)                                   # single-tile kernel, pair-list wgrad
        out_ref = ops.sparse_conv_forward(feat, w, rb.gather_table, None, n)
        gin_ref = ops.sparse_conv_dgrad(w, go, rb.scatter_table, n)
        gw_ref = ops.sparse_conv_wgrad(feat, w, go, rb.indice_pairs, rb.indice_pair_num)
    finally:
        L.ddf_set_tensor_cores(prev)
    for a, b, tol in ((out, out_ref, 1e-5), (gin, gin_ref, 1e-5), (gw, gw_ref, 1e-4)):
        assert float((a - b).abs().max()) <= tol * float(b.abs().max())


def test_bev_bf16_channels_last_handoff():
    """dense_bev_bf16() == dense().view(B, C*D, H, W) rounded to bf16, stored channels-last; backward gathers."""
    import ddf_b200.ops.spconv as sp
    shape = [2, 18, 20]
    idx = random_voxels(400, 3, shape, seed=5)
    feat = torch.randn(len(idx), 128, device="cuda", requires_grad=True)
    x = sp.SparseConvTensor(feat, torch.from_numpy(idx).cuda(), shape, 3)
    bev = x.dense_bev_bf16()
    assert bev.dtype == torch.bfloat16 and bev.shape == (3, 256, 18, 20)
    assert bev.is_contiguous(memory_format=torch.channels_last)
    ref = x.dense().view(3, 256, 18, 20)
    assert torch.equal(bev.float(), ref.detach().bfloat16().float())
    g = torch.randn(3, 256, 18, 20, device="cuda").bfloat16()
    bev.backward(g)
    g_ref, = torch.autograd.grad(ref, feat, g.float())
    assert torch.equal(feat.grad, g_ref)
