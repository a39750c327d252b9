"""GPU parity (bit-exact indices / values) of the point-set ops through the C-ABI against the
reference's known-answer vectors and the numpy oracle, plus the LocalTransformer module against the
same module driven by the oracle ops on the CPU."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(GOLDEN, "pointops_golden.npz"))


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reference_known_answers():
    from ddf_b200.ops import pointops as P
    idx = P.furthest_point_sample(cuda(G["fps0/xyz"]), int(G["fps0/npoint"]))
    assert np.array_equal(idx.cpu().numpy(), G["fps0/idx"])
    for i in range(2):
        p = "ball_query%d/" % i
        idx = P.ball_query(float(G[p + "min_r"]), float(G[p + "max_r"]), int(G[p + "nsample"]), cuda(G[p + "xyz"]), cuda(G[p + "new_xyz"]))
        assert np.array_equal(idx.cpu().numpy(), G[p + "idx"])
    out = P.grouping_operation(cuda(G["group0/features"]), cuda(G["group0/idx"]).int())
    assert np.array_equal(out.cpu().numpy(), G["group0/out"])
    out = P.gather_points(cuda(G["gather0/features"]), cuda(G["gather0/idx"]).int())
    assert np.array_equal(out.cpu().numpy(), G["gather0/out"])


def padded_cloud(B, N, n_real, seed):
    """Per-camera query layout of the fusion layer: real voxel centres then zero padding."""
    rng = np.random.default_rng(seed)
    xyz = np.zeros((B, N, 3), np.float32)
    for b in range(B):
        k = n_real[b % len(n_real)]
        xyz[b, :k] = (rng.uniform(-30, 30, (k, 3)) * [1, 1, 0.1]).astype(np.float32)
        xyz[b, :k] = np.round(xyz[b, :k] / 0.6) * 0.6 + 0.3   # voxel-centre grid -> many exact ties
    return xyz


@pytest.mark.parametrize("B,N,npoint", [(3, 700, 64), (2, 5000, 256), (2, 9000, 512), (1, 40000, 300)])
def test_fps_and_ball_query_vs_oracle(B, N, npoint):
    from ddf_b200.ops import pointops as P
    from oracle import pointops as op
    xyz = padded_cloud(B, N, [N, N * 2 // 3, N // 3], seed=N)
    idx = P.furthest_point_sample(cuda(xyz), npoint).cpu().numpy()
    ref = op.furthest_point_sample(xyz, npoint)
    assert np.array_equal(idx, ref)
    centres = np.stack([xyz[b][ref[b]] for b in range(B)])
    bq = P.ball_query(0.0, 2.0, 32, cuda(xyz), cuda(centres)).cpu().numpy()
    assert np.array_equal(bq, op.ball_query(0.0, 2.0, 32, xyz, centres))
    bq = P.ball_query(0.5, 1.5, 16, cuda(xyz), cuda(centres)).cpu().numpy()
    assert np.array_equal(bq, op.ball_query(0.5, 1.5, 16, xyz, centres))


def test_group_gather_forward_backward_vs_oracle():
    from ddf_b200.ops import pointops as P
    from oracle import pointops as op
    rng = np.random.default_rng(3)
    B, C, N, n_p, ns = 3, 37, 900, 50, 32
    f = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, n_p, ns)).astype(np.int32)
    g = rng.standard_normal((B, C, n_p, ns)).astype(np.float32)
    tf = cuda(f).requires_grad_()
    out = P.grouping_operation(tf, cuda(idx))
    out.backward(cuda(g))
    assert np.array_equal(out.detach().cpu().numpy(), op.grouping_operation(f, idx))
    np.testing.assert_allclose(tf.grad.cpu().numpy(), op.grouping_operation_grad(g, idx, N), rtol=1e-5, atol=1e-5)
    idx1 = rng.integers(0, N, (B, n_p)).astype(np.int32)
    g1 = rng.standard_normal((B, C, n_p)).astype(np.float32)
    tf = cuda(f).requires_grad_()
    out = P.gather_points(tf, cuda(idx1))
    out.backward(cuda(g1))
    assert np.array_equal(out.detach().cpu().numpy(), op.gather_points(f, idx1))
    np.testing.assert_allclose(tf.grad.cpu().numpy(), op.gather_points_grad(g1, idx1, N), rtol=1e-5, atol=1e-5)


def test_local_transformer_matches_oracle_cpu_path():
    from ddf_b200.fusion.pointformer import LocalTransformer
    from oracle import cpu_path
    torch.manual_seed(0)
    m_cpu = LocalTransformer(128, 2.0, 32, 64, 64, num_layers=2).eval()
    m_gpu = copy.deepcopy(m_cpu).cuda().eval()
    xyz = torch.from_numpy(padded_cloud(3, 1500, [1500, 900, 400], seed=5))
    feats = torch.randn(3, 64, 1500)
    with torch.no_grad():
        with cpu_path.reference_cpu_ops():
            ref = m_cpu(xyz, feats.clone())
        out = m_gpu(xyz.cuda(), feats.clone().cuda()).cpu()
    assert out.shape == ref.shape == (3, 1500, 64)
    assert float((out - ref).abs().max()) < 1e-3 * float(ref.abs().max())
    # voxels outside every ball keep their input feature (feat_agg_method='replace')
    untouched = (out == feats.permute(0, 2, 1)).all(-1)
    assert bool(untouched.any()) and not bool(untouched.all())


def test_scatter_first_occurrence_kernel_matches_reference_rule():
    """LocalTransformer.scatter 'unique' rule (pointformer.py:319-347): first occurrence in flattened (group, slot)
    order wins; forward and backward against the host implementation (which reproduces the reference's
    unique + flip + scatter_ result on the CPU)."""
    from ddf_b200.fusion.pointformer import first_occurrence_scatter
    from ddf_b200.ops import pointops
    torch.manual_seed(0)
    B, C, N, npnt, ns = 3, 16, 500, 64, 8
    idx = torch.randint(0, N // 2, (B, npnt, ns), dtype=torch.int32)          # duplicates; the upper half is never hit
    idx[0, 0, :] = torch.tensor([5, 3, 5, 7, 3, 3, 9, 5], dtype=torch.int32)  # SURVEY 3.3 example
    feats = torch.randn(B, C, npnt, ns)
    base = torch.randn(B, C, N)
    ref_base = base.clone().requires_grad_()
    ref_feats = feats.clone().requires_grad_()
    ref = first_occurrence_scatter(ref_base.clone(), ref_feats, idx)
    g = torch.randn(B, C, N)
    ref.backward(g)
    d_base = base.cuda().requires_grad_()
    d_feats = feats.cuda().requires_grad_()
    out = pointops.scatter_first(d_base, d_feats, idx.cuda())
    out.backward(g.cuda())
    assert torch.equal(out.detach().cpu(), ref.detach())
    assert torch.equal(d_base.grad.cpu(), ref_base.grad) and torch.equal(d_feats.grad.cpu(), ref_feats.grad)
    # the worked example: voxels 5, 3, 7, 9 take flattened positions 0, 1, 3, 6 of group 0
    for v, pos in ((5, 0), (3, 1), (7, 3), (9, 6)):
        assert torch.equal(out[0, :, v].cpu(), feats[0, :, 0, pos])
    assert torch.equal(d_base.detach().cpu(), base)   # input not modified


@pytest.mark.parametrize("C,heads", [(128, 4), (64, 4)])
def test_local_attention_kernel_matches_multihead_attention(C, heads):
    """csrc/local_attn.cu vs nn.MultiheadAttention's own attention (scaled_dot_product) on groups of 32 tokens."""
    from ddf_b200.ops import pointops
    torch.manual_seed(0)
    G, ns = 37, 32
    qkv = torch.randn(G * ns, 3 * C, device="cuda", dtype=torch.float64)
    hd = C // heads
    q, k, v = (t.view(G, ns, heads, hd).transpose(1, 2) for t in qkv.split(C, dim=1))
    qkv32 = qkv.float().requires_grad_()
    out = pointops.local_attention(qkv32, heads, ns)
    g = torch.randn(G * ns, C, device="cuda")
    out.backward(g)
    q, k, v = (t.detach().requires_grad_() for t in (q, k, v))
    ref = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, -1) @ v
    ref = ref.transpose(1, 2).reshape(G * ns, C)
    ref.backward(g.double())
    ref_g = torch.cat([t.grad.transpose(1, 2).reshape(G * ns, C) for t in (q, k, v)], 1)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((qkv32.grad.double() - ref_g).abs().max() / ref_g.abs().max()) < 1e-5


def test_local_transformer_token_path_equals_module_graph():
    """LocalTransformer on CUDA (token-major path, csrc/local_attn.cu, cached geometry) == the same module through
    the generic module graph (nn.MultiheadAttention on the permuted (32, B*np, C) tensor), forward and backward."""
    from ddf_b200.fusion.pointformer import LocalTransformer
    torch.backends.cudnn.allow_tf32 = False      # the module graph runs its 1x1 position convs through cuDNN
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    lt = LocalTransformer(64, 2.0, 32, 128, 128, num_layers=2).cuda().train()
    B, N = 3, 900
    xyz = (torch.rand(B, N, 3, device="cuda") * torch.tensor([20.0, 20.0, 4.0], device="cuda")).contiguous()
    xyz[:, -100:] = 0        # padded rows at the origin
    feats = torch.randn(B, N, 128, device="cuda")
    f1 = feats.clone().requires_grad_()
    out1 = lt(xyz, f1.permute(0, 2, 1))
    out1.square().sum().backward()
    g1 = {n: p.grad.clone() for n, p in lt.named_parameters()}
    lt.zero_grad()
    for m in lt.modules():                     # same batch statistics update twice: reset the momentum side effects
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
    f2 = feats.clone().requires_grad_()
    lt._token_path_ok = lambda features: False
    out2 = lt(xyz, f2.permute(0, 2, 1).contiguous())
    out2.square().sum().backward()
    assert float((out1 - out2).abs().max() / out2.abs().max()) < 1e-4
    assert float((f1.grad - f2.grad).abs().max() / f2.grad.abs().max()) < 1e-3
    for n, p in lt.named_parameters():
        assert float((g1[n] - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-12)) < 2e-3, n
