"""GPU parity: ddf_ms_deform_attn_forward/backward (through the C-ABI, via the reference-named
MSDeformAttnFunction) against (i) golden vectors from the reference's pure-PyTorch MSDA,
(ii) the C oracle on seeded inputs, (iii) size-independent properties at full hot-path size.
Tolerance: fp32 1e-3 relative (north_star), fp64 1e-9."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(GOLDEN, "msda_golden.npz"))
CASES = sorted({k.split("/")[0] for k in G.files})


def case(name):
    return {k.split("/")[1]: G[k] for k in G.files if k.startswith(name + "/")}


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def run_cuda(c, im2col_step=64):
    from ddf_b200.ops.msda import MSDeformAttnFunction
    dev = "cuda:0"
    value = torch.from_numpy(c["value"]).to(dev).requires_grad_(True)
    loc = torch.from_numpy(c["loc"]).to(dev).requires_grad_(True)
    attn = torch.from_numpy(c["attn"]).to(dev).requires_grad_(True)
    shapes = torch.from_numpy(np.asarray(c["shapes"], np.int64)).to(dev)
    lsi = torch.from_numpy(np.asarray(c["lsi"], np.int64)).to(dev)
    out = MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, im2col_step)
    out.backward(torch.from_numpy(c["gout"]).to(dev))
    return (out.detach().cpu().numpy(), value.grad.cpu().numpy(), loc.grad.cpu().numpy(),
            attn.grad.cpu().numpy())


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_golden(name):
    c = case(name)
    out, gv, gl, ga = run_cuda(c, im2col_step=2 if name.startswith("reftest") else 64)
    tol = 1e-9 if c["value"].dtype == np.float64 else 1e-3
    assert rel_err(out, c["out"]) < tol
    assert rel_err(gv, c["gvalue"]) < tol
    assert rel_err(gl, c["gloc"]) < tol
    assert rel_err(ga, c["gattn"]) < tol
    if c["value"].dtype == np.float32:
        # and elementwise at the reference test's own fp32 tolerance (ops/test.py:57)
        np.testing.assert_allclose(out, c["out"], rtol=1e-2, atol=1e-3)


def synth(N, M, D, Lq, P, shapes, seed, lo=-0.1, hi=1.1, dtype=np.float32):
    rng = np.random.default_rng(seed)
    shapes = np.asarray(shapes, np.int64)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    lsi = np.concatenate([[0], np.cumsum(shapes.prod(1))[:-1]]).astype(np.int64)
    value = rng.standard_normal((N, S, M, D)).astype(dtype)
    loc = rng.uniform(lo, hi, (N, Lq, M, L, P, 2)).astype(dtype)
    attn = rng.random((N, Lq, M, L, P)).astype(dtype) + 1e-5
    attn /= attn.sum((-1, -2), keepdims=True)
    gout = rng.standard_normal((N, Lq, M * D)).astype(dtype)
    return dict(value=value, shapes=shapes, lsi=lsi, loc=loc, attn=attn.astype(dtype), gout=gout)


@pytest.mark.parametrize("cfg", [
    # BASELINE config 1: 6 cams, 64x176 map, M=8, D=16, L=1, P=4
    dict(N=6, M=8, D=16, Lq=1500, P=4, shapes=[(64, 176)]),
    # Voxel-RCNN head (D=8), 1 camera 94x311
    dict(N=2, M=8, D=8, Lq=2000, P=4, shapes=[(94, 311)]),
    # multi-level, P not a multiple of 4, D=32
    dict(N=3, M=4, D=32, Lq=301, P=5, shapes=[(20, 30), (10, 15), (5, 8)]),
    # D=4 and D=64, D=128 lanes-per-head extremes
    dict(N=1, M=2, D=4, Lq=77, P=4, shapes=[(9, 11)]),
    dict(N=1, M=2, D=64, Lq=77, P=2, shapes=[(9, 11), (4, 5)]),
    dict(N=1, M=1, D=128, Lq=33, P=1, shapes=[(7, 9)]),
    # generic path: D not a multiple of 4
    dict(N=2, M=3, D=6, Lq=41, P=3, shapes=[(9, 11), (4, 5)]),
    dict(N=1, M=2, D=71, Lq=19, P=2, shapes=[(6, 4), (3, 2)]),
])
def test_cuda_matches_oracle_seeded(cfg):
    from oracle import msda as omsda
    c = synth(seed=17, **cfg)
    out, gv, gl, ga = run_cuda(c, im2col_step=64)
    o_out = omsda.msda_forward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    o_gv, o_gl, o_ga = omsda.msda_backward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"], c["gout"])
    assert rel_err(out, o_out) < 1e-5
    assert rel_err(ga, o_ga) < 1e-4
    assert rel_err(gl, o_gl) < 1e-4
    assert rel_err(gv, o_gv) < 1e-4  # atomics reorder the fp32 sums
    # exactly the same set of touched pixels: zero pattern of grad_value must agree
    assert np.array_equal(o_gv != 0, gv != 0)


def test_fp64_matches_oracle_and_gradcheck():
    from oracle import msda as omsda
    from ddf_b200.ops.msda import MSDeformAttnFunction
    c = synth(N=1, M=2, D=30, Lq=4, P=2, shapes=[(6, 4), (3, 2)], seed=3, lo=0.0, hi=1.0, dtype=np.float64)
    out, gv, gl, ga = run_cuda(c, im2col_step=2)
    assert rel_err(out, omsda.msda_forward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])) < 1e-12
    o_gv, o_gl, o_ga = omsda.msda_backward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"], c["gout"])
    assert rel_err(gv, o_gv) < 1e-12 and rel_err(gl, o_gl) < 1e-12 and rel_err(ga, o_ga) < 1e-12
    # reference's own check (ops/test.py:62-78): numerical gradcheck in fp64
    dev = "cuda:0"
    t = lambda k: torch.from_numpy(c[k]).to(dev)
    value, loc, attn = t("value").requires_grad_(), t("loc").requires_grad_(), t("attn").requires_grad_()
    assert torch.autograd.gradcheck(MSDeformAttnFunction.apply,
                                    (value, t("shapes"), t("lsi"), loc, attn, 2), nondet_tol=1e-9)


def test_full_size_properties_transfusion_shape():
    """C-TF sizes (N=12 cams, 112x200 map, Lq=8000): size-independent properties instead of the oracle."""
    from ddf_b200.ops.msda import MSDeformAttnFunction
    dev = "cuda:0"
    torch.manual_seed(0)
    N, H, W, M, D, Lq, P = 12, 112, 200, 8, 16, 8000, 4
    S = H * W
    shapes = torch.tensor([[H, W]], dtype=torch.long, device=dev)
    lsi = torch.zeros(1, dtype=torch.long, device=dev)
    value = torch.randn(N, S, M, D, device=dev)
    # (1) identity sampling: every point at a pixel centre, weights 1/P -> out == value at that pixel
    py = torch.randint(0, H, (N, Lq), device=dev)
    px = torch.randint(0, W, (N, Lq), device=dev)
    loc = torch.stack([(px + 0.5) / W, (py + 0.5) / H], -1)[:, :, None, None, None, :].expand(N, Lq, M, 1, P, 2).contiguous()
    attn = torch.full((N, Lq, M, 1, P), 1.0 / P, device=dev)
    out = MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, 64)
    want = value[torch.arange(N, device=dev)[:, None], py * W + px].reshape(N, Lq, M * D)
    assert (out - want).abs().max() < 1e-3 * want.abs().max()
    # (2) linearity in value and in attention weights
    loc = (torch.rand(N, Lq, M, 1, P, 2, device=dev) * 1.2 - 0.1).requires_grad_()
    attn = torch.softmax(torch.randn(N, Lq, M, 1 * P, device=dev), -1).view(N, Lq, M, 1, P).requires_grad_()
    v2 = torch.randn_like(value)
    f = lambda v, a: MSDeformAttnFunction.apply(v, shapes, lsi, loc, a, 64)
    o1, o2, o12 = f(value, attn), f(v2, attn), f(value + 2 * v2, attn)
    assert (o12 - (o1 + 2 * o2)).abs().max() < 1e-3 * o12.abs().max()
    # (3) checksum of grad_value: interior points only -> sum_s grad_value[b,:,m,c] == sum_q gout[b,q,m,c]
    loc_in = (torch.rand(N, Lq, M, 1, P, 2, device=dev) * 0.9 + 0.05)
    value.requires_grad_()
    o = MSDeformAttnFunction.apply(value, shapes, lsi, loc_in, attn.detach(), 64)
    gout = torch.randn_like(o)
    o.backward(gout)
    lhs = value.grad.sum(1).reshape(N, M * D)
    rhs = gout.sum(1)
    assert (lhs - rhs).abs().max() < 1e-3 * rhs.abs().max()


def test_empty_and_errors():
    from ddf_b200.ops.msda import MSDeformAttnFunction
    dev = "cuda:0"
    shapes = torch.tensor([[4, 5]], dtype=torch.long, device=dev)
    lsi = torch.zeros(1, dtype=torch.long, device=dev)
    v = torch.randn(2, 20, 2, 8, device=dev, requires_grad=True)
    out = MSDeformAttnFunction.apply(v, shapes, lsi, torch.zeros(2, 0, 2, 1, 4, 2, device=dev),
                                     torch.zeros(2, 0, 2, 1, 4, device=dev), 64)
    assert out.shape == (2, 0, 16)
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDeformAttnFunction.apply(v.detach().transpose(2, 3).contiguous().transpose(2, 3), shapes, lsi,
                                   torch.zeros(2, 3, 2, 1, 4, 2, device=dev), torch.zeros(2, 3, 2, 1, 4, device=dev), 64)
    v6 = torch.randn(6, 20, 2, 8, device=dev)
    with pytest.raises(RuntimeError, match="must divide"):
        MSDeformAttnFunction.apply(v6, shapes, lsi, torch.zeros(6, 3, 2, 1, 4, 2, device=dev),
                                   torch.zeros(6, 3, 2, 1, 4, device=dev), 4)
