"""Fused FFN-hidden and add+dropout+LayerNorm kernels against the reference's module chains
(actr_transformer.py:383-397) in float64 on the CPU; dropout checked through its invariants (kept
fraction, scaling, identical pattern in forward and backward). Tolerances: 1e-5 of the largest value for
outputs and input gradients, 1e-4 for parameter gradients (long sums)."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("rows,C,F_", [(1, 128, 256), (1000, 128, 1024), (4097, 256, 512)])
def test_ffn_hidden_eval_matches_module_chain(rows, C, F_):
    from ddf_b200.ops.fused import ffn_hidden
    torch.manual_seed(rows)
    lin, drop = nn.Linear(C, F_), nn.Dropout(0.1).eval()
    x = torch.randn(2, rows, C)
    go = torch.randn(2, rows, F_)
    lr = copy.deepcopy(lin).double()
    xr = x.double().requires_grad_()
    ref = torch.relu(lr(xr))
    ref.backward(go.double())
    ld = copy.deepcopy(lin).cuda()
    xd = x.cuda().requires_grad_()
    torch.backends.cuda.matmul.allow_tf32 = False
    out = ffn_hidden(ld, drop, xd)
    out.backward(go.cuda())
    assert rel(out.detach(), ref.detach()) < 1e-5
    assert rel(xd.grad, xr.grad) < 1e-5
    assert rel(ld.weight.grad, lr.weight.grad) < 1e-4
    assert rel(ld.bias.grad, lr.bias.grad) < 1e-4


def test_ffn_hidden_dropout_invariants():
    from ddf_b200.ops.fused import ffn_hidden
    torch.manual_seed(0)
    lin, drop = nn.Linear(128, 1024).cuda(), nn.Dropout(0.25).train()
    with torch.no_grad():
        lin.bias.fill_(5.0)          # every pre-activation positive: zeros in the output are drops
        lin.weight.mul_(0.01)
    x = torch.randn(4000, 128, device="cuda", requires_grad=True)
    out = ffn_hidden(lin, drop, x)
    pre = torch.nn.functional.linear(x.detach(), lin.weight, lin.bias)
    kept = out != 0
    assert abs(float(kept.float().mean()) - 0.75) < 0.005
    assert torch.allclose(out[kept], pre[kept] / 0.75, rtol=1e-6, atol=0)
    out.backward(torch.ones_like(out))
    # d out / d bias = kept / (1 - p), summed over the rows
    assert torch.allclose(lin.bias.grad, kept.float().sum(0) / 0.75, rtol=1e-5)
    out2 = ffn_hidden(lin, drop, x)                      # a new call draws a new pattern
    assert not torch.equal(out2 != 0, kept)


@pytest.mark.parametrize("rows,C", [(1, 128), (777, 128), (50001, 128), (300, 256), (65, 512), (1, 64), (50003, 64),
                                    (131072, 64), (9, 32), (4099, 32)])
@pytest.mark.parametrize("with_b", [True, False])
def test_add_layer_norm_eval_matches_module_chain(rows, C, with_b):
    from ddf_b200.ops.fused import add_dropout_layer_norm
    torch.manual_seed(rows + C)
    norm, drop = nn.LayerNorm(C), nn.Dropout(0.1).eval()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_()
    a, b, go = torch.randn(rows, C) * 2 + 1, torch.randn(rows, C), torch.randn(rows, C)
    nr = copy.deepcopy(norm).double()
    ar, br = a.double().requires_grad_(), b.double().requires_grad_()
    ref = nr(ar + br) if with_b else nr(ar)
    ref.backward(go.double())
    nd = copy.deepcopy(norm).cuda()
    ad, bd = a.cuda().requires_grad_(), b.cuda().requires_grad_()
    out = add_dropout_layer_norm(nd, drop, ad, bd if with_b else None)
    out.backward(go.cuda())
    assert rel(out.detach(), ref.detach()) < 1e-5
    assert rel(ad.grad, ar.grad) < 1e-5
    if with_b:
        assert rel(bd.grad, br.grad) < 1e-5
    assert rel(nd.weight.grad, nr.weight.grad) < 1e-4
    assert rel(nd.bias.grad, nr.bias.grad) < 1e-4


def test_add_dropout_layer_norm_train_consistency():
    from ddf_b200.ops.fused import add_dropout_layer_norm
    torch.manual_seed(3)
    norm, drop = nn.LayerNorm(128).cuda(), nn.Dropout(0.5).train()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
    a = torch.zeros(20000, 128, device="cuda", requires_grad=True)
    b = (torch.rand(20000, 128, device="cuda") + 1.0).requires_grad_()
    out = add_dropout_layer_norm(norm, drop, a, b)
    out.backward(torch.randn_like(out))
    # the branch gradient is zero exactly where the element was dropped; about half are kept
    kept = b.grad != 0
    assert abs(float(kept.float().mean()) - 0.5) < 0.01
    # forward used the same pattern: reconstruct s = dropout(b) and compare with LayerNorm of it
    s = torch.where(kept, b.detach() * 2.0, torch.zeros_like(b))
    ref = torch.nn.functional.layer_norm(s, (128,), norm.weight, norm.bias, norm.eps)
    # rows where a kept element has exactly zero gradient would be mis-detected; compare robustly
    ok = (out.detach() - ref).abs().max(1).values < 1e-4
    assert float(ok.float().mean()) > 0.99


def test_unsupported_shapes_use_library_ops_and_cpu_is_refused():
    from ddf_b200.ops.fused import add_dropout_layer_norm, ffn_hidden
    norm, drop = nn.LayerNorm(96).cuda(), nn.Dropout(0.0)
    x = torch.randn(10, 96, device="cuda")
    assert torch.allclose(add_dropout_layer_norm(norm, drop, x, x), norm(x + x))
    with pytest.raises(RuntimeError):
        ffn_hidden(nn.Linear(128, 256), drop, torch.randn(4, 128))


def test_linear_with_fast_bias_gradient_equals_nn_linear():
    from ddf_b200.ops import fused
    torch.manual_seed(0)
    lin = torch.nn.Linear(128, 64).cuda()
    x = torch.randn(3, 4000, 128, device="cuda", requires_grad=True)
    y = fused.linear(lin, x)
    g = torch.randn_like(y)
    y.backward(g)
    got = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None
    lin.zero_grad()
    y2 = lin(x)
    y2.backward(g)
    assert torch.equal(y, y2)
    for a, b in zip(got, (x.grad, lin.weight.grad, lin.bias.grad)):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    big = torch.randn(146016, 128, device="cuda")
    assert float((fused.col_sum(big) - big.double().sum(0).float()).abs().max()) < 1e-2


@pytest.mark.parametrize("K,M,N", [(146016, 128, 1024), (146016, 1024, 128), (20000, 128, 128), (5000, 32, 64),
                                   (4097, 96, 288), (30011, 256, 128), (12345, 64, 32)])
def test_xty_matches_matmul(K, M, N):
    """ddf_xty_tf32 (weight gradient of a Linear over K tokens, split-K tcgen05) against the fp64 product: exact on
    integer-valued operands (every tf32 product and fp32 partial sum is exact, so any layout / swizzle / split-K
    mistake shows), tf32-class on normal data (operands are truncated to tf32 by the hardware)."""
    from ddf_b200.ops import fused
    from ddf_b200 import lib as _lib
    assert _lib.get_lib().ddf_xty_supported(K, M, N)
    g = torch.Generator(device="cuda").manual_seed(K + M + N)
    a = torch.randint(-3, 4, (K, M), device="cuda", generator=g).float()
    b = torch.randint(-3, 4, (K, N), device="cuda", generator=g).float()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = fused.xty(a, b)
        ref = (a.double().t() @ b.double())
        assert torch.equal(out.double(), ref)
        a = torch.randn(K, M, device="cuda", generator=g)
        b = torch.randn(K, N, device="cuda", generator=g)
        out = fused.xty(a, b)
        ref = a.double().t() @ b.double()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        assert err < 4e-3, err
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_linear_weight_grad_through_xty():
    """fused.linear / ffn_hidden route W.grad through ddf_xty_tf32 when tf32 is allowed: gradients against the stock
    modules in fp64."""
    from ddf_b200.ops import fused
    torch.manual_seed(3)
    lin = torch.nn.Linear(128, 256).cuda()
    x = torch.randn(3, 5000, 128, device="cuda", requires_grad=True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        y = fused.linear(lin, x)
        go = torch.randn_like(y)
        y.backward(go)
        gw, gb, gx = lin.weight.grad.clone(), lin.bias.grad.clone(), x.grad.clone()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    lin64 = torch.nn.Linear(128, 256).cuda().double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    x64 = x.detach().double().requires_grad_()
    lin64(x64).backward(go.double())
    for got, ref in ((gw, lin64.weight.grad), (gb, lin64.bias.grad), (gx, x64.grad)):
        assert float((got.double() - ref).abs().max() / ref.abs().max()) < 4e-3


@pytest.mark.parametrize("cls_name", ["BiGateSum1D", "BiGateSum1D_2"])
@pytest.mark.parametrize("C,drop_second", [(128, False), (256, False), (128, True)])
def test_bigate_sum_kernel_matches_module_chain(cls_name, C, drop_second):
    """The fused gate (ddf_bigate_sum_*) against the module chain of attentions.py:89-117 in fp64: outputs and every
    gradient; ``drop_second``: the second output never reaches the loss (the last encoder layer)."""
    from ddf_b200.fusion import attentions
    torch.manual_seed(5)
    gate = getattr(attentions, cls_name)(C, C).cuda()
    f1 = torch.randn(3, 2500, C, device="cuda", requires_grad=True)
    f2 = torch.randn(3, 2500, C, device="cuda", requires_grad=True)
    o1, o2 = gate(f1, f2)
    g1, g2 = torch.randn_like(o1), torch.randn_like(o2)
    loss = (o1 * g1).sum() if drop_second else (o1 * g1).sum() + (o2 * g2).sum()
    loss.backward()
    ref = getattr(attentions, cls_name)(C, C).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in gate.state_dict().items()})
    r1, r2 = f1.detach().double().requires_grad_(), f2.detach().double().requires_grad_()
    s1 = torch.sigmoid(torch.nn.functional.linear(r1 + r2 if cls_name.endswith("_2") else r1,
                                                  ref.b_conv1d.weight.squeeze(-1), ref.b_conv1d.bias))
    s2 = torch.sigmoid(torch.nn.functional.linear(r1 + r2 if cls_name.endswith("_2") else r2,
                                                  ref.a_conv1d.weight.squeeze(-1), ref.a_conv1d.bias))
    q1, q2 = r1 + r2 * s1, r2 + r1 * s2
    rl = (q1 * g1.double()).sum() if drop_second else (q1 * g1.double()).sum() + (q2 * g2.double()).sum()
    rl.backward()
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(o1, q1) < 1e-5 and rel(o2, q2) < 1e-5
    assert rel(f1.grad, r1.grad) < 1e-4 and rel(f2.grad, r2.grad) < 1e-4
    assert rel(gate.b_conv1d.weight.grad, ref.b_conv1d.weight.grad) < 1e-3
    assert rel(gate.b_conv1d.bias.grad, ref.b_conv1d.bias.grad) < 1e-3
    if drop_second:
        assert gate.a_conv1d.weight.grad is None and gate.a_conv1d.bias.grad is None
    else:
        assert rel(gate.a_conv1d.weight.grad, ref.a_conv1d.weight.grad) < 1e-3
        assert rel(gate.a_conv1d.bias.grad, ref.a_conv1d.bias.grad) < 1e-3


@pytest.mark.parametrize("rows,C", [(5000, 192), (131072, 192), (4097, 384), (70000, 8), (33333, 1024)])
def test_col_sum_any_width(rows, C):
    """ddf_col_sum for widths whose quarter does not divide the CTA (192 = the qkv projection of the 64-channel
    LocalTransformer): idle threads instead of a fallback to ATen's sum(0)."""
    from ddf_b200.ops import fused
    x = torch.randn(rows, C, device="cuda")
    ref = x.double().sum(0)
    got = fused.col_sum(x)
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("T,F_", [(128, 64), (5000, 256), (146016, 1024), (4097, 1024)])
def test_fused_ffn_forward_matches_module_chain(T, F_):
    """ddf_ffn_forward (linear1 -> bias / ReLU / dropout -> linear2 in one kernel) against the module chain of
    actr_transformer.py:383-397 in fp64 (eval mode: no dropout), outputs and every gradient at tf32 accuracy; exact
    on integer-valued operands (layout / swizzle / pipeline mistakes show as gross errors there)."""
    from ddf_b200.ops import fused
    from ddf_b200 import lib as _lib
    torch.manual_seed(T + F_)
    l1, l2, drop = nn.Linear(128, F_).cuda(), nn.Linear(F_, 128).cuda(), nn.Dropout(0.1).eval()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        # exact case: small integers everywhere
        with torch.no_grad():
            for p_ in (l1.weight, l1.bias, l2.weight, l2.bias):
                p_.copy_(torch.randint(-2, 3, p_.shape, device="cuda").float())
        x = torch.randint(-2, 3, (T, 128), device="cuda").float()
        h = torch.empty(T, F_, device="cuda")
        y = torch.empty(T, 128, device="cuda")
        ws = torch.empty(2 * 128 * F_, device="cuda")
        assert _lib.get_lib().ddf_ffn_workspace_bytes(128, F_) == ws.numel() * 4
        rc = _lib.get_lib().ddf_ffn_forward(_lib.ptr(x), _lib.ptr(l1.weight), _lib.ptr(l1.bias), _lib.ptr(l2.weight),
                                            _lib.ptr(l2.bias), _lib.ptr(h), _lib.ptr(y), _lib.ptr(ws), T, 128, F_, 0.0, 0,
                                            _lib.current_stream())
        _lib.check(rc, "ffn_forward")
        href = torch.relu(x.double() @ l1.weight.double().t() + l1.bias.double())
        yref = href @ l2.weight.double().t() + l2.bias.double()
        assert torch.equal(h.double(), href)
        assert float((y.double() - yref).abs().max()) <= 1e-6 * float(yref.abs().max()) * F_ ** 0.5 + 1e-3
        # normal data through the autograd path.  The reference takes the ReLU decisions of the kernel (a unit whose
        # pre-activation is within tf32 rounding of zero may fall on either side; those are checked separately)
        torch.nn.init.xavier_uniform_(l1.weight); torch.nn.init.xavier_uniform_(l2.weight)
        with torch.no_grad():
            l1.bias.normal_(); l2.bias.normal_()
        x = torch.randn(T, 128, device="cuda", requires_grad=True)
        _lib.check(_lib.get_lib().ddf_ffn_forward(_lib.ptr(x.detach()), _lib.ptr(l1.weight), _lib.ptr(l1.bias),
                                                  _lib.ptr(l2.weight), _lib.ptr(l2.bias), _lib.ptr(h), _lib.ptr(y), _lib.ptr(ws),
                                                  T, 128, F_, 0.0, 0, _lib.current_stream()), "ffn_forward")
        on = h != 0
        out = fused.ffn(l1, drop, l2, x) if T >= 4096 else fused._FusedFFN.apply(x, l1.weight, l1.bias, l2.weight, l2.bias, 0.0)
        go = torch.randn_like(out)
        out.backward(go)
        r1, r2 = copy.deepcopy(l1).double(), copy.deepcopy(l2).double()
        r1.zero_grad(); r2.zero_grad()
        xr = x.detach().double().requires_grad_()
        pre = r1(xr)
        flips = on != (pre.detach() > 0)
        assert float(flips.float().mean()) < 2e-3 and float(pre.detach()[flips].abs().max() if bool(flips.any()) else 0) < 2e-2
        ref = r2(pre * on)
        ref.backward(go.double())
        assert rel(out.detach(), ref.detach().cpu()) < 3e-3
        assert rel(x.grad, xr.grad.cpu()) < 3e-3
        for got, want in ((l1.weight.grad, r1.weight.grad), (l1.bias.grad, r1.bias.grad), (l2.weight.grad, r2.weight.grad),
                          (l2.bias.grad, r2.bias.grad)):
            assert rel(got, want.cpu()) < 4e-3
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_fused_ffn_dropout_statistics_and_backward():
    """Training mode: the fused kernel keeps a fraction 1 - p of the positive pre-activations (scaled by 1 / (1 - p)),
    a different pattern per seed, and the autograd path uses exactly that pattern in backward."""
    from ddf_b200 import lib as _lib
    from ddf_b200.ops import fused
    torch.manual_seed(1)
    T, F_ = 8192, 512
    l1, l2 = nn.Linear(128, F_).cuda(), nn.Linear(F_, 128).cuda()
    x = torch.randn(T, 128, device="cuda")
    h = torch.empty(T, F_, device="cuda")
    y = torch.empty(T, 128, device="cuda")
    L = _lib.get_lib()
    ws = torch.empty(2 * 128 * F_, device="cuda")
    pats = []
    for seed in (1234, 99):
        _lib.check(L.ddf_ffn_forward(_lib.ptr(x), _lib.ptr(l1.weight), _lib.ptr(l1.bias), _lib.ptr(l2.weight),
                                     _lib.ptr(l2.bias), _lib.ptr(h), _lib.ptr(y), _lib.ptr(ws), T, 128, F_, 0.25, seed,
                                     _lib.current_stream()), "ffn_forward")
        pre = x.double() @ l1.weight.double().t() + l1.bias.double()
        pos = pre > 1e-2
        kept = (h != 0) & pos
        assert abs(float(kept.float().sum() / pos.float().sum()) - 0.75) < 0.005
        assert float((h[kept].double() - pre[kept] / 0.75).abs().max()) < 2e-2
        assert float((h[pre < -1e-2]).abs().max()) == 0.0
        # per column and per row the rate is right too (no stripes)
        assert float((kept.float().sum(0) / pos.float().sum(0).clamp_min(1) - 0.75).abs().max()) < 0.05
        pats.append(kept.clone())
        yref = h.double() @ l2.weight.double().t() + l2.bias.double()
        assert float((y.double() - yref).abs().max() / yref.abs().max()) < 3e-3
    assert 0.5 < float((pats[0] == pats[1]).float().mean()) < 0.9
    # autograd: the gradient wrt x is zero through dropped units - compare with a reference that masks with h != 0
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        xg = x.clone().requires_grad_()
        drop = nn.Dropout(0.25).train()
        torch.manual_seed(7)
        out = fused.ffn(l1, drop, l2, xg)
        go = torch.randn_like(out)
        out.backward(go)
        # recover the pattern from the output is not possible; check consistency instead: out == h' W2^T + b2 with the
        # h' that backward saved, through the bias gradient of linear1 (sum over tokens of the masked hidden gradient)
        assert l1.bias.grad is not None and float(l1.bias.grad.abs().sum()) > 0
        assert xg.grad is not None and bool(torch.isfinite(xg.grad).all())
        # the backward is the derivative of THIS forward: same pattern, and the kernel's quantised 1 / (1 - p)
        # (p = 0.1 drops 26 / 256: kept values are scaled by 256 / 230, not by 1 / 0.9)
        from unittest import mock
        drop1 = nn.Dropout(0.1).train()
        with mock.patch.object(fused, "_seed", return_value=4242):
            l1.zero_grad(); l2.zero_grad()
            xg2 = x.clone().requires_grad_()
            fused.ffn(l1, drop1, l2, xg2).backward(go)
        _lib.check(L.ddf_ffn_forward(_lib.ptr(x), _lib.ptr(l1.weight), _lib.ptr(l1.bias), _lib.ptr(l2.weight),
                                     _lib.ptr(l2.bias), _lib.ptr(h), _lib.ptr(y), _lib.ptr(ws), T, 128, F_, 0.1, 4242,
                                     _lib.current_stream()), "ffn_forward")
        unscaled = (go.double() @ l2.weight.double()) * (h != 0)
        assert abs(L.ddf_ffn_dropout_p(0.1) - 26.0 / 256.0) < 1e-7 and L.ddf_ffn_dropout_p(0.0) == 0.0
        gb1 = l1.bias.grad.double()
        ls = float((gb1 * unscaled.sum(0)).sum() / (unscaled.sum(0) ** 2).sum())
        assert abs(ls - 256.0 / 230.0) < 8e-4, ls
        gh_ref = unscaled * (256.0 / 230.0)
        assert rel(xg2.grad, (gh_ref @ l1.weight.double()).cpu()) < 4e-3
        assert rel(l1.weight.grad, (gh_ref.t() @ x.double()).cpu()) < 4e-3
        assert rel(gb1, gh_ref.sum(0).cpu()) < 4e-3
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("N,C,H,W", [(1, 32, 1, 1), (2, 256, 7, 9), (3, 100, 33, 31), (12, 256, 112, 200)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nchw_to_rows_is_the_transpose(N, C, H, W, dtype):
    """ddf_nchw_to_rows: (N, C, H, W) fp32 / bf16 -> token-major rows [N, H*W, C] fp32, bit-exact (a widening copy)."""
    from ddf_b200.ops import fused
    torch.manual_seed(N * C + H)
    x = torch.randn(N, C, H, W, device="cuda").to(dtype)
    cam = fused.nchw_to_rows(x)
    assert cam.rows.dtype == torch.float32 and tuple(cam.shape) == (N, C, H, W)
    assert torch.equal(cam.rows, x.float().flatten(2).transpose(1, 2))
    assert torch.equal(cam.nchw(), x.float())
    # the transformer's flatten(2).transpose(1, 2) of the NCHW view is the rows again, without a copy
    back = cam.nchw().flatten(2).transpose(1, 2)
    assert back.is_contiguous() and back.data_ptr() == cam.rows.data_ptr()


@pytest.mark.parametrize("N,L,C", [(1, 1, 128), (2, 37, 128), (12, 12168, 128), (3, 22400, 128), (5, 4099, 64), (2, 333, 256)])
def test_group_norm_rows_matches_torch_group_norm(N, L, C):
    """GroupNorm(C / 4 groups) on token-major [N, L, C] against torch.nn.GroupNorm on the transposed (N, C, L) tensor in
    float64 (actr.py:150-158, 172-187: Conv(k=1) + GroupNorm(32, 128)); forward 1e-5, gradients 1e-4 of the largest value."""
    from ddf_b200.ops import fused
    torch.manual_seed(L + C)
    gn = nn.GroupNorm(C // 4, C).cuda()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.normal_()
    x = (torch.randn(N, L, C, device="cuda") * 2 + 0.5).requires_grad_()
    go = torch.randn(N, L, C, device="cuda")
    assert fused.group_norm_rows_ok(gn, C)
    y = fused.group_norm_rows(gn, x)
    y.backward(go)
    ref_gn = copy.deepcopy(gn).double().cpu()
    ref_gn.zero_grad()
    xr = x.detach().double().cpu().requires_grad_()
    yr = ref_gn(xr.transpose(1, 2)).transpose(1, 2)
    yr.backward(go.double().cpu())
    assert rel(y.detach(), yr.detach()) < 1e-5
    assert rel(x.grad, xr.grad) < 1e-4
    assert rel(gn.weight.grad, ref_gn.weight.grad) < 1e-4
    assert rel(gn.bias.grad, ref_gn.bias.grad) < 1e-4


def test_group_norm_rows_falls_back_when_groups_are_not_four_wide():
    from ddf_b200.ops import fused
    gn = nn.GroupNorm(32, 64).cuda()          # 2 channels per group: the transposing path
    x = torch.randn(2, 50, 64, device="cuda")
    assert not fused.group_norm_rows_ok(gn, 64)
    assert torch.allclose(fused.group_norm_rows(gn, x), gn(x.transpose(1, 2)).transpose(1, 2))


def test_actr_projects_camera_rows_like_the_module_chain():
    """ACTR._project_map (rows: GEMM + GroupNorm on rows, result handed on as an NCHW view) against
    input_proj = Conv2d(k=1) + GroupNorm on the NCHW map, outputs and parameter gradients; fp32 and bf16 maps."""
    from ddf_b200.fusion import actr as actr_mod
    from ddf_b200.ops import fused
    torch.manual_seed(5)
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        proj = nn.Sequential(nn.Conv2d(256, 128, kernel_size=1), nn.GroupNorm(32, 128)).cuda()

        class Holder:
            input_proj = [proj]
        for dtype in (torch.float32, torch.bfloat16):
            x = torch.randn(4, 256, 40, 50, device="cuda").to(dtype)
            go = torch.randn(4, 128, 40, 50, device="cuda")
            proj.zero_grad()
            want = proj(x.float())
            want.backward(go)
            gw, gb, ggw = proj[0].weight.grad.clone(), proj[0].bias.grad.clone(), proj[1].weight.grad.clone()
            proj.zero_grad()
            got = actr_mod._project_map(Holder, 0, x)
            assert got.shape == want.shape
            got.backward(go)
            assert rel(got.detach(), want.detach().cpu()) < 1e-5
            assert rel(proj[0].weight.grad, gw.cpu()) < 2e-3          # xty reads tf32 operands
            assert rel(proj[0].bias.grad, gb.cpu()) < 1e-4
            assert rel(proj[1].weight.grad, ggw.cpu()) < 1e-4
            # a wrapper that already holds the rows passes them in
            got2 = actr_mod._project_map(Holder, 0, fused.nchw_to_rows(x))
            assert torch.equal(got2, got)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def test_nchw_to_rows_carries_the_gradient():
    """Maps that require grad (CenterPoint's gated camera features) go through the same kernel; the backward is the
    transpose back, bit-exact."""
    from ddf_b200.ops import fused
    torch.manual_seed(9)
    x = torch.randn(3, 40, 17, 23, device="cuda", requires_grad=True)
    cam = fused.nchw_to_rows(x)
    assert cam.rows.requires_grad and torch.equal(cam.rows.detach(), x.detach().flatten(2).transpose(1, 2))
    g = torch.randn_like(cam.rows)
    cam.rows.backward(g)
    assert torch.equal(x.grad, g.transpose(1, 2).reshape(x.shape))
