"""Fused BatchNorm1d (+ residual) (+ ReLU) over sparse features against the reference's module chain
nn.BatchNorm1d -> (+ identity) -> ReLU (TransFusion/mmdet3d/ops/sparse_block.py:102-120) evaluated in
float64 on the CPU. fp32 tolerance 1e-5 relative to the largest value (outputs), 1e-4 (gradients and
running statistics), far inside the 1e-3 north_star bar."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def chain(bn, x, res, relu, mask=None):
    """Reference module chain. ReLU is discontinuous: an element whose pre-activation is within fp32
    rounding of 0 may take the other branch on the device, which changes every gradient it feeds; with
    ``mask`` the reference takes the device's branch for those (|pre| < 1e-5) elements."""
    y = bn(x)
    if res is not None:
        y = y + res
    if not relu:
        return y
    if mask is None:
        return torch.relu(y)
    near = y.detach().abs() < 1e-5
    keep = torch.where(near, mask, y.detach() > 0)
    return y * keep.to(y.dtype)


@pytest.mark.parametrize("C", [4, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("n", [2, 257, 40001])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (False, False), (True, False)])
def test_batch_norm_act_matches_module_chain(C, n, training, with_res, relu):
    from ddf_b200.ops.sparse_norm import batch_norm_act
    g = torch.Generator().manual_seed(C * 1000 + n)
    x = torch.randn(n, C, generator=g) * 3 + torch.randn(1, C, generator=g) * 2
    res = torch.randn(n, C, generator=g) if with_res else None
    go = torch.randn(n, C, generator=g)
    bn = nn.BatchNorm1d(C, eps=1e-3, momentum=0.01)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.normal_(generator=g)
        bn.running_mean.normal_(generator=g)
        bn.running_var.uniform_(0.5, 2.0, generator=g)
    bn.train(training)
    ref_bn = copy.deepcopy(bn).double()
    xr = x.double().requires_grad_()
    rr = res.double().requires_grad_() if with_res else None
    gtol = 1e-4 if n >= 10 else 2e-3   # n = 2: xhat = +-1, the backward is a difference of equal terms

    dev_bn = copy.deepcopy(bn).cuda()
    xd = x.cuda().requires_grad_()
    rd = res.cuda().requires_grad_() if with_res else None
    out = batch_norm_act(dev_bn, xd, rd, relu)
    out.backward(go.cuda())
    ref = chain(ref_bn, xr, rr, relu, mask=(out.detach() > 0).cpu())
    ref.backward(go.double())
    assert rel(out.detach(), ref.detach()) < 1e-5
    assert rel(xd.grad, xr.grad) < gtol
    if with_res:
        assert rel(rd.grad, rr.grad) < 1e-6
    assert rel(dev_bn.weight.grad, ref_bn.weight.grad) < max(gtol, 1e-4)
    assert rel(dev_bn.bias.grad, ref_bn.bias.grad) < max(gtol, 1e-4)
    assert rel(dev_bn.running_mean, ref_bn.running_mean) < 1e-5
    assert rel(dev_bn.running_var, ref_bn.running_var) < 1e-5
    assert int(dev_bn.num_batches_tracked) == int(ref_bn.num_batches_tracked)


def test_workspace_is_reusable_back_to_back():
    from ddf_b200.ops.sparse_norm import batch_norm_act
    torch.manual_seed(0)
    bn = nn.BatchNorm1d(32, eps=1e-3, momentum=0.01).cuda().train()
    x = torch.randn(100000, 32, device="cuda")
    outs = [batch_norm_act(bn, x, None, True) for _ in range(5)]
    ref = torch.relu(torch.nn.functional.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.0, 1e-3))
    for o in outs:
        assert rel(o.detach(), ref.detach().cpu()) < 1e-5


def test_unsupported_width_uses_library_modules_and_cpu_is_refused():
    from ddf_b200.ops.sparse_norm import batch_norm_act
    bn = nn.BatchNorm1d(24).cuda().train()
    x = torch.randn(50, 24, device="cuda")
    y = batch_norm_act(bn, x, None, True)
    assert y.shape == x.shape and float(y.min()) >= 0
    with pytest.raises(RuntimeError):
        batch_norm_act(nn.BatchNorm1d(16), torch.randn(8, 16), None, True)


@pytest.mark.parametrize("C,with_res,relu", [(32, False, True), (64, True, True), (128, False, False)])
def test_bn_forward_split_writes_conv_operands(C, with_res, relu):
    """ddf_sparse_bn_forward_split: y is unchanged, the operand copies equal ddf_split_bf16x3(y) bit for bit, the
    conv Function picks them up (same conv output as from the separate split pass) and backward is untouched."""
    import ddf_b200.ops.spconv as sp
    from ddf_b200.ops import sparse_norm
    from ddf_b200.ops.spconv import ops, functional
    from ddf_b200 import lib as _lib
    torch.manual_seed(C)
    n = 20011
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    x = torch.randn(n, C, device="cuda", requires_grad=True)
    res = torch.randn(n, C, device="cuda") if with_res else None
    prev = _lib.get_lib().ddf_set_tensor_cores(4)
    try:
        y = sparse_norm.batch_norm_act(bn, x, res, relu)
        assert hasattr(y, "_ddf_operands")
        split, rounded = y._ddf_operands
        ref_split, ref_rounded = ops.split_bf16x3(y.detach(), want_rounded=True)
        assert torch.equal(split.view(torch.int32), ref_split.view(torch.int32))
        assert torch.equal(rounded, ref_rounded)
        _lib.get_lib().ddf_set_tensor_cores(0)
        bn2 = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
        bn2.load_state_dict({k: v.clone() for k, v in bn.state_dict().items() if k != "num_batches_tracked"}, strict=False)
        with torch.no_grad():
            bn2.running_mean.zero_(); bn2.running_var.fill_(1.0)
        x2 = x.detach().clone().requires_grad_()
        y2 = sparse_norm.batch_norm_act(bn2, x2, res, relu)
        assert not hasattr(y2, "_ddf_operands")
        assert torch.equal(y.detach(), y2.detach())
        go = torch.randn_like(y)
        y.backward(go)
        y2.backward(go)
        assert torch.equal(x.grad, x2.grad)
    finally:
        _lib.get_lib().ddf_set_tensor_cores(prev)
