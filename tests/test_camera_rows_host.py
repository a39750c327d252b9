"""Host side of the token-major camera path (ops/fused.py: CameraRows, nchw_to_rows, group_norm_rows;
fusion/actr.py: _project_map) on the CPU: without a GPU every piece takes the reference's module chain
(actr.py:150-158, 172-187), so the results must equal it exactly, gradients included. The kernels themselves are
checked on the GPU (tests/test_fused_gpu.py)."""
import torch
from torch import nn


def test_nchw_to_rows_cpu_is_a_view_chain_with_gradient():
    from ddf_b200.ops import fused
    torch.manual_seed(0)
    x = torch.randn(2, 6, 3, 5, requires_grad=True)
    cam = fused.nchw_to_rows(x)
    assert tuple(cam.shape) == (2, 6, 3, 5) and cam.rows.shape == (2, 15, 6)
    assert torch.equal(cam.rows, x.flatten(2).transpose(1, 2))
    assert torch.equal(cam.nchw(), x)
    back = cam.nchw().flatten(2).transpose(1, 2)
    assert back.data_ptr() == cam.rows.data_ptr() and back.is_contiguous()     # the transformer's flatten is free
    g = torch.randn_like(cam.rows)
    cam.rows.backward(g)
    assert torch.equal(x.grad, g.transpose(1, 2).reshape(x.shape))
    # bf16 maps are widened
    assert fused.nchw_to_rows(x.detach().to(torch.bfloat16)).rows.dtype == torch.float32


def test_group_norm_rows_cpu_is_group_norm_between_transposes():
    from ddf_b200.ops import fused
    torch.manual_seed(1)
    gn = nn.GroupNorm(32, 128)
    x = torch.randn(3, 17, 128)
    assert torch.equal(fused.group_norm_rows(gn, x), gn(x.transpose(1, 2)).transpose(1, 2))


def test_project_map_cpu_equals_input_proj():
    from ddf_b200.fusion import actr as actr_mod
    from ddf_b200.ops import fused
    torch.manual_seed(2)
    proj = nn.Sequential(nn.Conv2d(16, 128, kernel_size=1), nn.GroupNorm(32, 128))

    class Holder:
        input_proj = [proj]
    x = torch.randn(2, 16, 5, 7)
    want = proj(x)
    assert torch.equal(actr_mod._project_map(Holder, 0, x), want)
    # a wrapper may hand the rows over; off the fast path they are viewed back as NCHW
    got = actr_mod._project_map(Holder, 0, fused.nchw_to_rows(x))
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-6)


def test_wrapper_gather_from_rows_wraps_negative_pixels_like_advanced_indexing():
    """point_fusion.py:375-378 indexes img_feats[0][row, :, iy, ix]; negative iy / ix wrap. The row gather must too."""
    from ddf_b200.ops import fused
    torch.manual_seed(3)
    img = torch.randn(4, 8, 6, 9)
    cam = fused.nchw_to_rows(img)
    row = torch.tensor([0, 3, 2, 1, 3])
    iy = torch.tensor([0, -1, 5, -6, 2])
    ix = torch.tensor([8, -9, -1, 0, 4])
    want = img[row, :, iy, ix]
    flat = (row * cam.H + iy.remainder(cam.H)) * cam.W + ix.remainder(cam.W)
    got = cam.rows.view(-1, 8).index_select(0, flat)
    assert torch.equal(got, want)
