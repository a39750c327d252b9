"""Fused elementwise ops of the 3D-DF encoder layers (``ddf_bias_relu_dropout_*``,
``ddf_add_dropout_layer_norm_*``, include/ddf_b200.h) behind helpers that take the reference's own
modules (``nn.Linear``, ``nn.Dropout``, ``nn.LayerNorm``: parameters and state-dict keys untouched):

    ffn_hidden(linear1, dropout, x)            == dropout(relu(linear1(x)))
    add_dropout_layer_norm(norm, dropout, a, b) == norm(a + dropout(b))

(<proj>/models/model_utils/actr_transformer.py:383-397).
"""
import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib as _lib


def _seed():
    # drawn from the CPU generator, so torch.manual_seed makes the dropout pattern reproducible
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class _BiasReluDropout(Function):
    @staticmethod
    def forward(ctx, h, bias, p):
        _lib.require_cuda(h, bias)
        C = h.shape[-1]
        n = h.numel() // C
        seed = _seed() if p > 0 else 0
        with _lib.on_device(h.device):
            rc = _lib.get_lib().ddf_bias_relu_dropout_forward(_lib.ptr(h), _lib.ptr(bias), _lib.ptr(h), n, C,
                                                              float(p), seed, _lib.current_stream())
        _lib.check(rc, "bias_relu_dropout_forward")
        ctx.p = float(p)
        ctx.has_bias = bias is not None
        ctx.mark_dirty(h)
        ctx.save_for_backward(h)
        return h

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gh = torch.empty_like(out)
        C = out.shape[-1]
        want_gb = ctx.has_bias and ctx.needs_input_grad[1]
        in_kernel = want_gb and (C // 4) <= 256 and 256 % (C // 4) == 0
        gb = torch.zeros(C, dtype=torch.float32, device=out.device) if in_kernel else None
        with _lib.on_device(out.device):
            rc = _lib.get_lib().ddf_bias_relu_dropout_backward(_lib.ptr(grad_out), _lib.ptr(out), _lib.ptr(gh),
                                                               _lib.ptr(gb), out.numel() // C, C, ctx.p,
                                                               _lib.current_stream())
        _lib.check(rc, "bias_relu_dropout_backward")
        if want_gb and not in_kernel:
            gb = gh.reshape(-1, C).sum(0)
        return gh, gb, None


def col_sum(x2d):
    """Column sums of a contiguous fp32 (rows, C) CUDA matrix (ddf_col_sum)."""
    out = torch.empty(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    with _lib.on_device(x2d.device):
        rc = _lib.get_lib().ddf_col_sum(_lib.ptr(x2d), _lib.ptr(out), x2d.shape[0], x2d.shape[1], _lib.current_stream())
    _lib.check(rc, "col_sum")
    return out


def _col_sum_ok(C):
    return C % 4 == 0 and C // 4 <= 256


def _bias_in_kernel_ok(C):
    # relu_dropout_bwd_bias_kernel: a thread row of the CTA per C / 4 columns
    return C % 4 == 0 and C // 4 <= 256 and 256 % (C // 4) == 0


def xty(a2d, b2d):
    """a2d (K, M)^T @ b2d (K, N) -> (M, N): the weight gradient of a Linear over K tokens.  tf32 arithmetic allowed
    (``torch.backends.cuda.matmul.allow_tf32``) and widths that are multiples of 32: one split-K tcgen05 pass that
    streams both operands once (ddf_xty_tf32); otherwise the library product."""
    K, M = a2d.shape
    N = b2d.shape[1]
    if (torch.backends.cuda.matmul.allow_tf32 and K >= 4096 and a2d.is_cuda and a2d.dtype == torch.float32
            and b2d.dtype == torch.float32 and a2d.is_contiguous() and b2d.is_contiguous()
            and a2d.data_ptr() % 16 == 0 and b2d.data_ptr() % 16 == 0
            and _lib.get_lib().ddf_xty_supported(K, M, N)):
        out = torch.empty((M, N), dtype=torch.float32, device=a2d.device)
        with _lib.on_device(a2d.device):
            rc = _lib.get_lib().ddf_xty_tf32(_lib.ptr(a2d), _lib.ptr(b2d), _lib.ptr(out), K, M, N, _lib.current_stream())
        _lib.check(rc, "xty_tf32")
        return out
    return a2d.t() @ b2d


class _Linear(Function):
    """F.linear over (..., Cin) tokens whose bias gradient is one pass at the HBM rate (ddf_col_sum) instead of
    ATen's grad.sum(0) (168 us for a [146 k, 128] gradient on B200 = 14x below the copy rate) and whose weight
    gradient is the split-K tcgen05 pass ``xty``; the forward and input-gradient GEMMs stay in the library."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return F.linear(x, weight, bias)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            gw = xty(g2, x.reshape(-1, x.shape[-1]))
        if ctx.needs_input_grad[2]:
            gb = col_sum(g2) if _col_sum_ok(g2.shape[1]) else g2.sum(0)
        return gx, gw, gb


class _LinearNoBias(Function):
    """x @ weight.T over (..., Cin) tokens; the weight gradient through ``xty``."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return F.linear(x, weight)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = (g2 @ weight).view(x.shape) if ctx.needs_input_grad[0] else None
        gw = xty(g2, x.reshape(-1, x.shape[-1]).contiguous()) if ctx.needs_input_grad[1] else None
        return gx, gw


def linear(module, x):
    """``module(x)`` for an ``nn.Linear`` applied to many tokens (same parameters, same forward GEMM)."""
    if (module.bias is None or not x.is_cuda or x.dtype != torch.float32 or module.weight.dtype != torch.float32
            or not torch.is_grad_enabled() or x.numel() // max(x.shape[-1], 1) < 4096):
        return module(x)
    return _Linear.apply(x, module.weight, module.bias)


def linear_wb(x, weight, bias):
    """``F.linear(x, weight, bias)`` with the fast bias gradient (weight / bias given directly)."""
    if (not x.is_cuda or x.dtype != torch.float32 or weight.dtype != torch.float32
            or not torch.is_grad_enabled() or x.numel() // max(x.shape[-1], 1) < 4096):
        return F.linear(x, weight, bias)
    if bias is None:
        return _LinearNoBias.apply(x, weight)
    return _Linear.apply(x, weight, bias)


def ffn_hidden(linear, dropout, x):
    """``dropout(relu(linear(x)))``: the GEMM without bias, then bias + ReLU + dropout in one in-place pass."""
    C = linear.out_features
    if x.dtype != torch.float32 or C % 4:
        _lib.require_cuda(x)
        return dropout(F.relu(linear(x)))
    big = x.is_cuda and torch.is_grad_enabled() and x.numel() // max(x.shape[-1], 1) >= 4096
    h = _LinearNoBias.apply(x, linear.weight) if big else F.linear(x, linear.weight)   # fresh tensor: overwritten in place
    if h.dtype != torch.float32 or (linear.bias is not None and linear.bias.dtype != torch.float32):
        # e.g. under torch.autocast the GEMM returns half precision: the fp32 kernels must not see it
        h = h if linear.bias is None else h + linear.bias.to(h.dtype)
        return dropout(F.relu(h))
    p = dropout.p if dropout.training else 0.0
    return _BiasReluDropout.apply(h.contiguous(), linear.bias, p)


class _AddDropoutLayerNorm(Function):
    @staticmethod
    def forward(ctx, a, b, gamma, beta, p, eps):
        _lib.require_cuda(a, b, gamma, beta)
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        C = a.shape[-1]
        rows = a.numel() // C
        s = torch.empty_like(a)
        y = torch.empty_like(a)
        mean = torch.empty(rows, dtype=torch.float32, device=a.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=a.device)
        seed = _seed() if (p > 0 and b is not None) else 0
        with _lib.on_device(a.device):
            rc = _lib.get_lib().ddf_add_dropout_layer_norm_forward(
                _lib.ptr(a), _lib.ptr(b), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(s), _lib.ptr(y), _lib.ptr(mean),
                _lib.ptr(rstd), rows, C, float(p), seed, float(eps), _lib.current_stream())
        _lib.check(rc, "add_dropout_layer_norm_forward")
        ctx.p, ctx.seed, ctx.has_b = float(p), seed, b is not None
        ctx.save_for_backward(s, gamma, mean, rstd)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        s, gamma, mean, rstd = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        C = s.shape[-1]
        rows = s.numel() // C
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_b
        ga = torch.empty_like(s) if (need_a or (need_b and ctx.p == 0)) else None
        share = need_b and ctx.p == 0          # without dropout both branches get the same gradient
        gb = torch.empty_like(s) if (need_b and not share) else None
        gg = torch.zeros(C, dtype=torch.float32, device=s.device)
        gbeta = torch.zeros(C, dtype=torch.float32, device=s.device)
        with _lib.on_device(s.device):
            rc = _lib.get_lib().ddf_add_dropout_layer_norm_backward(
                _lib.ptr(grad_y), _lib.ptr(s), _lib.ptr(gamma), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(ga),
                _lib.ptr(gb), _lib.ptr(gg), _lib.ptr(gbeta), rows, C, ctx.p, ctx.seed, _lib.current_stream())
        _lib.check(rc, "add_dropout_layer_norm_backward")
        return (ga if need_a else None), (ga if share else gb), gg, gbeta, None, None


def add_dropout_layer_norm(norm, dropout, a, b):
    """``norm(a + dropout(b))`` with ``norm`` an ``nn.LayerNorm`` over the last dim."""
    C = a.shape[-1]
    if (a.dtype != torch.float32 or C not in (32, 64, 128, 256, 512) or not norm.elementwise_affine
            or norm.bias is None or tuple(norm.normalized_shape) != (C,)
            or (b is not None and (b.dtype != torch.float32 or b.shape != a.shape))
            or norm.weight.dtype != torch.float32 or norm.bias.dtype != torch.float32):
        _lib.require_cuda(a)
        if b is None:
            return norm(a)
        return norm(a + (dropout(b) if dropout is not None else b))
    p = dropout.p if (dropout is not None and dropout.training) else 0.0
    return _AddDropoutLayerNorm.apply(a, b, norm.weight, norm.bias, p, norm.eps)


class _BiGateSum(Function):
    """o1 = f1 + f2 * sigmoid(wb . u1 + bb), o2 = f2 + f1 * sigmoid(wa . u2 + ba) in one pass each way
    (ddf_bigate_sum_forward / _backward); u1 = u2 = f1 + f2 when ``fuse_in`` else u1 = f1, u2 = f2."""

    @staticmethod
    def forward(ctx, f1, f2, wb, bb, wa, ba, fuse_in):
        _lib.require_cuda(f1, f2, wb, wa)
        f1, f2 = f1.contiguous(), f2.contiguous()
        C = f1.shape[-1]
        rows = f1.numel() // C
        wbv, wav = wb.reshape(-1).contiguous(), wa.reshape(-1).contiguous()
        o1, o2 = torch.empty_like(f1), torch.empty_like(f2)
        gates = torch.empty((rows, 2), dtype=torch.float32, device=f1.device)
        with _lib.on_device(f1.device):
            rc = _lib.get_lib().ddf_bigate_sum_forward(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(wbv), _lib.ptr(bb),
                                                       _lib.ptr(wav), _lib.ptr(ba), _lib.ptr(o1), _lib.ptr(o2),
                                                       _lib.ptr(gates), rows, C, int(fuse_in), _lib.current_stream())
        _lib.check(rc, "bigate_sum_forward")
        ctx.fuse_in = bool(fuse_in)
        ctx.wshape = (wb.shape, wa.shape)
        ctx.has_bias = (bb is not None, ba is not None)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(f1, f2, gates, wbv, wav)
        return o1, o2

    @staticmethod
    @once_differentiable
    def backward(ctx, go1, go2):
        f1, f2, gates, wbv, wav = ctx.saved_tensors
        C = f1.shape[-1]
        rows = f1.numel() // C
        go1 = go1.contiguous() if go1 is not None else None
        go2 = go2.contiguous() if go2 is not None else None
        need = ctx.needs_input_grad
        new = lambda n: torch.empty(n, dtype=torch.float32, device=f1.device)
        gf1 = torch.empty_like(f1) if need[0] else None
        gf2 = torch.empty_like(f2) if need[1] else None
        # a gate only feeds its own output: with that output unused its parameters get no gradient at all (None, as
        # autograd leaves them in the module chain - the last layer's a_conv1d is structurally unused)
        gwb = new(C) if (need[2] and go1 is not None) else None
        gbb = new(1) if (need[3] and ctx.has_bias[0] and go1 is not None) else None
        gwa = new(C) if (need[4] and go2 is not None) else None
        gba = new(1) if (need[5] and ctx.has_bias[1] and go2 is not None) else None
        with _lib.on_device(f1.device):
            rc = _lib.get_lib().ddf_bigate_sum_backward(
                _lib.ptr(go1), _lib.ptr(go2), _lib.ptr(f1), _lib.ptr(f2), _lib.ptr(gates), _lib.ptr(wbv), _lib.ptr(wav),
                _lib.ptr(gf1), _lib.ptr(gf2), _lib.ptr(gwb), _lib.ptr(gbb), _lib.ptr(gwa), _lib.ptr(gba), rows, C,
                int(ctx.fuse_in), _lib.current_stream())
        _lib.check(rc, "bigate_sum_backward")
        return (gf1, gf2, gwb.view(ctx.wshape[0]) if gwb is not None else None, gbb,
                gwa.view(ctx.wshape[1]) if gwa is not None else None, gba, None)


def bigate_sum(b_conv, a_conv, feat1, feat2, fuse_in):
    """BiGateSum1D (``fuse_in=False``) / BiGateSum1D_2 (``True``) on (..., C) token rows with the module's own
    ``Conv1d(C, 1, 1)`` gates; ``None`` when the shape / dtype is not the fused kernel's (the caller then runs the
    module chain)."""
    C = feat1.shape[-1]
    if (not feat1.is_cuda or feat1.dtype != torch.float32 or feat2.dtype != torch.float32 or feat1.shape != feat2.shape
            or C not in (128, 256) or b_conv.weight.dtype != torch.float32 or a_conv.weight.dtype != torch.float32
            or b_conv.weight.numel() != C or a_conv.weight.numel() != C):
        return None
    return _BiGateSum.apply(feat1, feat2, b_conv.weight, b_conv.bias, a_conv.weight, a_conv.bias, fuse_in)


class CameraRows:
    """A camera feature map kept token-major: ``rows`` [N, H*W, C] fp32 (what ``nchw_to_rows`` returns) with its
    (H, W). ACTR accepts it in place of the NCHW tensor of the same map, so a wrapper that already needs the rows (the
    per-query camera feature is a row gather) converts once."""

    def __init__(self, rows, H, W):
        self.rows, self.H, self.W = rows, int(H), int(W)

    @property
    def shape(self):            # the NCHW shape of the map this stands for
        n, _, c = self.rows.shape
        return torch.Size((n, c, self.H, self.W))

    def nchw(self):
        n, _, c = self.rows.shape
        return self.rows.view(n, self.H, self.W, c).permute(0, 3, 1, 2)


def _transpose_last2(x3, dtype_code):
    """[N, A, B] -> [N, B, A] fp32 through ddf_nchw_to_rows (a batched 2-D transpose; widens bf16)."""
    N, A, B = x3.shape
    out = torch.empty((N, B, A), dtype=torch.float32, device=x3.device)
    with _lib.on_device(x3.device):
        rc = _lib.get_lib().ddf_nchw_to_rows(_lib.ptr(x3), dtype_code, _lib.ptr(out), N, A, B, _lib.current_stream())
    _lib.check(rc, "nchw_to_rows")
    return out


class _NchwToRows(Function):
    """(N, C, H, W) -> rows [N, H*W, C] for maps that carry a gradient (CenterPoint's gated camera features): the
    backward is the same transposing kernel the other way round."""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        ctx.hw = (H, W)
        return _transpose_last2(x.view(N, C, H * W), 0)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        N, HW, C = g.shape
        if HW > 65535 * 32:      # the kernel's grid.y covers the source's middle dimension in tiles of 32
            return g.transpose(1, 2).reshape(N, C, *ctx.hw)
        return _transpose_last2(g.contiguous(), 0).view(N, C, *ctx.hw)


def nchw_to_rows(x):
    """(N, C, H, W) fp32 / bf16 -> CameraRows with rows [N, H*W, C] fp32 in one transposing pass (ddf_nchw_to_rows)."""
    N, C, H, W = x.shape
    if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16) or N >= 65536 or not x.is_contiguous():
        return CameraRows(x.flatten(2).transpose(1, 2).float().contiguous(), H, W)
    if x.requires_grad and torch.is_grad_enabled():
        if x.dtype != torch.float32:
            return CameraRows(x.flatten(2).transpose(1, 2).float().contiguous(), H, W)
        return CameraRows(_NchwToRows.apply(x), H, W)
    return CameraRows(_transpose_last2(x.detach().view(N, C, H * W), 1 if x.dtype == torch.bfloat16 else 0), H, W)


class _GroupNormRows(Function):
    """torch.nn.GroupNorm over token-major data x [N, L, C] (statistics over the (L, C / G) elements of a (sample,
    group)) without the transposes to and from (N, C, L): ddf_group_norm_rows_forward / _backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, G, eps):
        _lib.require_cuda(x, weight, bias)
        x = x.contiguous()
        N, L, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty((N, G), dtype=torch.float32, device=x.device)
        rstd = torch.empty((N, G), dtype=torch.float32, device=x.device)
        ws = torch.empty(2 * N * G, dtype=torch.float64, device=x.device)
        with _lib.on_device(x.device):
            rc = _lib.get_lib().ddf_group_norm_rows_forward(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y),
                                                            _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(ws), N, L, C, G,
                                                            float(eps), _lib.current_stream())
        _lib.check(rc, "group_norm_rows_forward")
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.G = G
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, weight, mean, rstd = ctx.saved_tensors
        N, L, C = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        need = ctx.needs_input_grad
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if need[1] else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if need[2] else None
        ws = torch.empty(2 * N * ctx.G, dtype=torch.float64, device=x.device)
        with _lib.on_device(x.device):
            rc = _lib.get_lib().ddf_group_norm_rows_backward(_lib.ptr(gy), _lib.ptr(x), _lib.ptr(weight), _lib.ptr(mean),
                                                             _lib.ptr(rstd), _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb),
                                                             _lib.ptr(ws), N, L, C, ctx.G, _lib.current_stream())
        _lib.check(rc, "group_norm_rows_backward")
        return (gx if need[0] else None), gw, gb, None, None


def group_norm_rows_ok(gn, C, device_is_cuda=True):
    return bool(device_is_cuda and gn.affine and gn.weight.dtype == torch.float32 and gn.num_channels == C
                and _lib.get_lib().ddf_group_norm_rows_supported(C, gn.num_groups))


def group_norm_rows(gn, x):
    """``gn(x.transpose(1, 2)).transpose(1, 2)`` for token-major x [N, L, C] (the reference's GroupNorm around Conv1d /
    Conv2d(k=1) projections, actr.py:150-158): the rows kernel when C == 4 * groups, else the transposes."""
    if x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and group_norm_rows_ok(gn, x.shape[-1]):
        return _GroupNormRows.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
    return gn(x.transpose(1, 2)).transpose(1, 2)


class _FusedFFN(Function):
    """y = linear2(dropout(relu(linear1(x)))) with the forward as ONE kernel (ddf_ffn_forward: the [T, d_ffn] hidden
    activation is written once for backward and never read again in forward). Backward: the split-K weight
    gradients (xty), the fused ReLU / dropout / bias-gradient pass, library GEMMs for the two input gradients."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, p):
        _lib.require_cuda(x, w1, b1, w2, b2)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        T, D = x2.shape
        F_ = w1.shape[0]
        w1c, w2c = w1.contiguous(), w2.contiguous()
        h = torch.empty((T, F_), dtype=torch.float32, device=x.device)
        y = torch.empty((T, D), dtype=torch.float32, device=x.device)
        seed = _seed() if p > 0 else 0
        ws = torch.empty(2 * D * F_, dtype=torch.float32, device=x.device)     # re-laid weights (ddf_ffn_workspace_bytes)
        with _lib.on_device(x.device):
            rc = _lib.get_lib().ddf_ffn_forward(_lib.ptr(x2), _lib.ptr(w1c), _lib.ptr(b1), _lib.ptr(w2c), _lib.ptr(b2),
                                                _lib.ptr(h), _lib.ptr(y), _lib.ptr(ws), T, D, F_, float(p), seed,
                                                _lib.current_stream())
        _lib.check(rc, "ffn_forward")
        # the kernel's drop threshold is quantised to 1 / 256: backward scales by 1 / (1 - the probability it applied)
        ctx.p = float(_lib.get_lib().ddf_ffn_dropout_p(float(p))) if p > 0 else 0.0
        ctx.has_b2 = b2 is not None
        ctx.save_for_backward(x2, h, w1c, w2c)
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, h, w1, w2 = ctx.saved_tensors
        T, D = x2.shape
        F_ = w1.shape[0]
        g2 = gy.reshape(T, D)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        need = ctx.needs_input_grad
        gb2 = (col_sum(g2) if _col_sum_ok(D) else g2.sum(0)) if (need[4] and ctx.has_b2) else None
        gw2 = xty(g2, h) if need[3] else None                         # [D, F]
        gx = gw1 = gb1 = None
        if need[0] or need[1] or need[2]:
            gh = g2 @ w2                                               # [T, F]
            in_kernel = need[2] and _bias_in_kernel_ok(F_)
            gb1 = torch.zeros(F_, dtype=torch.float32, device=h.device) if in_kernel else None
            with _lib.on_device(h.device):
                rc = _lib.get_lib().ddf_bias_relu_dropout_backward(_lib.ptr(gh), _lib.ptr(h), _lib.ptr(gh), _lib.ptr(gb1),
                                                                   T, F_, ctx.p, _lib.current_stream())
            _lib.check(rc, "bias_relu_dropout_backward")
            if need[2] and not in_kernel:
                gb1 = gh.sum(0)
            if need[1]:
                gw1 = xty(gh, x2)                                      # [F, D]
            if need[0]:
                gx = (gh @ w1).view(gy.shape)
        return gx, gw1, gb1, gw2, gb2, None


def ffn(linear1, dropout, linear2, x):
    """``linear2(dropout(relu(linear1(x))))`` over (..., d_model) tokens: the fused forward kernel when the shape is
    its (d_model 128, d_ffn % 64 == 0, fp32, tf32 products allowed, many tokens), else the two-GEMM chain."""
    D, F_ = linear1.in_features, linear1.out_features
    T = x.numel() // max(x.shape[-1], 1)
    if (x.is_cuda and x.dtype == torch.float32 and torch.backends.cuda.matmul.allow_tf32 and T >= 4096
            and linear1.bias is not None and linear1.weight.dtype == torch.float32
            and linear2.weight.dtype == torch.float32 and linear2.in_features == F_ and linear2.out_features == D
            and _lib.get_lib().ddf_ffn_supported(T, D, F_)):
        p = dropout.p if (dropout is not None and dropout.training) else 0.0
        return _FusedFFN.apply(x, linear1.weight, linear1.bias, linear2.weight, linear2.bias, p)
    return linear(linear2, ffn_hidden(linear1, dropout, x))
