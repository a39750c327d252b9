"""Oracle pinning (CPU): oracle/spconv_ref.c against the reference's own sparse_conv_ext built
unmodified into oracle/_ref (CPU path), and against dense torch conv3d as an independent check."""
import numpy as np
import pytest
import torch

from oracle import ref_build, spconv as osp


def random_voxels(n, batch, shape, seed):
    rng = np.random.default_rng(seed)
    cells = batch * int(np.prod(shape))
    flat = rng.choice(cells, size=min(n, cells), replace=False)
    idx = np.stack(np.unravel_index(flat, (batch, *shape)), 1).astype(np.int32)
    return idx


GEOMS = [
    # (spatial_shape, ksize, stride, padding, dilation, subm)  — the SparseEncoder layer shapes
    ([11, 40, 40], [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True),
    ([11, 40, 40], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False),
    ([11, 40, 40], [3, 3, 3], [2, 2, 2], [0, 1, 1], [1, 1, 1], False),
    ([5, 24, 24], [3, 1, 1], [2, 1, 1], [0, 0, 0], [1, 1, 1], False),
    ([9, 17, 13], [3, 3, 3], [1, 1, 1], [0, 0, 0], [2, 2, 2], True),   # dilated SubM
    ([9, 17, 13], [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], False),  # stride-1 regular conv
    ([8, 16, 16], [2, 2, 2], [2, 2, 2], [0, 0, 0], [1, 1, 1], False),
]


@pytest.mark.parametrize("geom", GEOMS)
def test_rulebook_equals_reference_extension(geom):
    ext = ref_build.load("sparse_conv_ext")
    if ext is None:
        pytest.skip("oracle/_ref/sparse_conv_ext.so not built (needs /root/reference)")
    shape, ks, st, pad, dil, subm = geom
    idx = random_voxels(700, 2, shape, seed=3)
    outids, pairs, num, out_shape = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm)
    r_out, r_pairs, r_num = ext.get_indice_pairs_3d(torch.from_numpy(idx), 2, out_shape, shape, ks, st, pad,
                                                    dil, [0, 0, 0], int(subm), 0)
    assert np.array_equal(r_num.numpy(), num)
    assert np.array_equal(r_out.numpy(), outids)
    assert np.array_equal(r_pairs.numpy(), pairs)


@pytest.mark.parametrize("geom", GEOMS[:4])
def test_conv_fwd_bwd_equals_reference_extension(geom):
    ext = ref_build.load("sparse_conv_ext")
    if ext is None:
        pytest.skip("oracle/_ref/sparse_conv_ext.so not built (needs /root/reference)")
    shape, ks, st, pad, dil, subm = geom
    rng = np.random.default_rng(5)
    idx = random_voxels(500, 2, shape, seed=4)
    outids, pairs, num, out_shape = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm)
    cin, cout = 16, 32
    feat = rng.standard_normal((len(idx), cin)).astype(np.float32)
    w = rng.standard_normal((*ks, cin, cout)).astype(np.float32)
    out = osp.indice_conv(feat, w, pairs, num, len(outids))
    r = ext.indice_conv_fp32(torch.from_numpy(feat), torch.from_numpy(w), torch.from_numpy(pairs),
                             torch.from_numpy(num), len(outids), 0, int(subm))
    np.testing.assert_allclose(out, r.numpy(), rtol=1e-4, atol=1e-4)
    go = rng.standard_normal(out.shape).astype(np.float32)
    gin, gw = osp.indice_conv_backward(feat, w, go, pairs, num)
    rgi, rgw = ext.indice_conv_backward_fp32(torch.from_numpy(feat), torch.from_numpy(w), torch.from_numpy(go),
                                             torch.from_numpy(pairs), torch.from_numpy(num), 0, int(subm))
    np.testing.assert_allclose(gin, rgi.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gw, rgw.numpy(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_equals_dense_conv3d(geom):
    """Derived known-answer: densify, torch conv3d with the weight permuted
    [kd,kh,kw,Cin,Cout] -> [Cout,Cin,kd,kh,kw], compare on the active outputs."""
    shape, ks, st, pad, dil, subm = geom
    rng = np.random.default_rng(7)
    idx = random_voxels(400, 2, shape, seed=6)
    outids, pairs, num, out_shape = osp.get_indice_pairs(idx, 2, shape, ks, st, pad, dil, subm, order="gpu")
    cin, cout = 5, 7
    feat = rng.standard_normal((len(idx), cin)).astype(np.float32)
    w = rng.standard_normal((*ks, cin, cout)).astype(np.float32)
    out = osp.indice_conv(feat, w, pairs, num, len(outids))
    dense_in = torch.from_numpy(osp.dense(feat, idx, shape, 2))
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous()
    p = [k // 2 for k in ks] if subm else pad
    s = [1, 1, 1] if subm else st
    if subm and dil != [1, 1, 1]:
        p = [d * (k // 2) for d, k in zip(dil, ks)]
        # the reference keeps padding = k/2 even when dilated (spconv_ops.h:76-79): emulate by shifting
        pytest.skip("dilated SubM keeps pad=k/2 (asymmetric window); covered by the extension test")
    dense_out = torch.nn.functional.conv3d(dense_in, wt, stride=s, padding=p, dilation=dil).numpy()
    got = dense_out[outids[:, 0], :, outids[:, 1], outids[:, 2], outids[:, 3]]
    np.testing.assert_allclose(out, got, rtol=1e-4, atol=1e-4)
    if not subm:
        # regular conv: every non-zero dense output cell is an active output (and sorted order)
        flat = ((outids[:, 0].astype(np.int64) * out_shape[0] + outids[:, 1]) * out_shape[1] + outids[:, 2]) * out_shape[2] + outids[:, 3]
        assert np.all(np.diff(flat) > 0)
