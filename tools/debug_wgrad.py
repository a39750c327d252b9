"""Debug driver: table-driven wgrad on a dense block (bench-like size)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from ddf_b200.ops.spconv import ops
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 200
c = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator().manual_seed(0)
zz, yy, xx = torch.meshgrid(torch.arange(8), torch.arange(n_side), torch.arange(n_side), indexing="ij")
idx = torch.stack([torch.zeros_like(zz), zz, yy, xx], -1).reshape(-1, 4).int().cuda().contiguous()
n = idx.shape[0]
rb = ops.build_rulebook(idx, 1, [8, n_side, n_side], 3, 1, 1, 1, 0, True, False)
feat = ops.round_tf32(torch.randn(n, c, generator=g).cuda())
go = ops.round_tf32(torch.randn(n, c, generator=g).cuda())
w = torch.zeros(3, 3, 3, c, c).cuda()
a = ops.sparse_conv_wgrad_table(feat, w, go, rb.gather_table)
torch.cuda.synchronize()
b = ops.sparse_conv_wgrad(feat, w, go, rb.indice_pairs, rb.indice_pair_num)
torch.cuda.synchronize()
print("rows", n, "max diff table vs pair-list", float((a - b).abs().max()), "ref max", float(b.abs().max()))
