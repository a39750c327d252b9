"""Debug driver: one dense SubM conv at bench size through the multi-tile kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from ddf_b200.ops.spconv import ops
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 120
cin = cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = torch.Generator().manual_seed(0)
# dense block of voxels: every interior voxel has all 27 neighbours
zz, yy, xx = torch.meshgrid(torch.arange(8), torch.arange(n_side), torch.arange(n_side), indexing="ij")
idx = torch.stack([torch.zeros_like(zz), zz, yy, xx], -1).reshape(-1, 4).int().cuda().contiguous()
n = idx.shape[0]
print("rows", n, "tiles", (n + 127) // 128, flush=True)
rb = ops.build_rulebook(idx, 1, [8, n_side, n_side], 3, 1, 1, 1, 0, True, False)
feat = ops.round_tf32(torch.randn(n, cin, generator=g).cuda())
w = (torch.randn(3, 3, 3, cin, cout, generator=g) / (27 * cin) ** 0.5).cuda()
out = ops.sparse_conv_forward(feat, w, rb.gather_table, None, n)
torch.cuda.synchronize()
from ddf_b200 import lib
lib.get_lib().ddf_set_tensor_cores(2)
ref = ops.sparse_conv_forward(feat, w, rb.gather_table, None, n)
torch.cuda.synchronize()
print("max diff vs single-tile kernel", float((out - ref).abs().max()), "ref max", float(ref.abs().max()))
