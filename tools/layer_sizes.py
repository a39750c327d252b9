"""Print per-conv sizes (rows, pairs, channels) of the bench workload and time each conv kernel family."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import torch
import bench
from ddf_b200.ops.spconv import ops
model = bench.build_model("cuda")
pts, feats, metas = bench.host_batch(0, 2)
pts = [p.cuda() for p in pts]; feats = feats.cuda()
rows = []
of, od, ow = ops.sparse_conv_forward, ops.sparse_conv_dgrad, ops.sparse_conv_wgrad
def ev():
    return torch.cuda.Event(enable_timing=True)
def fwd(features, filters, table, bias, n_out):
    a, b = ev(), ev(); a.record(); o = of(features, filters, table, bias, n_out); b.record()
    rows.append(["fwd", features.shape[0], n_out, int((table >= 0).sum()), filters.shape[-2], filters.shape[-1], a, b]); return o
def dgrad(filters, gout, table, n_in):
    a, b = ev(), ev(); a.record(); o = od(filters, gout, table, n_in); b.record()
    rows.append(["dgrad", gout.shape[0], n_in, int((table >= 0).sum()), filters.shape[-2], filters.shape[-1], a, b]); return o
def wgrad(features, filters, gout, pairs, num):
    a, b = ev(), ev(); a.record(); o = ow(features, filters, gout, pairs, num); b.record()
    rows.append(["wgrad", features.shape[0], gout.shape[0], int(num.sum()), filters.shape[-2], filters.shape[-1], a, b]); return o
for _ in range(2):
    model(pts, [feats], metas).square().mean().backward()
ops.sparse_conv_forward, ops.sparse_conv_dgrad, ops.sparse_conv_wgrad = fwd, dgrad, wgrad
model(pts, [feats], metas).square().mean().backward()
torch.cuda.synchronize()
tot = {}
for kind, n_src, n_dst, pairs, cin, cout, a, b in rows:
    ms = a.elapsed_time(b); fl = 2.0 * pairs * cin * cout
    tot[kind] = tot.get(kind, 0) + ms
    print("%-5s src %7d dst %7d pairs %9d  %3d->%3d  %7.3f ms  %7.1f TFLOP/s  %6.1f GB/s(alg)" % (
        kind, n_src, n_dst, pairs, cin, cout, ms, fl / ms / 1e9, (4.0 * (n_src * cin + n_dst * cout) + 8 * pairs) / ms / 1e6))
print(tot)
