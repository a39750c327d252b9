"""Per-CTA phase clocks of the multi-tile conv kernel (instrumented library, tools/build_trace.py -DDDF_PHASES):
runs one layer's forward launch and summarises, per SM, how the CTAs' lifetimes split into alloc / table prologue /
main loop / accumulator wait / epilogue / teardown, and the gaps between successive CTAs of an SM."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench_ops
from ddf_b200 import lib as _lib
from ddf_b200.ops.spconv import ops

class A: pass
args = A(); args.batch = 2; args.points = 260000
idx, _ = bench_ops._scene(args)
shape = [41, 1440, 1440]
want = sys.argv[1] if len(sys.argv) > 1 else "32->32"
L = _lib.get_lib()
L.ddf_phase_dump.argtypes = [ctypes.c_char_p]
for name, subm, ks, st, pad, cin, cout in bench_ops.STAGES:
    n = idx.shape[0]
    rb = ops.build_rulebook(idx, 2, shape, ks, st, pad, 1, 0, subm, False)
    n_out = rb.outids.shape[0]
    if want in name:
        feat = torch.randn(n, cin, device="cuda")
        w = torch.randn(*ks, cin, cout, device="cuda") / (cin * 27) ** 0.5
        f_in = ops.split_bf16x3(feat)[0]
        for _ in range(3):
            ops.sparse_conv_forward(f_in, w, rb.gather_table, None, n_out, 1)
        torch.cuda.synchronize()
        L.ddf_phase_dump(b"/dev/null")
        ops.sparse_conv_forward(f_in, w, rb.gather_table, None, n_out, 1)
        torch.cuda.synchronize()
        path = os.path.join(ROOT, "gpurun_out", "phases_%s.csv" % name.split()[-1].replace("->", "_"))
        L.ddf_phase_dump(path.encode())
        d = np.loadtxt(path, delimiter=",", dtype=np.int64).reshape(-1, 10)
        ph = d[:, 2:9]
        dur = np.diff(ph, axis=1)
        names = ["entry->alloc", "table prologue", "main loop (producers)", "accum wait", "epilogue", "teardown"]
        print(name, "CTAs", len(d), "T", d[0, 9], "mean CTA life", (ph[:, 6] - ph[:, 0]).mean())
        for i, nm in enumerate(names):
            print("  %-24s mean %8.0f  p90 %8.0f cycles" % (nm, dur[:, i].mean(), np.percentile(dur[:, i], 90)))
        # per-SM: span, sum of lives, gaps
        spans, lives = [], []
        for sm in np.unique(d[:, 0]):
            e = d[d[:, 0] == sm]
            spans.append(e[:, 8].max() - e[:, 2].min())
            lives.append((e[:, 8] - e[:, 2]).sum())
        print("  per SM: kernel span mean %.0f cycles, sum of CTA lives / span = %.2f (CTA slots busy)" % (np.mean(spans), np.sum(lives) / np.sum(spans)))
    if not subm:
        idx = rb.outids
        shape = rb.out_spatial_shape
