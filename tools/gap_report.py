"""Where the GPU idles inside a bench step: torch.profiler trace of one warm step, kernels sorted by start time, every
gap longer than 15 us between consecutive kernels (any stream) listed with the kernels around it."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
sys.argv = ["bench.py"] + sys.argv[1:]
import bench
args = bench.parse()
wl = bench.WORKLOADS[args.config]()
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
model = bench.build_model(wl, dev)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01, fused=True)
t, static = wl.host_batch(0)
t = bench.map_tensors(t, lambda x: x.to(dev))
def step():
    loss = wl.forward(model, t, static).square().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
    opt.step()
for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_step.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e], key=lambda e: e["ts"])
busy_end = ks[0]["ts"]
gaps = []
total_gap = 0.0
for a in ks:
    if a["ts"] > busy_end:
        g = a["ts"] - busy_end
        total_gap += g
        if g > 15:
            gaps.append((g, prev["name"][:60], a["name"][:60], a["ts"] - ks[0]["ts"]))
    if a["ts"] + a["dur"] > busy_end:
        busy_end = a["ts"] + a["dur"]
        prev = a
span = busy_end - ks[0]["ts"]
print("span %.2f ms, idle %.2f ms in %d gaps (> 15 us: %d gaps, %.2f ms)" % (span / 1e3, total_gap / 1e3, len(ks), len(gaps), sum(g[0] for g in gaps) / 1e3))
for g in sorted(gaps, reverse=True)[:40]:
    print("%7.0f us at %7.2f ms  after %-60s before %s" % (g[0], g[3] / 1e3, g[1], g[2]))
