"""N identical bench steps and nothing else (the target of the ncu launch-list pass: the last 1/N of the captured
launches is one warm step).  python tools/step_only.py --config tf [N=4 via DDF_STEPS]"""
import os, sys
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
for _ in range(int(os.environ.get("DDF_STEPS", "4"))):
    loss = wl.forward(model, t, static).square().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
