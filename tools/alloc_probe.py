import os, sys, time
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
s0 = torch.cuda.memory_stats()
# time every torch.empty
orig = torch.empty
slow = []
def timed_empty(*a, **k):
    t0 = time.perf_counter()
    r = orig(*a, **k)
    dt = time.perf_counter() - t0
    if dt > 20e-6:
        slow.append((dt * 1e6, tuple(r.shape), str(r.dtype)))
    return r
torch.empty = timed_empty
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("ms/step wall", (time.perf_counter() - t0) / 5 * 1e3)
torch.empty = orig
s1 = torch.cuda.memory_stats()
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"):
    print(k, s0.get(k), "->", s1.get(k))
print("reserved GB", torch.cuda.memory_reserved() / 1e9, "max allocated GB", torch.cuda.max_memory_allocated() / 1e9)
print("slow torch.empty calls:", len(slow), "total us", sum(s[0] for s in slow))
slow.sort(reverse=True)
for s in slow[:25]:
    print(s)
