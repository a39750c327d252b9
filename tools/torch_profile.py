"""torch.profiler view of one bench step (operator names + input shapes by device time): which ATen ops the non-library
share of the step is made of."""
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
def step():
    loss = wl.forward(model, t, static).square().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=130, max_name_column_width=45,
                                                          max_shapes_column_width=70))
if os.environ.get("DDF_PROFILE_FLAT"):
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=110, max_name_column_width=90))
if os.environ.get("DDF_PROFILE_STACKS"):
    # second pass: attribute the ATen kernels to source lines of this package
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof2:
        step()
        torch.cuda.synchronize()
    print(prof2.key_averages(group_by_stack_n=6).table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=40,
                                                       max_src_column_width=110))
