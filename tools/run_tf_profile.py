import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools'), os.path.join(ROOT, 'tests')]
import torch, numpy as np
import configs, synth
from ddf_b200.fusion.detector import TransFusionPtsBranch
import ddf_b200.fusion.point_fusion, ddf_b200.fusion.sparse_encoder, ddf_b200.fusion.voxel_encoder
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
m = TransFusionPtsBranch(**configs.transfusion_f()).cuda().train()
B = 2
pts = [torch.from_numpy(synth.lidar_points(260000, seed=b)).cuda() for b in range(B)]
img_feats = [torch.from_numpy(synth.camera_features(B, 6, (112, 200))).cuda()]
metas = [synth.nusc_img_meta() for _ in range(B)]
opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.time()
    out = m(pts, img_feats, metas)
    loss = out.square().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    print(it, out.shape, float(loss), 'ms', (time.time() - t0) * 1e3)
print('max mem GB', torch.cuda.max_memory_allocated() / 1e9)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    out = m(pts, img_feats, metas); loss = out.square().mean(); opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=60))
