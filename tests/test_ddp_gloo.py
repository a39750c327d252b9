"""N>1 host-side logic on CPU: two gloo ranks run the data-parallel harness of bench.py (DDP over the
hot-path module, structurally unused parameters frozen, per-rank synthetic samples) with the ops
driven by the oracle CPU path, and must end the step with identical, all-reduced gradients."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SMALL_RANGE = [-12.0, -12.0, -5.0, 12.0, 12.0, 3.0]   # 320 x 320 x 41 grid at the nuScenes voxel size


def small_model():
    import configs
    import ddf_b200.fusion.point_fusion  # noqa: F401
    import ddf_b200.fusion.sparse_encoder  # noqa: F401
    import ddf_b200.fusion.voxel_encoder  # noqa: F401
    from ddf_b200.fusion import structurally_unused_parameters
    from ddf_b200.fusion.detector import TransFusionPtsBranch
    cfg = configs.transfusion_f()
    cfg["pts_voxel_layer"]["point_cloud_range"] = SMALL_RANGE
    cfg["pts_middle_encoder"]["point_cloud_range"] = SMALL_RANGE
    cfg["pts_middle_encoder"]["sparse_shape"] = [41, 320, 320]
    torch.manual_seed(0)
    m = TransFusionPtsBranch(**cfg).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    frozen = set(structurally_unused_parameters(m))
    for n, p in m.named_parameters():
        if n in frozen:
            p.requires_grad_(False)
    return m, frozen


def sample(rank):
    pts = synth.lidar_points(6000, seed=100 + rank)
    pts = pts[(np.abs(pts[:, 0]) < 12) & (np.abs(pts[:, 1]) < 12)]
    feats = torch.from_numpy(synth.camera_features(1, 6, (28, 50), seed=rank))
    meta = synth.nusc_img_meta(6, input_hw=(112, 200))
    return [torch.from_numpy(pts)], feats, [meta]


def worker(rank, world, port, out_dir, flat=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_path
    torch.set_num_threads(2)
    model, frozen = small_model()
    exchange = None
    if flat:
        # bench.py's default N > 1 path: plain module + one flat all-reduce after backward
        from ddf_b200.data_parallel import GradientExchange
        if rank != 0:
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)                     # the broadcast must undo this
        GradientExchange.broadcast_initial_state(model)
        exchange = GradientExchange(model.parameters())
        net = model
    else:
        net = torch.nn.parallel.DistributedDataParallel(model, broadcast_buffers=False)
    pts, feats, metas = sample(rank)
    with cpu_path.reference_cpu_ops():
        out = net(pts, [feats], metas)
        loss = out.square().mean()
        loss.backward()
    if exchange is not None:
        exchange.exchange()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(exchange.params, exchange.views))
        exchange.exchange()                         # gradients already in the flat buffer: no copy, averages again (no-op on equal values)
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert not (set(grads) & frozen)
    torch.save(dict(grads=grads, loss=float(loss.detach())), os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("flat", [False, True], ids=["ddp_wrapper", "gradient_exchange"])
def test_two_rank_data_parallel_step(tmp_path, flat):
    port = 29500 + (os.getpid() + 7 * flat) % 2000
    mp.spawn(worker, args=(2, port, str(tmp_path), flat), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["loss"] != r1["loss"]                      # different samples per rank (weak scaling)
    assert set(r0["grads"]) == set(r1["grads"]) and len(r0["grads"]) > 100
    for k in r0["grads"]:
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k   # all-reduced (averaged) gradients
    # and they are the mean of the single-rank gradients
    from oracle import cpu_path
    model, _ = small_model()
    acc = None
    for rank in range(2):
        model.zero_grad()
        pts, feats, metas = sample(rank)
        with cpu_path.reference_cpu_ops():
            model(pts, [feats], metas).square().mean().backward()
        g = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    for k in ("pts_middle_encoder.conv_input.0.weight", "pts_middle_encoder.conv_out.0.weight",
              "pts_middle_encoder.fusion_layer.actr.transformer.encoder.layers.0.linear1.weight"):
        want = acc[k] / 2   # thread count differs between the runs: fp32 reassociation noise only
        assert float((r0["grads"][k] - want).abs().max()) < 2e-3 * float(want.abs().max()), k
