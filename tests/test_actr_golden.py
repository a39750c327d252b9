"""Module-graph parity of the 3D-DF fusion encoder against the REFERENCE's own classes.

tests/golden/actr_golden.{npz,json} hold the outputs (and the state-dict signature) of the reference's
``build(...)`` -> ``ACTR`` for the three forks' live configurations and for every query-mixing /
pos-encoding mode (generator: tests/golden/make_actr_golden.py, which imports the reference's files).
Weights and inputs are regenerated on both sides by tests/golden/detfill.py from key names, so

* the state-dict KEYS and SHAPES of ``ddf_b200.fusion.actr.build`` must equal the reference's, and
* the outputs must agree: 1e-5 on the CPU (our module graph, native ops swapped for the oracle), 1e-3 rel on
  CUDA through the real kernels (north_star tolerance).
"""
import json
import os
import sys

import numpy as np
import pytest
import torch
from torch import nn

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import detfill  # noqa: E402
import recipes  # noqa: E402

with open(os.path.join(GOLDEN, "actr_golden.json")) as f:
    META = json.load(f)
GOLD = np.load(os.path.join(GOLDEN, "actr_golden.npz"))
CASES = sorted(META)


def make_inputs(name, cfg_or_case, dims=None, valid=None):
    if dims is None:
        cfg_or_case, dims, valid = cfg_or_case["cfg"], cfg_or_case["dims"], cfg_or_case["valid"]
    return recipes.actr_inputs(name, cfg_or_case, dims, valid)


def build_ours(case):
    from ddf_b200.fusion import actr
    net = actr.build(case["cfg"], model_name=case["model_name"], lt_cfg=case.get("lt"),
                     hybrid_cfg=case.get("hybrid"), gate_first=case["flavour"] == "VR")
    sig = [[k, list(s)] for k, s in detfill.state_dict_signature(net)]
    assert sig == case["signature"], "state-dict keys / shapes differ from the reference class"
    detfill.fill_state_dict(net)
    return net


def run(net, inputs, modal, device):
    v_feat, grid, i_feat, v_i_feat, lidar = [t.clone().to(device) for t in inputs]
    with torch.no_grad():
        out = net(v_feat, grid, [i_feat], v_i_feat if modal in ("image", "hybrid") else None, lidar)
    return out.float().cpu().numpy()


def zero_dropout(net):
    for m in net.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0


def check(name, device, tol):
    case = META[name]
    net = build_ours(case).to(device)
    inputs = make_inputs(name, case)
    modal = case["cfg"].get("feature_modal", "lidar")
    net.eval()
    errs = {}
    got = run(net, inputs, modal, device)
    ref = GOLD[name + "/eval"]
    errs["eval"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    if case.get("train"):
        zero_dropout(net)
        net.train()
        got = run(net, inputs, modal, device)
        ref = GOLD[name + "/train"]
        errs["train"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert all(e <= tol for e in errs.values()), (name, errs)
    return errs


@pytest.mark.parametrize("name", CASES)
def test_module_graph_matches_reference_classes_cpu(name):
    from oracle import cpu_path
    with cpu_path.reference_cpu_ops():
        check(name, "cpu", 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_module_graph_matches_reference_classes_cuda(name):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    check(name, "cuda", 1e-3)
