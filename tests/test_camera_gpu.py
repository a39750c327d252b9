"""Camera branch (f-4): the CUDA-graphed bf16 channels-last branch returns what the same network returns eagerly in
fp32 (bf16 tolerance), keeps its static output across replays, and stays frozen."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_camera_branch_graph_matches_fp32_eager():
    from ddf_b200.fusion.camera import CameraBranch, ResNet50FPN0
    torch.manual_seed(0)
    ref = ResNet50FPN0().cuda().eval()
    cam = CameraBranch().cuda()
    cam.net.load_state_dict(ref.state_dict())
    cam.train()                                   # must stay frozen / eval
    assert not cam.training and all(not p.requires_grad for p in cam.parameters())
    imgs = torch.randint(0, 256, (2, 3, 128, 160), dtype=torch.uint8, device="cuda")
    out = cam(imgs).float().clone()
    assert out.shape == (2, 256, 32, 40)
    with torch.no_grad():
        expect = ref(imgs.float() - cam.mean)
    err = float((out - expect).abs().max() / expect.abs().max())
    assert err < 5e-2, err                        # bf16 activations through 50 layers
    imgs2 = torch.randint(0, 256, (2, 3, 128, 160), dtype=torch.uint8, device="cuda")
    out2 = cam(imgs2).float().clone()
    assert float((out2 - out).abs().max()) > 0    # the replay really consumed the new images
    assert torch.equal(cam(imgs).float(), out)    # and is deterministic
