"""Bi-directional gated fusion blocks between the LiDAR and image query streams
(<proj>/models/model_utils/attentions.py:48-117; ``attn_dict`` :144-149). Parameter names
``b_conv1d`` / ``a_conv1d`` (Conv1d C->1, kernel 1) as in the reference."""
import torch
from torch import nn

from ..ops import fused as _fused


class _BiGateBase(nn.Module):
    def __init__(self, g_channel, g_channel_):
        super().__init__()
        self.g_channel = g_channel
        self.g_channel_ = g_channel_
        self.b_conv1d = nn.Conv1d(g_channel, 1, kernel_size=1, stride=1, padding=0)
        self.a_conv1d = nn.Conv1d(g_channel_, 1, kernel_size=1, stride=1, padding=0)

    @staticmethod
    def _gate(conv, x):
        # Conv1d(C -> 1, k=1) over (B, L, C) laid out channel-last is a per-row dot product: run it as
        # a matmul on the rows (same parameters, weight [1, C, 1]) instead of a cuDNN grouped conv
        return torch.sigmoid(torch.nn.functional.linear(x, conv.weight.squeeze(-1), conv.bias))


class BiGate1D(_BiGateBase):
    """s1 = sigmoid(b(f1)), s2 = sigmoid(a(f2)); each stream is scaled by the OTHER stream's gate
    (attentions.py:44-49)."""

    def forward(self, feat1, feat2):
        s1 = self._gate(self.b_conv1d, feat1)
        s2 = self._gate(self.a_conv1d, feat2)
        return feat1 * s2, feat2 * s1


class BiGate1D_2(_BiGateBase):
    """Gates from the summed streams, each stream scaled by its own gate (attentions.py:66-72)."""

    def forward(self, feat1, feat2):
        fuse = feat1 + feat2
        s1 = self._gate(self.b_conv1d, fuse)
        s2 = self._gate(self.a_conv1d, fuse)
        return feat1 * s1, feat2 * s2


class BiGateSum1D(_BiGateBase):
    """Per-stream gates, residual mix: out1 = f1 + f2 * sigmoid(b(f1)); out2 = f2 + f1 * sigmoid(a(f2))
    (attentions.py:89-94)."""

    def forward(self, feat1, feat2):
        out = _fused.bigate_sum(self.b_conv1d, self.a_conv1d, feat1, feat2, False)   # one kernel each way on CUDA
        if out is not None:
            return out
        s1 = self._gate(self.b_conv1d, feat1)
        s2 = self._gate(self.a_conv1d, feat2)
        return feat1 + feat2 * s1, feat2 + feat1 * s2


class BiGateSum1D_2(_BiGateBase):
    """out1 = f1 + f2 * sigmoid(b(f1+f2)); out2 = f2 + f1 * sigmoid(a(f1+f2)) (attentions.py:111-117)."""

    def forward(self, feat1, feat2):
        out = _fused.bigate_sum(self.b_conv1d, self.a_conv1d, feat1, feat2, True)    # one kernel each way on CUDA
        if out is not None:
            return out
        fuse = feat1 + feat2
        s1 = self._gate(self.b_conv1d, fuse)
        s2 = self._gate(self.a_conv1d, fuse)
        return feat1 + feat2 * s1, feat2 + feat1 * s2


attn_dict = {
    "BiGate1D": BiGate1D,
    "BiGate1D_2": BiGate1D_2,
    "BiGateSum1D": BiGateSum1D,
    "BiGateSum1D_2": BiGateSum1D_2,
}
