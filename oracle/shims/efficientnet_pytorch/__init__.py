"""Stand-in for the `efficientnet_pytorch` package (not installed; its pretrained weights need network).

TEST / BENCH SCAFFOLDING ONLY.  It exposes exactly the attribute surface the reference's
unmodified `Encoder.get_eff_depth` walks (reference src/modules.py:33,44-60): `_conv_stem`, `_bn0`, `_swish`,
`_blocks` (each called as block(x, drop_connect_rate=...)), `_global_params.drop_connect_rate`, and produces the
EfficientNet-B4 endpoint shapes the reference's `Up(448+160, 512)` expects: reduction_4 = 160 channels at 1/16,
reduction_5 = 448 channels at 1/32.  The trunk is a handful of strided convolutions -- the backbone is out of
scope for this build; only its output contract matters to the lift-splat stage.
"""
import types

import torch
from torch import nn


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x, drop_connect_rate=None):
        return torch.relu(self.bn(self.conv(x)))


class EfficientNet(nn.Module):
    def __init__(self):
        super().__init__()
        self._conv_stem = nn.Conv2d(3, 16, 3, stride=2, padding=1, bias=False)   # 1/2
        self._bn0 = nn.BatchNorm2d(16)
        self._swish = nn.SiLU()
        self._blocks = nn.ModuleList([
            _Block(16, 24, 2),      # 1/4
            _Block(24, 56, 2),      # 1/8
            _Block(56, 160, 2),     # 1/16  -> reduction_4
            _Block(160, 448, 2),    # 1/32  -> reduction_5
        ])
        self._global_params = types.SimpleNamespace(drop_connect_rate=0.2)

    @classmethod
    def from_pretrained(cls, name, **kwargs):
        return cls()

    @classmethod
    def from_name(cls, name, **kwargs):
        return cls()
