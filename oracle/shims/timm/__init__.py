"""Stand-in for `timm` (not installed).  TEST / BENCH SCAFFOLDING ONLY.

The reference's VoVNetV2 wrapper (reference src/vovnet_timm.py:27,48-58) calls
`timm.create_model('ese_vovnet39b' | 'ese_vovnet57b', pretrained=..., features_only=True, out_indices=(3, 4))`
and expects a module returning `[c3 (768 channels, 1/16), c4 (1024 channels, 1/32)]`.  This returns a few strided
convolutions with that output contract; the backbone itself is out of scope for this build.
"""
import torch
from torch import nn


class _Features(nn.Module):
    def __init__(self):
        super().__init__()

        def block(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(cout),
                                 nn.ReLU(inplace=True))
        self.s1 = block(3, 32)       # 1/2
        self.s2 = block(32, 64)      # 1/4
        self.s3 = block(64, 128)     # 1/8
        self.c3 = block(128, 768)    # 1/16
        self.c4 = block(768, 1024)   # 1/32

    def forward(self, x):
        c3 = self.c3(self.s3(self.s2(self.s1(x))))
        return [c3, self.c4(c3)]


def create_model(name, pretrained=False, features_only=False, out_indices=None, **kwargs):
    if not features_only:
        raise NotImplementedError("timm shim: only features_only=True is provided")
    return _Features()
