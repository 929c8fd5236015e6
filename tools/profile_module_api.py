#!/usr/bin/env python
"""Where the host time of the eager drop-in path goes: model.get_voxels + backward on the reference's LSS class
(oracle/_ref) with patch.install, under cProfile; prints device time per step and the top host functions."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
from lss2_multimodal_nu_b200 import patch, synthetic as S  # noqa: E402

cfg = S.config("config2")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).to(dev).train()
patch.install(m)
cal = [torch.from_numpy(v).to(dev) for v in S.make_calibration(cfg, 1234).values()]
x = torch.randn(cfg.B * cfg.N, 512, cfg.fH, cfg.fW, device=dev, requires_grad=True)
dbev = torch.randn(cfg.B, 200, 200, cfg.C, device=dev).permute(0, 3, 1, 2)


def one():
    x.grad = None
    m.get_voxels(x, *cal).backward(dbev)


for _ in range(20):
    one()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
pr = cProfile.Profile()
e0.record()
pr.enable()
for _ in range(200):
    one()
pr.disable()
e1.record()
torch.cuda.synchronize()
print("%.1f us per step (device clock, host-bound if >> 70)" % (e0.elapsed_time(e1) * 1e3 / 200))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
