#!/usr/bin/env python
"""Per-phase timeline of the plan kernels (debug build with -DLSS_PHASE_TIMING).

    python tools/phase_timing.py        # builds /tmp/liblss_timing.so, runs config2, prints phase times
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = os.path.join(ROOT, "lss2_multimodal_nu_b200", "csrc")
so = os.path.join(ROOT, "gpurun_out", "liblss_timing.so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                "-shared", "-DLSS_PHASE_TIMING", "-I", os.path.join(ROOT, "include"), "-I", src, "-o", so,
                os.path.join(src, "lss_abi.cu")], check=True)
from lss2_multimodal_nu_b200 import _abi
_abi.LIB_PATH = so
from lss2_multimodal_nu_b200 import functional as F, synthetic as S
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lss_oracle as O

cfg = S.config(sys.argv[1] if len(sys.argv) > 1 else "config2")
dev = "cuda:0"
cal = {k: torch.from_numpy(v).to(dev) for k, v in S.make_calibration(cfg).items()}
us, vs, ds = (torch.from_numpy(a).to(dev) for a in O.frustum_axes(cfg.final_dim, cfg.downsample, cfg.dbound))
grid = F.GridSpec.from_bounds(cfg.xbound, cfg.ybound, cfg.zbound)
ft = {k: torch.from_numpy(v).to(dev) for k, v in S.make_features(cfg).items()}
for _ in range(5):
    plan = F.build_plan(us, vs, ds, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"], grid)
    bev = F.lift_splat(ft["depth"], ft["feat"], plan)
torch.cuda.synchronize()
lib = _abi.load()
lib.lss_debug_phase_ts.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
for kernel, name, nblk in ((0, "partition_coop", (cfg.P + 1023) // 1024), (1, "local_sort", 1024),
                           (2, "pool_fwd (0-1 FILL warp, 2-6 first REDUCE warp)", 148 * 8)):
    buf = np.zeros(4096 * 8, np.uint64)
    lib.lss_debug_phase_ts(kernel, buf.ctypes.data, buf.size)
    ts = buf.reshape(4096, 8)[:nblk].astype(np.int64)
    t0 = ts[:, 0].min()
    rel = (ts - t0) / 1e3
    print(name, "blocks", nblk)
    for s in range(8):
        col = rel[:, s][ts[:, s] > 0]
        if len(col):
            print("  stamp %d: min %.2f  median %.2f  max %.2f us  (n=%d)" % (s, col.min(), np.median(col), col.max(), len(col)))
    if kernel == 0:
        d = (ts[:, 4] - ts[:, 3]) / 1e3
        print("  phase B duration by tile: ", " ".join("%.1f" % v for v in d[::8]))
        d = (ts[:, 7] - ts[:, 6]) / 1e3
        print("  rank+scatter duration:    ", " ".join("%.1f" % v for v in d[::8]))
        d = (ts[:, 2] - ts[:, 1]) / 1e3
        print("  geometry duration:        ", " ".join("%.1f" % v for v in d[::8]))
