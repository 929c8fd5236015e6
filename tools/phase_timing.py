#!/usr/bin/env python
"""Per-warp timeline of the fused forward kernel (debug build with -DLSS_PHASE_TIMING).

    python tools/phase_timing.py [config2]   # builds gpurun_out/liblss_timing.so, runs the config, prints the timeline
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
os.makedirs(os.path.dirname(so), exist_ok=True)
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    plan = F.build_plan(us, vs, ds, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"], grid)
    flush.zero_()
    bev = F.lift_splat(ft["depth"], ft["feat"], plan)
torch.cuda.synchronize()
lib = _abi.load()
lib.lss_debug_phase_ts.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
n_keys = F.n_keys(grid, cfg.B)
fill = min(((n_keys + 31) // 32 + 7) // 8, 148)
fill = int(os.environ.get("LSS_FILL_CTAS", fill))
chunk = 32; nblk = min(4096, fill + (cfg.P + chunk * 8 - 1) // (chunk * 8))
buf = np.zeros(4096 * 16, np.uint64)
lib.lss_debug_phase_ts(2, buf.ctypes.data, buf.size)
ts = buf.reshape(4096, 8, 2)[:nblk].astype(np.int64)
ok = (ts[:, :, 0] > 0) & (ts[:, :, 1] > 0)
t0 = ts[:, :, 0][ok].min()
start = (ts[:, :, 0] - t0) / 1e3
end = (ts[:, :, 1] - t0) / 1e3
dur = end - start
print("pool_fwd: %d CTAs (%d fill); kernel span %.1f us" % (nblk, fill, end[ok].max()))
for name, sl in (("fill", slice(0, fill)), ("reduce", slice(fill, nblk))):
    o = ok[sl]
    if not o.any():
        continue
    print("%-6s warps %5d: start p50 %.1f p90 %.1f max %.1f | duration p50 %.2f p90 %.2f p99 %.2f max %.2f | end p50 %.1f p99 %.1f max %.1f us" % (
        name, o.sum(), *np.percentile(start[sl][o], [50, 90, 100]), *np.percentile(dur[sl][o], [50, 90, 99, 100]),
        *np.percentile(end[sl][o], [50, 99, 100])))
hist, edges = np.histogram(start[ok], bins=12)
print("  warp starts per %.1f us bin:" % (edges[1] - edges[0]), hist.tolist())
hist, edges = np.histogram(end[ok], bins=12)
print("  warp ends   per %.1f us bin:" % (edges[1] - edges[0]), hist.tolist())
