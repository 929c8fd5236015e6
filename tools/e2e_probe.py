"""Probe: host pipeline variants (slots in flight, copies inside/outside the slot graph)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss2_multimodal_nu_b200 import functional as F, synthetic as S
from lss2_multimodal_nu_b200.pipeline import HostPipeline, LiftSplatStep
cfg = S.config("config2"); dev = torch.device("cuda:0")
grid = F.GridSpec.from_bounds(cfg.xbound, cfg.ybound, cfg.zbound)
us, vs, ds = F.frustum_axes(F.make_frustum(cfg.final_dim, cfg.downsample, cfg.dbound).to(dev))
cal = S.make_calibration(cfg); ft = S.make_features(cfg)
host = {k: torch.from_numpy(v).pin_memory() for k, v in {**cal, **ft}.items()}
mk = lambda: LiftSplatStep(cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, cfg.C, grid, us, vs, ds, device=dev)
for depth in (2, 3, 4, 6):
    for gio in (False, True):
        pipe = HostPipeline(mk, depth=depth, graph_io=gio)
        for k in range(depth):
            pipe.pack(host, pipe.input_block(k))
        def run(n):
            t_sub = t_col = 0.0
            for i in range(n):
                if pipe.in_flight() == depth:
                    t = time.perf_counter(); pipe.collect(); t_col += time.perf_counter() - t
                t = time.perf_counter(); pipe.submit(); t_sub += time.perf_counter() - t
            while pipe.in_flight():
                pipe.collect()
            return t_sub, t_col
        run(10); torch.cuda.synchronize()
        t0 = time.perf_counter(); ts, tc = run(300); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("depth %d graph_io %-5s: %.1f us/step (%.0f samples/s); host submit %.1f us, collect wait %.1f us per step" % (
            depth, gio, dt / 300 * 1e6, cfg.B * 300 / dt, ts / 300 * 1e6, tc / 300 * 1e6))
        del pipe
