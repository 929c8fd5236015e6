#!/usr/bin/env python
"""Time single phases of the step (plan / stage / fwd / bwd) the way bench.py does -- one CUDA graph of the phase's
launches over 4 rotating batch sets, replayed between CUDA events -- without running whole steps, so that debug
builds whose later phases are meaningless can still be timed.  LSS_B200_LIB selects the library.

    python tools/time_phase.py [--config config2] plan stage fwd bwd
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from lss2_multimodal_nu_b200 import functional as F, synthetic as S  # noqa: E402
from lss2_multimodal_nu_b200.pipeline import LiftSplatStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config2")
    ap.add_argument("--B", type=int, default=0)
    ap.add_argument("phases", nargs="*", default=["plan", "stage", "fwd", "bwd"])
    a = ap.parse_args()
    cfg = S.config(a.config, **({"B": a.B} if a.B else {}))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    grid = F.GridSpec.from_bounds(cfg.xbound, cfg.ybound, cfg.zbound)
    us, vs, ds = F.frustum_axes(F.make_frustum(cfg.final_dim, cfg.downsample, cfg.dbound).to(dev))
    stream = torch.cuda.Stream(dev)
    steps = []
    for s in range(4):
        st = LiftSplatStep(cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, cfg.C, grid, us, vs, ds, device=dev, capture=False,
                           stream=stream)
        cal = S.make_calibration(cfg, 1234 + s); ft = S.make_features(cfg, 1234 + s)
        st.load({k: torch.from_numpy(v) for k, v in {**cal, **ft}.items()})
        st._dbev.normal_()
        st.run()                       # every phase once, in order: later phases need the earlier ones' outputs
        steps.append(st)
    torch.cuda.synchronize()
    out = []
    for name in a.phases:
        with torch.cuda.stream(stream):
            for d in steps:
                getattr(d, "enqueue_" + name)(stream.cuda_stream)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            cur = torch.cuda.current_stream(dev).cuda_stream
            for _ in range(2):
                for d in steps:
                    getattr(d, "enqueue_" + name)(cur)
        ts = []
        with torch.cuda.stream(stream):
            g.replay()
            for _ in range(20):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); g.replay(); e1.record(stream)
                ts.append((e0, e1))
        stream.synchronize()
        per = sorted(x.elapsed_time(y) * 1e3 / 8 for x, y in ts)
        out.append("%s %.2f" % (name, per[len(per) // 2]))
    print(os.environ.get("LSS_B200_LIB", "default"), " ".join(out), "us")


if __name__ == "__main__":
    main()
