#!/usr/bin/env python
"""Informational baseline: the reference's lift + voxel_pooling OP SEQUENCE in stock PyTorch on the GPU.

The north star quotes the speed-up against "the reference PyTorch-on-GPU voxel_pooling".  The reference
itself cannot travel to the GPU box (it imports packages this image does not have and lives outside the
repo), so this script restates its sequence of ATen calls with plain torch ops on the same synthetic
inputs -- the same sequence oracle/lss_oracle.py restates in numpy -- and times forward + backward with
CUDA events:

    geometry (sub, inverse, matmul, mul, cat, matmul, add)          src/model_baseline.py:50-70
    lift outer product depth.unsqueeze(1) * feat.unsqueeze(2)       src/modules.py:84
    view + permute + reshape (copy)                                 src/model_baseline.py:79-80,89
    quantise, batch index, bounds mask, boolean-mask gathers        src/model_baseline.py:92-103
    rank, argsort, gathers                                          src/model_baseline.py:106-111
    cumsum trick with a custom backward (gather)                    src/tools.py:192-218
    zeros, index_put, cat(unbind)                                   src/model_baseline.py:120-124

It is NOT the product path and nothing in the package, the tests or bench.py imports it; it also checks
its own output against the CUDA path so the two numbers describe the same computation.

    python tools/torch_baseline_gpu.py [config2] [--steps 30]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lss_oracle as O  # noqa: E402
from lss2_multimodal_nu_b200 import functional as F, synthetic as S  # noqa: E402


class CumsumTrick(torch.autograd.Function):
    """Segmented sum through a global prefix sum; backward = gather of the voxel gradient."""

    @staticmethod
    def forward(ctx, x, coords, ranks):
        x = x.cumsum(0)
        last = torch.ones(x.shape[0], device=x.device, dtype=torch.bool)
        last[:-1] = ranks[1:] != ranks[:-1]
        x, coords = x[last], coords[last]
        x = torch.cat((x[:1], x[1:] - x[:-1]))
        ctx.save_for_backward(last)
        ctx.mark_non_differentiable(coords)
        return x, coords

    @staticmethod
    def backward(ctx, gx, gcoords):
        (last,) = ctx.saved_tensors
        run = torch.cumsum(last, 0)
        run[last] -= 1
        return gx[run], None, None


def geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    B, N, _ = trans.shape
    pts = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
    pts = torch.cat((pts[:, :, :, :, :, :2] * pts[:, :, :, :, :, 2:3], pts[:, :, :, :, :, 2:3]), 5)
    comb = rots.matmul(torch.inverse(intrins))
    pts = comb.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
    return pts + trans.view(B, N, 1, 1, 1, 3)


def lift_and_pool(depth, feat, geom, dx, bx, nx, B, N):
    BN, D, fH, fW = depth.shape
    C = feat.shape[1]
    x = depth.unsqueeze(1) * feat.unsqueeze(2)                       # (BN, C, D, fH, fW)
    x = x.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    P = B * N * D * fH * fW
    x = x.reshape(P, C)
    g = ((geom - (bx - dx / 2.)) / dx).long().view(P, 3)
    bix = torch.cat([torch.full([P // B, 1], i, device=x.device, dtype=torch.long) for i in range(B)])
    g = torch.cat((g, bix), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x, g = x[kept], g[kept]
    ranks = g[:, 0] * (nx[1] * nx[2] * B) + g[:, 1] * (nx[2] * B) + g[:, 2] * B + g[:, 3]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x, g = CumsumTrick.apply(x, g, ranks)
    final = torch.zeros((B, C, int(nx[2]), int(nx[0]), int(nx[1])), device=x.device)
    final[g[:, 3], :, g[:, 2], g[:, 0], g[:, 1]] = x
    return torch.cat(final.unbind(dim=2), 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="config2")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    cfg = S.config(a.config)
    dev = torch.device("cuda:0")
    cal = {k: torch.from_numpy(v).to(dev) for k, v in S.make_calibration(cfg).items()}
    ft = {k: torch.from_numpy(v).to(dev) for k, v in S.make_features(cfg).items()}
    dbev = torch.from_numpy(S.make_dbev(cfg)).to(dev)
    frustum = torch.from_numpy(O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound)).to(dev)
    dxn, bxn, nxn = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    dx, bx, nx = torch.from_numpy(dxn).to(dev), torch.from_numpy(bxn).to(dev), torch.from_numpy(nxn).to(dev)

    def step(geom=None):
        depth = ft["depth"].clone().requires_grad_(True)
        feat = ft["feat"].clone().requires_grad_(True)
        if geom is None:
            geom = geometry(frustum, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
        bev = lift_and_pool(depth, feat, geom, dx, bx, nx, cfg.B, cfg.N)
        bev.backward(dbev)
        return bev, depth.grad, feat.grad

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # the CUDA path on the same inputs: same computation?
    grid = F.GridSpec(tuple(map(float, dxn)), tuple(map(float, bxn)), tuple(map(int, nxn)))
    us, vs, ds = F.frustum_axes(frustum)
    plan = F.build_plan(us, vs, ds, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"], grid)
    d = ft["depth"].clone().requires_grad_(True); f = ft["feat"].clone().requires_grad_(True)
    ours = F.lift_splat(d, f, plan); ours.backward(dbev)
    # (cuBLAS evaluates the batched 3x3 products with FMAs, so a few points near voxel borders land in
    #  other voxels than on the CPU: SURVEY.md 7.3-1; the check therefore feeds the torch pooling the
    #  bit-exact geometry of the CUDA path)
    g_exact = F.geometry(us, vs, ds, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"],
                         grid, want_geom=True)["geom"]
    out = step(g_exact)
    gt = geometry(frustum, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
    moved = int((((gt - (bx - dx / 2.)) / dx).long() != ((g_exact - (bx - dx / 2.)) / dx).long()).any(-1).sum())
    err = (ours - out[0]).abs().max().item()
    gerr = max((d.grad - out[1]).abs().max().item(), (f.grad - out[2]).abs().max().item())
    print("stock PyTorch on the GPU, %s: %.3f ms per fwd+bwd step = %.0f samples/s  (max |diff| to the CUDA path: "
          "bev %.2e, grads %.2e on identical geometry; torch's own GPU geometry moves %d of %d points to "
          "another voxel; peak memory %.2f GB)" % (cfg.name, ms, cfg.B / ms * 1e3, err, gerr, moved, cfg.P,
                                                  torch.cuda.max_memory_allocated() / 1e9))
    # where the time goes
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))


if __name__ == "__main__":
    main()
