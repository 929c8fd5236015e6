#!/usr/bin/env python
"""The hot path inside a data-parallel training step (BASELINE config 3 in miniature).

The reference's full models need packages and weights this image does not have (efficientnet_pytorch,
timm, nuScenes), so this harness wires the patched Lift-Splat stage between two small stand-ins -- a
strided conv "backbone" in front, a conv "BEV encoder" + head behind -- exactly where
BEV_TXT.forward has it (reference src/model_BEV_TXT.py:279-283), wraps the model in
DistributedDataParallel (one process per GPU, NCCL) and trains on synthetic frames.  The lift-splat stage
has no parameters, so NCCL carries only the surrounding model's gradient all-reduce; the path itself
shards by sample.

    python -m torch.distributed.run --nproc-per-node 2 tools/ddp_train_standin.py [--steps 30] [--bsize 8]

Prints per-step time, aggregate frames/s and checks that (a) every rank holds identical parameters after
training (the all-reduce worked), (b) the loss went down, (c) the state_dict has no extra entries.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as Fnn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lss_oracle as O  # noqa: E402  (frustum / grid constants only)
from lss2_multimodal_nu_b200 import patch, synthetic as S  # noqa: E402


class CamEncode(nn.Module):
    """Attribute surface of the reference's CamEncode (src/modules.py:69-91)."""

    def __init__(self, D, C, cin):
        super().__init__()
        self.D, self.C = D, C
        self.depthnet = nn.Conv2d(cin, D + C, kernel_size=1, padding=0)

    def get_depth_dist(self, x, eps=1e-20):
        return x.softmax(dim=1)


class StandInBevModel(nn.Module):
    """Backbone stand-in -> lift-splat (the reference's method surface, patched) -> BEV head stand-in."""

    def __init__(self, cfg, outC=4, cin=32):
        super().__init__()
        dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        self.dx = nn.Parameter(torch.from_numpy(dx), requires_grad=False)
        self.bx = nn.Parameter(torch.from_numpy(bx), requires_grad=False)
        self.nx = nn.Parameter(torch.from_numpy(nx), requires_grad=False)
        fr = O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound)
        self.frustum = nn.Parameter(torch.from_numpy(fr), requires_grad=False)
        self.D, self.camC, self.bsize, self.downsample = fr.shape[0], cfg.C, cfg.B, cfg.downsample
        self.camencode = CamEncode(self.D, cfg.C, cin)
        self.backbone = nn.Sequential(                     # /16, like the EfficientNet trunk + Up
            nn.Conv2d(3, 16, 3, stride=4, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(16, cin, 3, stride=4, padding=1), nn.ReLU(inplace=True))
        self.bevencode = nn.Sequential(
            nn.Conv2d(cfg.C * int(nx[2]), 32, 7, stride=2, padding=3), nn.ReLU(inplace=True),
            nn.Conv2d(32, outC, 1), nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))

    # the four methods the reference defines; patch.install rebinds them
    def get_geometry(self, *a): raise NotImplementedError
    def get_cam_feats(self, x): raise NotImplementedError
    def voxel_pooling(self, g, x): raise NotImplementedError
    def get_voxels(self, *a): raise NotImplementedError

    def forward(self, imgs, rots, trans, intrins, post_rots, post_trans):
        B, N, C, H, W = imgs.shape
        x = self.backbone(imgs.view(B * N, C, H, W))
        y = self.get_voxels(x, rots, trans, intrins, post_rots, post_trans)
        return self.bevencode(y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--bsize", type=int, default=8)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = S.config("config2", B=a.bsize)
    torch.manual_seed(0)
    model = StandInBevModel(cfg).to(dev)
    keys = list(model.state_dict().keys())
    patch.install(model)
    assert list(model.state_dict().keys()) == keys
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    # one fixed synthetic batch per rank (different frames on every rank)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    imgs = torch.randn(cfg.B, cfg.N, 3, *cfg.final_dim, device=dev, generator=g)
    cal = {k: torch.from_numpy(v).to(dev) for k, v in S.make_calibration(cfg, 1234 + rank).items()}
    X, Y = int(model.nx[0]), int(model.nx[1])
    target = torch.randint(0, 4, (cfg.B, X, Y), device=dev, generator=g)
    losses, t_steps = [], []
    for step in range(a.steps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = net(imgs, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
        loss = Fnn.cross_entropy(out, target)
        loss.backward()
        nn.utils.clip_grad_norm_(net.parameters(), 5.0)    # as train.py:64
        opt.step()
        torch.cuda.synchronize(); t_steps.append(time.perf_counter() - t0)
        losses.append(loss.detach().item())
    ms = 1e3 * sorted(t_steps[3:])[len(t_steps[3:]) // 2]
    # (a) identical parameters on every rank
    flat = torch.cat([p.detach().flatten() for p in model.parameters() if p.requires_grad])
    same = True
    if world > 1:
        ref = flat.clone(); dist.broadcast(ref, 0)
        ok = torch.tensor([float(torch.equal(ref, flat))], device=dev); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
        lt = torch.tensor([losses[0], losses[-1], ms], device=dev); dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        first, last, ms = (float(v) for v in lt)
    else:
        first, last = losses[0], losses[-1]
    if rank == 0:
        print("ddp stand-in: %d GPU(s), batch %d/GPU, %.2f ms/step (slowest rank) = %.0f frames/s; loss %.4f -> %.4f; "
              "parameters identical across ranks: %s; fused softmax path: %s"
              % (world, cfg.B, ms, world * cfg.B / ms * 1e3, first, last, same, patch._cache(model).fuse_softmax))
        assert same and last < first
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
