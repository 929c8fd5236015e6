#!/usr/bin/env python
"""BASELINE.json config 3: the reference's BEV_TXT full training step under DDP on 1 / 2 / 4 / 8 B200s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/ddp_train_bevtxt.py [--steps 30] [--bsize 8] [--stock]

The model is the UNMODIFIED `src.model_BEV_TXT.BEV_TXT` (imported from /root/reference or its staged copy
oracle/_ref through oracle/ref_import.py; the EfficientNet package is the stand-in of oracle/shims/ -- the
reference's own Encoder / CamEncode / BevEncode / SceneUnder / heads run as written).  The step is the body of
train.py:49-65: forward, the reference's MultiLoss (src/tools.py:232-251), backward, clip_grad_norm_(5.0), Adam
(lr 1e-3, weight decay 1e-7, train.py:110-113), wrapped in DistributedDataParallel (one process per GPU, NCCL over
NVLink).  The lift-splat stage has no parameters: NCCL carries only the surrounding model's gradient all-reduce and
the stage shards by sample.  Without --stock the drop-in is installed (patch.install); with --stock
the reference's PyTorch lift-splat runs, as the baseline.  Frames are synthetic.

Checks: every rank holds identical parameters afterwards (the all-reduce worked), the loss is finite and went
down, the state_dict has no extra entries, and `ConfusionMatrix.reduce_from_all_processes` (src/tools.py:567-573,
the reference's one collective call site) sums the per-rank matrices.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402
from lss2_multimodal_nu_b200 import patch, synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--bsize", type=int, default=8)
    ap.add_argument("--stock", action="store_true", help="the reference's own PyTorch lift-splat (baseline)")
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tools = ref_import.load_module("tools")
    cfg = S.config("config2", B=a.bsize)
    torch.manual_seed(0)                                            # identical initial weights on every rank
    model = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf(), outC=4, cls="BEV_TXT",
                                 module="model_BEV_TXT", backbone=True).to(dev)
    keys = list(model.state_dict().keys())
    if not a.stock:
        patch.install(model)
    assert list(model.state_dict().keys()) == keys
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True) \
        if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-7)
    loss_args = types.SimpleNamespace(gpuid=local)                  # MultiLoss reads args.gpuid

    def batch(step):
        seed = 1000 * rank + step                                   # every rank its own frames (sample sharding)
        cal = {k: torch.from_numpy(v).to(dev) for k, v in S.make_calibration(cfg, seed).items()}
        g = torch.Generator(device=dev); g.manual_seed(seed)
        imgs = torch.randn(cfg.B, cfg.N, 3, *cfg.final_dim, device=dev, generator=g)
        binimgs = torch.randint(0, 4, (cfg.B, 200, 200), device=dev, generator=g)
        acts = torch.randint(0, 2, (cfg.B, 4), device=dev, generator=g).float()
        descs = torch.randint(0, 2, (cfg.B, 8), device=dev, generator=g).float()
        return imgs, cal, binimgs, acts, descs

    batches = [batch(s) for s in range(4)]
    losses = []

    def step(i):
        imgs, cal, binimgs, acts, descs = batches[i % len(batches)]
        opt.zero_grad()
        bev, act, desc = net(imgs, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
        loss = tools.MultiLoss(bev, act, desc, binimgs, acts, descs, loss_args)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        return loss

    net.train()
    for i in range(a.warmup):
        losses.append(float(step(i)))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        loss = step(a.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    losses.append(float(loss))
    ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)

    # ---- checks ----
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    assert torch.isfinite(flat).all()
    same = True
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, flat))
        ok = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    # the reference's one collective: the confusion matrix of the validation loop
    cm = tools.ConfusionMatrix(4)
    gt = torch.arange(4, device=dev).repeat(25 * (rank + 1))
    cm.update(gt, gt.flip(0))
    mine = cm.mat.clone()
    cm.reduce_from_all_processes()
    want = sum(torch.bincount(4 * torch.arange(4).repeat(25 * (r + 1)) + torch.arange(4).repeat(25 * (r + 1)).flip(0),
                              minlength=16).reshape(4, 4) for r in range(world))
    cm_ok = bool(torch.equal(cm.mat.cpu(), want)) and (world == 1 or not torch.equal(cm.mat, mine))
    if rank == 0:
        print(json.dumps({
            "workload": "model_BEV_TXT.BEV_TXT full training step (BEV + text head), batch %d/GPU, DDP x%d" % (cfg.B, world),
            "lift_splat": "reference PyTorch (stock)" if a.stock else "lss2_multimodal_nu_b200 drop-in",
            "n_gpus": world, "steps": a.steps, "ms_per_step": ms, "frames_per_s": cfg.B * world / (ms * 1e-3),
            "wall_ms_per_step": wall * 1e3 / a.steps, "loss_first": losses[0], "loss_last": losses[-1],
            "params_identical_across_ranks": same, "confusion_matrix_allreduce_ok": cm_ok,
            "prefetch_hits": patch._cache(model).prefetch_hits if not a.stock else None,
            "backbone": "oracle/shims stand-in EfficientNet (output contract only)"}))
    assert same and cm_ok and losses[-1] == losses[-1]
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
