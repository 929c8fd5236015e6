"""world_size=2 gloo test of the N>1 host logic (sample sharding + max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lss2_multimodal_nu_b200 import shard, synthetic as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.env_world() == (rank, rank, world)
        lo, hi = shard.shard_range(17, rank, world)
        # ranks own disjoint, covering, contiguous sample ranges
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi))
        assert gathered[0][0] == 0 and gathered[-1][1] == 17
        assert all(gathered[i][1] == gathered[i + 1][0] for i in range(world - 1))
        # different ranks draw different synthetic calibrations
        cfg = S.config("tiny")
        cal = S.make_calibration(cfg, shard.rank_seed(1234, rank))
        sums = [None] * world
        dist.all_gather_object(sums, float(cal["post_trans"].sum()))
        assert len(set(sums)) == world
        # timing = max over ranks, throughput = all samples / slowest rank
        ms = 10.0 * (rank + 1)
        assert shard.max_over_ranks(ms) == 10.0 * world
        assert shard.gather_floats(float(rank) + 0.5) == [r + 0.5 for r in range(world)]
        tput = shard.aggregate_throughput(8 * 100, ms)
        assert abs(tput - (8 * 100 * world) / (10.0 * world * 1e-3)) < 1e-6
        out[rank] = tput
    finally:
        dist.destroy_process_group()


def test_world_size_two():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world and len(set(out.values())) == 1


def test_shard_range_covers_everything():
    for gb in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(gb, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
