"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): batch sharding covers the
batch exactly once, and the benchmark's aggregate = sum(units) / max(time)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import sharding, synth


@pytest.mark.parametrize("total,world", [(64, 1), (64, 2), (256, 8), (10, 4), (3, 8)])
def test_shard_range_partitions_the_batch(total, world):
    seen = []
    for r in range(world):
        seen += list(sharding.shard_range(total, r, world))
    assert seen == list(range(total))
    sizes = [len(sharding.shard_range(total, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(total, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, L, H, W = 6, 3, 8, 8
        x = synth.make_layers(B, L, H, W, "W", seed=0)            # every rank builds the same global batch
        mine = sharding.shard_range(B, rank, world)
        shard = x[mine.start:mine.stop]
        # each rank's shard checksum, gathered: together they must reproduce the global checksum
        local = torch.tensor([shard.double().sum().item(), float(len(mine))], dtype=torch.float64)
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        total = sum(g[0].item() for g in gathered)
        count = sum(g[1].item() for g in gathered)
        units, ms, thr = sharding.aggregate_throughput(len(mine) * L * H * W, 10.0 * (rank + 1))
        q.put((rank, abs(total - x.double().sum().item()) < 1e-9, count == B, units, ms, thr))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharding_and_aggregate():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, checksum_ok, count_ok, units, ms, thr in res:
        assert checksum_ok and count_ok
        assert units == 6 * 3 * 8 * 8            # sum of units over ranks
        assert ms == 20.0                        # max over ranks (rank 1 reported 20 ms)
        assert abs(thr - units / 0.020) < 1e-6


def test_aggregate_without_process_group():
    units, ms, thr = sharding.aggregate_throughput(1000.0, 4.0)
    assert (units, ms) == (1000.0, 4.0) and abs(thr - 250000.0) < 1e-9
