"""N>1 path on CPU: world_size-2 gloo ranks shard the batch with no data-path collective."""
import os
import random
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adt_str_b200 import sharding
    from adt_str_b200.config import setting_1
    from adt_str_b200.planner import plan_batch
    from adt_str_b200.synthetic import make_bank, make_segments
    segs = make_segments(10, seed=1)
    lo, hi = sharding.shard_range(len(segs), rank, world)
    bank = make_bank(78, min_len=200, max_len=2000, seed=0)             # replicated bank: same seed on every rank
    plan = plan_batch(segs[lo:hi], setting_1(), bank, rng=sharding.rank_rng(1234, rank))
    audio_s = float(plan.wave_lengths.sum()) / 24000
    total = sharding.reduce_stats(audio_s, elapsed_ms=10.0 + rank)        # host-side stats only
    if rank == 0:
        torch.save({"total": total, "range": (lo, hi)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["range"] == (0, 5)
    audio_s, ms = got["total"]
    assert ms == 11.0 and audio_s > 2 * 2.56 * 4                         # sum of units, max of times


def test_shard_ranges_partition_everything():
    from adt_str_b200 import sharding
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    a, b = sharding.rank_rng(7, 0), sharding.rank_rng(7, 1)
    assert a.random() != b.random() and sharding.rank_rng(7, 0).random() == sharding.rank_rng(7, 0).random()
