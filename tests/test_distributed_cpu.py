"""CPU, world_size 2 over gloo: the one collective of the sharded sweep (gather of winners)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from integrated_path_planning_b200 import gather_winners, shard_bounds


def _worker(rank, world, port, n_q, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_q, world, rank)
        ids = torch.arange(lo, hi)
        out = {"best_idx": (ids * 3 % 7).to(torch.int32), "best_cost": ids.to(torch.float64) * 0.5,
               "stats": torch.stack([ids, ids + 1], dim=1).to(torch.int32)}
        got = gather_winners(out)
        q.put((rank, {k: v.numpy() for k, v in got.items()}))
    finally:
        dist.destroy_process_group()


def test_gather_winners_gloo_world2():
    world, n_q = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_q, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(n_q)
    for r in range(world):
        assert np.array_equal(results[r]["best_idx"], (ids * 3 % 7).astype(np.int32))
        assert np.array_equal(results[r]["best_cost"], ids * 0.5)
        assert np.array_equal(results[r]["stats"], np.stack([ids, ids + 1], 1).astype(np.int32))


def test_gather_winners_is_identity_without_process_group():
    out = {"best_idx": torch.arange(4, dtype=torch.int32)}
    assert gather_winners(out) is out
