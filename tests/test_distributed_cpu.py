"""CPU, world_size 2 and 3 over gloo: the one collective of the sharded sweep (gather of the packed winner block),
including uneven and empty shards (`shard_bounds` gives the last rank a shorter or empty block)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from integrated_path_planning_b200 import WinnerBlock, gather_winners, shard_bounds

N_T = 5


def _local_block(lo, hi, as_block):
    ids = torch.arange(lo, hi)
    n = hi - lo
    vals = {"best_idx": (ids * 3 % 7).to(torch.int32) - 1, "winner_len": (ids % N_T).to(torch.int32),
            "stats": torch.stack([ids + k for k in range(8)], dim=1).to(torch.int32) if n else torch.zeros((0, 8), dtype=torch.int32),
            "best_cost": ids.to(torch.float64) * 0.5,
            "winner": (ids.to(torch.float64)[:, None, None] + torch.arange(15)[None, :, None] * 0.01 +
                       torch.arange(N_T)[None, None, :] * 0.0001)}
    if not as_block:
        return vals
    blk = WinnerBlock(n, N_T)
    for k, v in vals.items():
        blk.views[k].copy_(v)
    return blk.views


def _worker(rank, world, port, n_q, q, as_block, know_counts):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_q, world, rank)
        counts = [shard_bounds(n_q, world, r)[1] - shard_bounds(n_q, world, r)[0] for r in range(world)] if know_counts else None
        got = gather_winners(_local_block(lo, hi, as_block), counts=counts)
        q.put((rank, {k: v.numpy() for k, v in got.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_q,as_block,know_counts", [(2, 10, True, True), (2, 11, True, False), (2, 11, False, True),
                                                            (3, 4, True, False), (3, 7, False, False)])
def test_gather_winners_gloo(world, n_q, as_block, know_counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + world * 131 + n_q) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_q, q, as_block, know_counts)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = {k: v.numpy() for k, v in _local_block(0, n_q, False).items()}
    for r in range(world):
        assert set(results[r]) == set(want)
        for k in want:
            assert results[r][k].shape == want[k].shape, (r, k, results[r][k].shape)
            assert np.array_equal(results[r][k], want[k]), (r, k)


def test_gather_winners_is_identity_without_process_group():
    out = {"best_idx": torch.arange(4, dtype=torch.int32)}
    assert gather_winners(out) is out


def test_winner_block_is_one_contiguous_buffer():
    blk = WinnerBlock(7, 51)
    assert blk.buf.dtype == torch.uint8 and blk.buf.is_contiguous()
    base = blk.buf.data_ptr()
    for key, (off, n) in blk.offsets.items():
        v = blk.views[key]
        assert v.data_ptr() == base + off and off % 256 == 0 and v.numel() * v.element_size() == n
        assert v.shape[0] == 7
    assert blk.views["winner"].shape == (7, 15, 51) and blk.views["stats"].shape == (7, 8)
    per_query = sum(n for _, n in blk.offsets.values()) / 7
    assert per_query == 4 + 4 + 32 + 8 + 15 * 51 * 8          # the 6.2 KB per query of SURVEY.md section 8e
