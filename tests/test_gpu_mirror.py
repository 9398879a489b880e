"""GPU: the winner block's second destination (fot_set_result_mirror) -- the gather of the query-sharded sweep without
a collective call.  On one GPU the mirror is a second local buffer; across GPUs it is a slice of the
root's buffer mapped over NVLink peer memory (bench.py --gpus N exercises that and verifies every rank's slice)."""
import ctypes as C

import numpy as np
import pytest

from tests import scenarios

pytestmark = pytest.mark.gpu


def test_winner_kernel_writes_the_mirror_block_and_publishes_its_sequence_number():
    import torch
    import bench
    from integrated_path_planning_b200 import BatchFrenetPlanner, CubicSpline2D, DeviceBatch, WinnerBlock, _lib
    spline, frenet, dyn = bench.make_queries(21000, 300)
    pl = BatchFrenetPlanner(CubicSpline2D(*scenarios.STRAIGHT_60), **scenarios.S1_KNOBS)
    batch = DeviceBatch(pl, frenet, 6.0, dyn, _lib.FOT_DYN_SINGLE)
    eng = pl.engine
    mirror = WinnerBlock(300, eng.n_t_max, device="cuda")
    mirror.buf.fill_(0x5a)
    flag = torch.zeros(2, dtype=torch.int32, device="cuda")
    m = _lib.FotResult()
    v = mirror.views
    m.best_idx, m.best_cost, m.stats = v["best_idx"].data_ptr(), v["best_cost"].data_ptr(), v["stats"].data_ptr()
    m.winner_len, m.winner = v["winner_len"].data_ptr(), v["winner"].data_ptr()
    _lib.check(eng.lib.fot_set_result_mirror(eng._h, C.byref(m), C.c_void_p(flag.data_ptr())), "fot_set_result_mirror")
    try:
        for launch in (1, 2, 3):
            batch.launch(None)
            assert int(flag[0].item()) == launch                 # 1, 2, 3, ... behind each mirrored launch
        out = batch.out
        assert (out["best_idx"] >= 0).any() and (out["best_idx"] < 0).any()
        for key in ("best_idx", "winner_len", "stats"):
            assert torch.equal(v[key], out[key]), key
        assert torch.equal(v["best_cost"].view(torch.int64), out["best_cost"].view(torch.int64))
        keep = torch.arange(eng.n_t_max, device="cuda")[None, None, :] < out["winner_len"][:, None, None]
        assert torch.equal(torch.where(keep, v["winner"], 0.0).view(torch.int64), torch.where(keep, out["winner"], 0.0).view(torch.int64))
        # rows beyond winner_len are never written, in either copy
        untouched = v["winner"].view(torch.uint8).reshape(300, -1)[(out["best_idx"] < 0)]
        assert bool((untouched == 0x5a).all())
    finally:
        _lib.check(eng.lib.fot_set_result_mirror(eng._h, None, None), "fot_set_result_mirror")
    flag.zero_()
    batch.launch(None)
    assert int(flag[0].item()) == 0                              # mirror off: nothing published


def test_peer_alloc_export_and_await_on_one_device():
    """fot_peer_alloc / fot_peer_await without a second process: the buffer is zeroed, the handle is 64 bytes, and the
    wait kernel passes once every flag has reached the sequence number (and reports a time-out otherwise... not waited
    for here: a 2 s spin is the designed bound)."""
    import torch
    from integrated_path_planning_b200 import _lib
    from integrated_path_planning_b200.batch import _RawCuda
    lib = _lib.load()
    ptr = C.c_void_p()
    handle = C.create_string_buffer(64)
    _lib.check(lib.fot_peer_alloc(0, 4096, C.byref(ptr), handle), "fot_peer_alloc")
    try:
        t = torch.as_tensor(_RawCuda(ptr.value, 4096), device="cuda:0")
        assert int(t.sum().item()) == 0
        flags = t[:16].view(torch.int32)
        flags[:3] = torch.tensor([5, 7, 6], dtype=torch.int32, device="cuda:0")
        _lib.check(lib.fot_peer_await(0, None, C.c_void_p(ptr.value), 3, 5, C.c_void_p(ptr.value + 64)), "fot_peer_await")
        torch.cuda.synchronize()
        assert int(t[64:68].view(torch.int32).item()) == 0       # no time-out
    finally:
        _lib.check(lib.fot_peer_free(ptr), "fot_peer_free")
