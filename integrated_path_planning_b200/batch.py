"""Batched planning: many independent queries per launch, sharded over GPUs.

A query is one reference `plan()` call (ego Frenet state + obstacle field + per-call knobs).
Queries share only the planner constants, so a batch shards across ranks with no traffic
during compute; `gather_winners` is the single collective.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .engine import SweepEngine, SweepResult, speed_grid_batch
from .planner import FrenetPlanner


def shard_bounds(n_q: int, world_size: int, rank: int):
    """Contiguous block of ceil(n_q / world_size) queries for `rank`."""
    per = (n_q + world_size - 1) // world_size
    lo = min(rank * per, n_q)
    return lo, min(lo + per, n_q)


class BatchFrenetPlanner(FrenetPlanner):
    """FrenetPlanner plus `plan_batch` (host arrays) and `plan_batch_device` (torch buffers)."""

    def plan_batch(self, frenet_states, target_speed, dynamic_obstacles=None, distribution=None,
                   static_obstacles=None, constraint_overrides=None, limits=None, max_stop_distance=None,
                   want_candidates: bool = False) -> SweepResult:
        """frenet_states [n_q,6]; target_speed scalar or [n_q]; dynamic_obstacles [n_q,P,T,2] or
        distribution [n_q,S,P,T,2]; static_obstacles [M,2] (shared) or [n_q,M,2]; limits [n_q,4]
        overrides constraint_overrides; max_stop_distance scalar / [n_q] (NaN = none)."""
        fs = np.asarray(frenet_states, dtype=np.float64).reshape(-1, 6)
        n_q = fs.shape[0]
        if limits is None:
            limits = self.resolve_limits(constraint_overrides)
        mode, dyn = _lib.FOT_DYN_NONE, None
        if distribution is not None and np.size(distribution) > 0:
            mode, dyn = _lib.FOT_DYN_DISTRIBUTION, np.asarray(distribution, dtype=np.float64)
        elif dynamic_obstacles is not None and np.size(dynamic_obstacles) > 0:
            d = np.asarray(dynamic_obstacles, dtype=np.float64)
            mode, dyn = _lib.FOT_DYN_SINGLE, d.reshape(n_q, 1, *d.shape[1:])
        per_query = static_obstacles is not None and np.ndim(static_obstacles) == 3
        return self.engine.run_host(fs, target_speed, limits, max_stop_distance, static_obstacles, dyn, mode,
                                    static_per_query=per_query, want_candidates=want_candidates)


class DeviceBatch:
    """A batch whose inputs and outputs live in torch CUDA tensors (buffers only): the timed
    region of `bench.py`'s device-resident leg calls `launch()` and nothing else."""

    def __init__(self, planner: FrenetPlanner, frenet_states, target_speed, dyn, dyn_mode, limits=None,
                 max_stop_distance=None, static_obstacles=None):
        import torch
        eng = planner.engine
        dev = torch.device("cuda", eng.device)
        fs = np.ascontiguousarray(frenet_states, dtype=np.float64).reshape(-1, 6)
        n_q = fs.shape[0]
        target = np.array(np.broadcast_to(np.asarray(target_speed, np.float64), (n_q,)))
        lim = planner.resolve_limits(None) if limits is None else limits
        lim = np.array(np.broadcast_to(np.asarray(lim, np.float64), (n_q, 4)))
        stop = np.full(n_q, np.nan) if max_stop_distance is None else \
            np.array(np.broadcast_to(np.asarray(max_stop_distance, np.float64), (n_q,)))
        v_grid, n_v = speed_grid_batch(target, eng.d_t_s)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.t = {"frenet": up(fs), "target": up(target), "limits": up(lim), "stop": up(stop),
                  "v_grid": up(v_grid), "n_v": up(n_v)}
        self.engine, self.n_q = eng, n_q
        self.n_v_host, self.frenet_host = n_v, fs
        b = _lib.FotBatch()
        b.n_q, b.n_v_max = n_q, v_grid.shape[1]
        b.frenet, b.target_speed, b.limits = self.t["frenet"].data_ptr(), self.t["target"].data_ptr(), self.t["limits"].data_ptr()
        b.stop_dist, b.v_grid, b.n_v = self.t["stop"].data_ptr(), self.t["v_grid"].data_ptr(), self.t["n_v"].data_ptr()
        if static_obstacles is not None and np.size(static_obstacles) > 0:
            st = np.ascontiguousarray(static_obstacles, dtype=np.float64)
            self.t["static"] = up(st)
            b.static_obs, b.n_static, b.static_per_query = self.t["static"].data_ptr(), st.shape[-2], int(st.ndim == 3)
        if dyn_mode != _lib.FOT_DYN_NONE:
            if isinstance(dyn, np.ndarray):
                dyn = up(np.ascontiguousarray(dyn, dtype=np.float64))
            assert dyn.dim() == 5 and dyn.shape[0] == n_q and dyn.shape[-1] == 2 and dyn.dtype == torch.float64
            self.t["dyn"] = dyn.contiguous()
            b.dyn, b.S, b.P, b.T_obs, b.dyn_mode = self.t["dyn"].data_ptr(), dyn.shape[1], dyn.shape[2], dyn.shape[3], dyn_mode
        self.batch = b
        self.out = {
            "best_idx": torch.empty(n_q, dtype=torch.int32, device=dev),
            "best_cost": torch.empty(n_q, dtype=torch.float64, device=dev),
            "stats": torch.empty(n_q, _lib.FOT_N_STATS, dtype=torch.int32, device=dev),
            "winner_len": torch.empty(n_q, dtype=torch.int32, device=dev),
            "winner": torch.empty(n_q, _lib.FOT_N_SERIES, eng.n_t_max, dtype=torch.float64, device=dev),
        }
        r = _lib.FotResult()
        r.best_idx, r.best_cost, r.stats = self.out["best_idx"].data_ptr(), self.out["best_cost"].data_ptr(), self.out["stats"].data_ptr()
        r.winner_len, r.winner = self.out["winner_len"].data_ptr(), self.out["winner"].data_ptr()
        self.result = r

    def launch(self, stream: Optional[int] = None) -> None:
        """Enqueue prepass + sweep + winner kernels on `stream` (raw cudaStream_t; None = the
        handle's stream, synchronised before returning)."""
        self.engine.run_device(self.batch, self.result, stream)

    def launch_to_host(self, host_out: dict, stream: Optional[int] = None) -> None:
        """fot_plan_batch_device_to_host: sweep the (device-resident) batch and deliver the winners into the host
        arrays of `host_out` (keys of `self.out`; torch CPU tensors or NumPy arrays, ideally page-locked): the
        queries run in ranges and each range's winners are copied back while the next range is swept.  `stream`:
        the raw cudaStream_t on which the batch's inputs become ready (None: they are ready)."""
        import ctypes as C
        ptr = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        r = _lib.FotResult()
        r.best_idx, r.best_cost, r.stats = ptr(host_out["best_idx"]), ptr(host_out["best_cost"]), ptr(host_out["stats"])
        r.winner_len, r.winner = ptr(host_out["winner_len"]), ptr(host_out["winner"])
        _lib.check(self.engine.lib.fot_plan_batch_device_to_host(self.engine._h, C.byref(self.batch), C.byref(r),
                                                                 C.c_void_p(stream) if stream else None),
                   "fot_plan_batch_device_to_host")

    def dense_evals(self) -> int:
        """Densely credited point-vs-obstacle tests of this batch."""
        pts = self.engine.points_per_query(self.frenet_host, self.n_v_host)
        n_circ = max(1, self.engine.n_circles)
        per_point = self.batch.S * self.batch.P + self.batch.n_static
        return int(pts.sum()) * n_circ * int(per_point)


def gather_winners(out: dict, group=None) -> dict:
    """The one collective of the sharded sweep: all_gather of each rank's winner block
    (best_idx, best_cost, stats, winner_len, winner) over the default process group
    (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out
    world = dist.get_world_size(group)
    gathered = {}
    for key, t in out.items():
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        gathered[key] = torch.cat(parts, dim=0)
    return gathered
