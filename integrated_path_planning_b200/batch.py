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
                   want_candidates: bool = False, winner_samples: int = 0) -> SweepResult:
        """frenet_states [n_q,6]; target_speed scalar or [n_q]; dynamic_obstacles [n_q,P,T,2] or
        distribution [n_q,S,P,T,2]; static_obstacles [M,2] (shared) or [n_q,M,2]; limits [n_q,4]
        overrides constraint_overrides; max_stop_distance scalar / [n_q] (NaN = none); winner_samples = k > 0
        reads back only the first k samples of each winner series (`engine.fetch_winners` for the rest)."""
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
                                    static_per_query=per_query, want_candidates=want_candidates,
                                    winner_samples=winner_samples)


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
        # the whole winner block in one contiguous buffer (what the sharded sweep gathers with one collective)
        self.block = WinnerBlock(n_q, eng.n_t_max, device=dev)
        self.out = self.block.views
        r = _lib.FotResult()
        r.best_idx, r.best_cost, r.stats = self.out["best_idx"].data_ptr(), self.out["best_cost"].data_ptr(), self.out["stats"].data_ptr()
        r.winner_len, r.winner = self.out["winner_len"].data_ptr(), self.out["winner"].data_ptr()
        self.result = r

    def launch(self, stream: Optional[int] = None) -> None:
        """Enqueue prepass + sweep + winner kernels on `stream` (raw cudaStream_t).  None: torch's current stream
        of the batch's device -- ordered behind whatever torch work produced the input tensors -- synchronised
        before returning."""
        if stream is None:
            import torch
            cur = torch.cuda.current_stream(self.out["best_idx"].device)
            self.engine.run_device(self.batch, self.result, cur.cuda_stream or None)
            cur.synchronize()
            return
        self.engine.run_device(self.batch, self.result, stream)

    def launch_to_host(self, host_out: dict, stream: Optional[int] = None, winner_samples: int = 0) -> None:
        """fot_plan_batch_device_to_host: sweep the (device-resident) batch and deliver the winners into the host
        arrays of `host_out` (keys of `self.out`; torch CPU tensors or NumPy arrays, ideally page-locked): the
        queries run in ranges and each range's winners are copied back while the next range is swept.  `stream`:
        the raw cudaStream_t on which the batch's inputs become ready (None: they are ready).  `winner_samples = k > 0`:
        only the first k samples of every winner series come back (`host_out["winner"]` is [n_q, 15, k]) -- what a
        closed-loop caller consumes (integrated_simulator.py:663), 240 B instead of 6.1 KB per query at k = 2."""
        import ctypes as C
        ptr = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        r = _lib.FotResult()
        r.best_idx, r.best_cost, r.stats = ptr(host_out["best_idx"]), ptr(host_out["best_cost"]), ptr(host_out["stats"])
        r.winner_len, r.winner = ptr(host_out["winner_len"]), ptr(host_out["winner"])
        r.winner_samples = int(winner_samples)
        _lib.check(self.engine.lib.fot_plan_batch_device_to_host(self.engine._h, C.byref(self.batch), C.byref(r),
                                                                 C.c_void_p(stream) if stream else None),
                   "fot_plan_batch_device_to_host")

    def dense_evals(self) -> int:
        """Densely credited point-vs-obstacle tests of this batch."""
        pts = self.engine.points_per_query(self.frenet_host, self.n_v_host)
        n_circ = max(1, self.engine.n_circles)
        per_point = self.batch.S * self.batch.P + self.batch.n_static
        return int(pts.sum()) * n_circ * int(per_point)


class WinnerBlock:
    """A rank's winner block for `cap` queries in ONE contiguous buffer: best_idx | winner_len | stats | best_cost |
    winner, each section 256-byte aligned (what `gather_winners` moves with one collective, SURVEY.md section 8e:
    about 6.2 KB per query).  `views[key]` are typed tensors into the buffer; the kernels write straight into them."""

    def __init__(self, cap: int, n_t_max: int, device=None, sections=None):
        import torch
        self.cap = int(cap)
        self.sections = sections or (("best_idx", torch.int32, ()), ("winner_len", torch.int32, ()),
                                     ("stats", torch.int32, (_lib.FOT_N_STATS,)), ("best_cost", torch.float64, ()),
                                     ("winner", torch.float64, (_lib.FOT_N_SERIES, int(n_t_max))))
        self.offsets, off = {}, 0
        for key, dt, shape in self.sections:
            n = self.cap * int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
            self.offsets[key] = (off, n)
            off = (off + n + 255) // 256 * 256
        self.nbytes = max(off, 256)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.views = self.unpack(self.buf)

    def unpack(self, buf):
        """Typed views of a buffer laid out like this block (leading dims of `buf` are kept: [world, nbytes] ->
        [world, cap, ...])."""
        lead = tuple(buf.shape[:-1])
        out = {}
        for key, dt, shape in self.sections:
            off, n = self.offsets[key]
            out[key] = buf[..., off:off + n].contiguous().view(dt).reshape(lead + (self.cap,) + tuple(shape)) if lead \
                else buf[off:off + n].view(dt).reshape((self.cap,) + tuple(shape))
        return out


def gather_winners(out: dict, group=None, counts=None) -> dict:
    """The one collective of the sharded sweep: every rank's winner block (best_idx, best_cost, stats, winner_len,
    winner -- any dict of tensors with the queries on dim 0) packed into ONE contiguous buffer and moved by ONE
    `all_gather_into_tensor` over the process group (NCCL on GPUs, gloo in the CPU tests); returns the blocks of all
    ranks concatenated in rank order.

    Ranks may hold different numbers of queries (`shard_bounds` gives the last rank a shorter or empty block): every
    block is padded to the longest one and the padding is trimmed after the gather.  `counts` = queries per rank when
    the caller knows them (e.g. from `shard_bounds`); otherwise they are exchanged first (one extra, 8-byte
    collective).  A `WinnerBlock`'s views are gathered without the packing copy when its capacity already equals the
    longest block."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out
    world = dist.get_world_size(group)
    first = next(iter(out.values()))
    n_local, dev = int(first.shape[0]), first.device
    if counts is None:
        mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
        allc = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)
        counts = [int(c) for c in allc.cpu()]
    counts = [int(c) for c in counts]
    assert len(counts) == world and counts[dist.get_rank(group)] == n_local, (counts, n_local)
    per = max(counts)
    sections = tuple((k, t.dtype, tuple(t.shape[1:])) for k, t in out.items())
    blk = WinnerBlock(per, 0, device=dev, sections=sections)
    own = getattr(first, "_base", None)
    zero_copy = per == n_local and own is not None and own.dtype == torch.uint8 and own.numel() == blk.nbytes and \
        all(t._base is own and t.data_ptr() == own.data_ptr() + blk.offsets[k][0] for k, t in out.items())
    if zero_copy:
        send = own
    else:
        send = blk.buf
        send.zero_()
        for k, t in out.items():
            blk.views[k][:n_local].copy_(t)
    gathered = torch.empty((world, blk.nbytes), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), send, group=group)
    parts = blk.unpack(gathered)
    if all(c == per for c in counts):
        return {k: v.reshape((world * per,) + tuple(v.shape[2:])) for k, v in parts.items()}
    return {k: torch.cat([v[r, :counts[r]] for r in range(world)], dim=0) for k, v in parts.items()}


class _RawCuda:
    """A raw device pointer as something torch can wrap (`torch.as_tensor(obj, device=...)`)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """The one gather of the query-sharded sweep without a collective call: every rank's winner block lands in the
    ROOT's memory, stored there by that rank's own `fot_winner` kernel through NVLink peer memory
    (fot_set_result_mirror), while the rank's next sweep is already running.

    The root owns `depth` buffers of `world x WinnerBlock(cap)` bytes (fot_peer_alloc) and a flag word per rank; the
    other ranks map them (CUDA IPC handle -> fot_peer_open; the 64-byte handles travel through torch.distributed).
    Step k of every rank is mirrored into buffer k % depth, slice `rank`; behind it the rank publishes k + 1 in its flag
    word.  `await_step(k)` on the root enqueues a device-side wait for all flags; `views(k)` are typed tensors
    [world, cap, ...] into the buffer.  One process per GPU, one handle per process.
    """

    def __init__(self, engine: SweepEngine, cap: int, group=None, root: int = 0, depth: int = 2):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.engine, self.lib, self.root, self.depth = engine, engine.lib, int(root), int(depth)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device("cuda", engine.device)
        self.layout = WinnerBlock(cap, engine.n_t_max, device="meta")             # offsets only
        self.nbytes = self.layout.nbytes
        self.flag_bytes = 256 * ((4 * self.world + 255) // 256)
        total = self.depth * self.world * self.nbytes + self.flag_bytes + 256
        handle = torch.zeros(64, dtype=torch.uint8)
        self._base = C.c_void_p()
        self._owned = self.rank == self.root
        if self._owned:
            buf = C.create_string_buffer(64)
            _lib.check(self.lib.fot_peer_alloc(engine.device, total, C.byref(self._base), buf), "fot_peer_alloc")
            handle = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        h_dev = handle.to(self.dev)
        dist.broadcast(h_dev, src=dist.get_global_rank(group, self.root) if group is not None else self.root, group=group)
        if not self._owned:
            raw = bytes(h_dev.cpu().numpy().tobytes())
            _lib.check(self.lib.fot_peer_open(engine.device, raw, C.byref(self._base)), "fot_peer_open")
        self.base = int(self._base.value)
        self.flags_ptr = self.base + self.depth * self.world * self.nbytes
        self.err_ptr = self.flags_ptr + self.flag_bytes
        self.step = 0

    def _slice_ptr(self, k: int, rank: int) -> int:
        return self.base + ((k % self.depth) * self.world + rank) * self.nbytes

    def attach(self, k: Optional[int] = None) -> None:
        """The engine's next fot_plan_batch_device launch mirrors its winner block into buffer k % depth (default: the
        running step counter, which then advances)."""
        import ctypes as C
        if k is None:
            k = self.step
            self.step += 1
        p = self._slice_ptr(k, self.rank)
        m = _lib.FotResult()
        off = self.layout.offsets
        m.best_idx, m.winner_len, m.stats = p + off["best_idx"][0], p + off["winner_len"][0], p + off["stats"][0]
        m.best_cost, m.winner = p + off["best_cost"][0], p + off["winner"][0]
        _lib.check(self.lib.fot_set_result_mirror(self.engine._h, C.byref(m), C.c_void_p(self.flags_ptr + 4 * self.rank)),
                   "fot_set_result_mirror")

    def detach(self) -> None:
        _lib.check(self.lib.fot_set_result_mirror(self.engine._h, None, None), "fot_set_result_mirror")

    def await_step(self, seq: int, stream: Optional[int] = None) -> None:
        """Root: enqueue a wait until every rank has published launch number `seq` (1-based count of its mirrored
        launches)."""
        import ctypes as C
        assert self._owned, "await_step is a root-side call"
        _lib.check(self.lib.fot_peer_await(self.engine.device, C.c_void_p(stream) if stream else None, C.c_void_p(self.flags_ptr),
                                           self.world, int(seq), C.c_void_p(self.err_ptr)), "fot_peer_await")

    def views(self, k: int) -> dict:
        """Root: the gathered blocks of step k as tensors [world, cap, ...] (views into the peer buffer)."""
        import torch
        assert self._owned
        raw = torch.as_tensor(_RawCuda(self._slice_ptr(k, 0), self.world * self.nbytes), device=self.dev)
        return self.layout.unpack(raw.view(self.world, self.nbytes))

    def timed_out(self) -> bool:
        import torch
        assert self._owned
        return bool(torch.as_tensor(_RawCuda(self.err_ptr, 4), device=self.dev).view(torch.int32).item())

    def close(self) -> None:
        if self._base is not None and self._base.value:
            self.detach()
            if self._owned:
                self.lib.fot_peer_free(self._base)
            else:
                self.lib.fot_peer_close(self._base)
            self._base = None
