"""Host side of the sweep: planner-constant tables, the fot handle, one launch per batch.

Everything numeric that the reference computes once per planner or once per call OUTSIDE the
candidate loops is built here with the reference's own NumPy expressions, so the kernel starts
from bit-identical grids, time tables and thresholds:
  horizon grid      frenet_planner.py:397-398      lateral grid   :419-420
  speed grid        :410-413                       brake ladder   :473-485
  TimeCache inverses :599-616 (numpy.linalg.inv)   collision radii :1172-1175
The candidate sweep itself runs only on the GPU (csrc/fot_kernels.cuh) through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .spline import spline_tables
from .types import SERIES

BRAKE_T_MIN = 0.5
BRAKE_T_STEP = 0.5


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def horizon_grid(min_t, max_t, dt) -> np.ndarray:
    n_ti = int((max_t - min_t) / dt + 1e-9)
    return np.asarray(min_t + np.arange(n_ti + 1) * dt, dtype=np.float64)


def lateral_grid(d_road_w, max_road_width) -> np.ndarray:
    n_side = int(max_road_width / d_road_w + 1e-9)
    return np.asarray(np.arange(-n_side, n_side + 1) * d_road_w, dtype=np.float64)


def speed_grid(target_speed, d_t_s) -> np.ndarray:
    """One query's terminal-speed grid (frenet_planner.py:410-413)."""
    n_down = int(target_speed / d_t_s + 1e-9)
    if n_down < 0:
        raise ValueError("target_speed too negative for the speed grid")
    tv = target_speed - np.arange(n_down + 1) * d_t_s
    if tv[-1] > 1e-9:
        tv = np.append(tv, 0.0)
    return np.asarray(tv, dtype=np.float64)


def speed_grid_batch(target_speed: np.ndarray, d_t_s: float):
    """Vectorised speed_grid: returns (v_grid [n_q, n_v_max], n_v [n_q]); same elementwise
    arithmetic as the scalar version."""
    target = np.asarray(target_speed, dtype=np.float64).reshape(-1)
    if target.size > 1 and target.min() == target.max():
        # one target speed for the whole batch (the usual case): one row, repeated -- same arithmetic per element
        grid, n_v = speed_grid_batch(target[:1], d_t_s)
        return np.repeat(grid, target.size, axis=0), np.repeat(n_v, target.size)
    n_down = np.trunc(target / d_t_s + 1e-9).astype(np.int64)
    if np.any(n_down < 0):
        raise ValueError("target_speed too negative for the speed grid")
    width = int(n_down.max()) + 2
    k = np.arange(width)
    grid = target[:, None] - k[None, :] * d_t_s
    last = grid[np.arange(target.size), n_down]
    extra = last > 1e-9
    n_v = (n_down + 1 + extra).astype(np.int32)
    cols = np.arange(width)[None, :]
    grid = np.where(cols <= n_down[:, None], grid, 0.0)
    n_v_max = int(n_v.max())
    return np.ascontiguousarray(grid[:, :n_v_max]), n_v


def time_table(T, dt):
    """n_steps and the two inverse boundary matrices of one horizon (frenet_planner.py:593-616)."""
    n_steps = int(round(T / dt))
    ts = float(T)
    quartic = np.array([[3.0 * ts ** 2, 4.0 * ts ** 3],
                        [6.0 * ts, 12.0 * ts ** 2]])
    quintic = np.array([[ts ** 3, ts ** 4, ts ** 5],
                        [3.0 * ts ** 2, 4.0 * ts ** 3, 5.0 * ts ** 4],
                        [6.0 * ts, 12.0 * ts ** 2, 20.0 * ts ** 3]])
    return n_steps, np.linalg.inv(quartic), np.linalg.inv(quintic)


@dataclass
class SweepResult:
    """Host copy of fot_result_t for n_q queries."""
    best_idx: np.ndarray      # [n_q] int32
    best_cost: np.ndarray     # [n_q]
    stats: np.ndarray         # [n_q, 8] int32, slot order _lib.STAT_KEYS
    winner_len: np.ndarray    # [n_q] int32
    winner: np.ndarray        # [n_q, 15, n_t_max]
    cand_cat: Optional[np.ndarray] = None    # [n_q, stride] uint8
    cand_cost: Optional[np.ndarray] = None   # [n_q, stride]
    n_cand: Optional[np.ndarray] = None      # [n_q] candidates generated per query
    kernel_ms: float = float("nan")

    def series(self, q: int = 0):
        """The 15 winner sequences of query q truncated to their kept length, or None."""
        if self.best_idx[q] < 0:
            return None
        n = min(int(self.winner_len[q]), self.winner.shape[2])      # winner_samples = k: the first k samples only
        return {name: self.winner[q, j, :n].copy() for j, name in enumerate(SERIES)}


class SweepEngine:
    """Owns one fot handle (one device).  Built from the planner's constructor knobs."""

    def __init__(self, reference_path, *, max_speed, dt, d_road_w, max_road_width, robot_radius,
                 obstacle_radius, min_t, max_t, d_t_s, k_j, k_t, k_d, k_s_dot, k_lat, k_lon,
                 chance_epsilon=0.0, collision_margin_inflation=1.0, footprint=None, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        self.dt = dt
        self.d_t_s = d_t_s
        sp = spline_tables(reference_path)

        T = horizon_grid(min_t, max_t, dt)
        tabs = [time_table(t, dt) for t in T]
        d_grid = lateral_grid(d_road_w, max_road_width)
        if T.size == 0 or d_grid.size == 0:
            raise ValueError("empty horizon or lateral grid")
        n_total = int(round(max_t / dt)) + 1
        Tb_all = np.arange(BRAKE_T_MIN, min_t - 1e-9, BRAKE_T_STEP)
        btabs = [(tb, *time_table(tb, dt)) for tb in Tb_all]
        btabs = [b for b in btabs if n_total - (b[1] + 1) >= 0]      # :483-485 n_pad < 0 -> skipped

        if footprint is None:
            ego_radius = robot_radius
            offsets = np.zeros(0)
        else:
            ego_radius = footprint.radius
            offsets = np.asarray(footprint.offsets, dtype=np.float64).reshape(-1)
            if offsets.size > _lib.FOT_MAX_CIRCLES or offsets.size == 0:
                raise ValueError(f"footprint with {offsets.size} circles unsupported (1..{_lib.FOT_MAX_CIRCLES})")
        inflated_radius = max(ego_radius + obstacle_radius, 1e-6)
        dyn_radius = inflated_radius * collision_margin_inflation
        sq_rubicon = inflated_radius ** 2
        sq_rubicon_dyn = dyn_radius ** 2

        cfg = _lib.FotConfig()
        cfg.dt, cfg.max_speed, cfg.max_road_width = dt, max_speed, max_road_width
        cfg.k_j, cfg.k_t, cfg.k_d, cfg.k_s_dot, cfg.k_lat, cfg.k_lon = k_j, k_t, k_d, k_s_dot, k_lat, k_lon
        cfg.collide_r2, cfg.collide_r2_single, cfg.chance_epsilon = sq_rubicon, sq_rubicon_dyn, chance_epsilon
        for i, o in enumerate(offsets):
            cfg.circle_offsets[i] = float(o)
        cfg.n_circles = int(offsets.size)
        cfg.n_T, cfg.n_d, cfg.n_B, cfg.n_total, cfg.nx = len(T), len(d_grid), len(btabs), n_total, len(sp["knots"])

        f64 = lambda seq, shape: np.ascontiguousarray(np.array(seq, dtype=np.float64).reshape(shape))
        i32 = lambda seq: np.ascontiguousarray(np.array(seq, dtype=np.int32).reshape(-1))
        self._keep = {                                     # keep host arrays alive during create
            "T": f64(T, (-1,)), "n_steps": i32([t[0] for t in tabs]),
            "inv4": f64([t[1] for t in tabs], (-1,)), "inv5": f64([t[2] for t in tabs], (-1,)),
            "Tb": f64([b[0] for b in btabs], (-1,)), "n_steps_b": i32([b[1] for b in btabs]),
            "inv4b": f64([b[2] for b in btabs], (-1,)), "inv5b": f64([b[3] for b in btabs], (-1,)),
            "d_grid": f64(d_grid, (-1,)), **sp,
        }
        tb = _lib.FotTables()
        for name, _ in _lib.FotTables._fields_:
            arr = self._keep[name]
            ctype = _lib.c_int32_p if arr.dtype == np.int32 else _lib.c_double_p
            setattr(tb, name, arr.ctypes.data_as(ctype) if arr.size else None)
        self.n_T, self.n_d, self.n_B = cfg.n_T, cfg.n_d, cfg.n_B
        self.n_circles = int(cfg.n_circles)
        self.T, self.d_grid, self.Tb = self._keep["T"], self._keep["d_grid"], self._keep["Tb"]
        self.n_steps, self.n_steps_b, self.n_total = self._keep["n_steps"], self._keep["n_steps_b"], n_total
        self._pinned = {}
        self._grid_cache = None
        self._h = C.c_void_p()
        _lib.check(self.lib.fot_create(C.byref(cfg), C.byref(tb), self.device, C.byref(self._h)), "fot_create")
        self.n_t_max = self.lib.fot_n_t_max(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fot_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------
    def candidate_counts(self, frenet: np.ndarray, n_v: np.ndarray) -> np.ndarray:
        has_brake = (frenet[:, 1] > 0.1) & (self.n_B > 0)
        return (self.n_T * n_v.astype(np.int64) * self.n_d + has_brake * self.n_B).astype(np.int64)

    def points_per_query(self, frenet: np.ndarray, n_v: np.ndarray) -> np.ndarray:
        """Un-truncated trajectory samples per query (the E of SURVEY.md section 8d divides by this)."""
        per_T = (self.n_steps.astype(np.int64) + 1).sum()
        has_brake = (frenet[:, 1] > 0.1) & (self.n_B > 0)
        return per_T * n_v.astype(np.int64) * self.n_d + has_brake * self.n_B * self.n_total

    def run_host(self, frenet, target_speed, limits, stop_dist=None, static=None, dyn=None,
                 dyn_mode=_lib.FOT_DYN_NONE, static_per_query=False, want_candidates=False,
                 winner_samples: int = 0) -> SweepResult:
        """One fot_plan_batch_host call.  Shapes: frenet [n_q,6], target_speed [n_q], limits [n_q,4],
        stop_dist [n_q] (NaN = none), static [M,2] or [n_q,M,2], dyn [n_q,S,P,T,2].
        winner_samples = k > 0: only the first k samples of each winner series come back (`winner` is
        [n_q,15,k]); `fetch_winners` reads full series of the last call from the device."""
        f64c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        frenet = f64c(frenet).reshape(-1, 6)
        n_q = frenet.shape[0]
        target = f64c(np.broadcast_to(np.asarray(target_speed, dtype=np.float64), (n_q,)))
        limits = f64c(np.broadcast_to(np.asarray(limits, dtype=np.float64), (n_q, 4)))
        stop = f64c(np.full(n_q, np.nan) if stop_dist is None else np.broadcast_to(np.asarray(stop_dist, dtype=np.float64), (n_q,)))
        # one target speed for the whole batch (the usual case): the grid of the previous call with the same target
        key = (float(target_speed), n_q) if np.ndim(target_speed) == 0 else None
        if key is not None and self._grid_cache is not None and self._grid_cache[0] == key:
            v_grid, n_v = self._grid_cache[1]
        else:
            v_grid, n_v = speed_grid_batch(target, self.d_t_s)
            self._grid_cache = (key, (v_grid, n_v)) if key is not None else None
        b = _lib.FotBatch()
        b.n_q, b.n_v_max = n_q, v_grid.shape[1]
        b.frenet, b.target_speed, b.limits, b.stop_dist = _ptr(frenet), _ptr(target), _ptr(limits), _ptr(stop)
        b.v_grid, b.n_v = _ptr(v_grid), _ptr(n_v)
        keep = [frenet, target, limits, stop, v_grid, n_v]
        if static is not None and np.size(static) > 0:
            static = f64c(static)
            b.n_static = static.shape[-2]
            b.static_per_query = int(bool(static_per_query))
            b.static_obs = _ptr(static)
            keep.append(static)
        if dyn_mode != _lib.FOT_DYN_NONE:
            dyn = f64c(dyn)
            if dyn.ndim != 5 or dyn.shape[0] != n_q or dyn.shape[-1] != 2:
                raise ValueError(f"dyn must be [n_q,S,P,T,2], got {dyn.shape}")
            b.S, b.P, b.T_obs = dyn.shape[1], dyn.shape[2], dyn.shape[3]
            b.dyn, b.dyn_mode = _ptr(dyn), dyn_mode
            keep.append(dyn)
        k_head = min(int(winner_samples), self.n_t_max) if winner_samples and winner_samples > 0 else 0
        res = self._result_buffers(n_q, want_candidates, int(b.n_v_max), k_head)
        r = _lib.FotResult()
        r.winner_samples = k_head
        r.best_idx, r.best_cost, r.stats = _ptr(res.best_idx), _ptr(res.best_cost), _ptr(res.stats)
        r.winner_len, r.winner = _ptr(res.winner_len), _ptr(res.winner)
        if want_candidates:
            r.cand_cat, r.cand_cost, r.cand_stride = _ptr(res.cand_cat), _ptr(res.cand_cost), res.cand_cat.shape[1]
        _lib.check(self.lib.fot_plan_batch_host(self._h, C.byref(b), C.byref(r)), "fot_plan_batch_host")
        res.n_cand = self.candidate_counts(frenet, n_v)
        res.kernel_ms = float(self.lib.fot_last_kernel_ms(self._h))
        return res

    def fetch_winners(self, q0: int, n: int) -> np.ndarray:
        """Full winner series [n,15,n_t_max] of queries q0..q0+n-1 of the last run_host call (fot_fetch_winners)."""
        out = np.empty((n, _lib.FOT_N_SERIES, self.n_t_max), dtype=np.float64)
        _lib.check(self.lib.fot_fetch_winners(self._h, int(q0), int(n), _ptr(out)), "fot_fetch_winners")
        return out

    def reload_options(self) -> None:
        """Re-read the FOT_* tuning environment for this handle (tests / tuning; the planning calls never read it)."""
        _lib.check(self.lib.fot_reload_options(self._h), "fot_reload_options")

    def _result_buffers(self, n_q: int, want_candidates: bool, n_v_max: int, k_head: int = 0) -> SweepResult:
        """Result arrays in pinned host memory (torch is only the allocator), cached per batch size so
        the device->host copies are true async DMAs.  The arrays of a returned SweepResult stay valid
        until the next call on this engine with the same batch size."""
        import torch
        stride = self.lib.fot_candidate_count(self._h, n_v_max, 1) if want_candidates else 0
        key = (n_q, stride, k_head)
        cached = self._pinned.get(key)
        if cached is None:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
            cached = dict(best_idx=pin((n_q,), torch.int32), best_cost=pin((n_q,), torch.float64),
                          stats=pin((n_q, _lib.FOT_N_STATS), torch.int32), winner_len=pin((n_q,), torch.int32),
                          winner=pin((n_q, _lib.FOT_N_SERIES, k_head or self.n_t_max), torch.float64))
            if want_candidates:
                cached["cand_cat"] = pin((n_q, stride), torch.uint8)
                cached["cand_cost"] = pin((n_q, stride), torch.float64)
            if len(self._pinned) > 8:
                self._pinned.clear()
            self._pinned[key] = cached
        return SweepResult(**cached)

    def run_device(self, batch: "_lib.FotBatch", result: "_lib.FotResult", stream=None) -> None:
        """fot_plan_batch_device with caller-owned device pointers (torch tensors as buffers)."""
        _lib.check(self.lib.fot_plan_batch_device(self._h, C.byref(batch), C.byref(result),
                                                  C.c_void_p(stream) if stream else None),
                   "fot_plan_batch_device")

    def launch_stage_ms(self, back: int = 0):
        """(prepass, sweep, winner) device ms of the `back`-th most recent launch."""
        ms = (C.c_float * 3)()
        _lib.check(self.lib.fot_launch_stage_ms(self._h, int(back), C.byref(ms)), "fot_launch_stage_ms")
        return tuple(float(v) for v in ms)

    def last_kernel_ms(self) -> float:
        return float(self.lib.fot_last_kernel_ms(self._h))
