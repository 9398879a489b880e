"""Batched closed-loop roll-out driver.

The reference runs its campaigns (Monte-Carlo seeds, sensitivity sweeps: examples/run_da_poc.py:179+,
examples/run_rq1b_sensitivity.py:64, examples/run_statistical_benchmark.py:243-261) as sequential loops of
`IntegratedSimulator.step()`.  This driver advances N independent simulations in lock-step and issues ONE
batched sweep per (re)planning attempt:

    replayed pedestrians -> observer (0.4 s sampling) -> CV prediction + t = 0 column   [device]
                         -> safety metrics (clearance, clearance ahead)                  [device]
                         -> fail-safe state machine: target speed, limits, stop room     [host, vectorised]
                         -> FrenetPlanner sweep, up to 3 escalation retries per step     [device]
                         -> ego update from the winner's second sample, or the adaptive emergency stop

It restates, per simulation, exactly what `IntegratedSimulator.step()` / `run()` do with a replayed
pedestrian source and the constant-velocity predictor (src/simulation/integrated_simulator.py:406-422,
:424-527, :529-653, :655-676, :678-747, :749-802, :842-892; src/core/state_machine.py:116-278;
src/pedestrian/observer.py:52-86; src/simulation/replay_source.py:31-111), so that a batch reproduces the
reference's trajectories (tests/test_gpu_rollout.py, golden roll-outs recorded from the unmodified
reference by tests/golden/make_golden_rollout.py).  Pedestrian ground truth stays outside (replay), as in
the reference's own pysocialforce-free source; static obstacles and the multi-circle footprint are not
wired into this driver yet.
"""
from __future__ import annotations

import math
import time
from typing import Dict, Optional

import numpy as np

from . import _lib
from .batch import BatchFrenetPlanner, DeviceBatch
from .frenet_host import BatchCoordinateConverter, ego_to_frenet_many
from .prediction import DevicePredictionPostprocessor, safety_metrics
from .spline import CubicSpline2D

NORMAL, CAUTION, EMERGENCY = 0, 1, 2
SGAN_DT = 0.4          # the observer samples at the SGAN rate whatever the simulation dt (integrated_simulator.py:325)


def _knob(knobs, key, default):
    v = knobs.get(key, default)
    if v is None or (isinstance(v, float) and math.isnan(v)):
        return default
    return v


class _StateMachines:
    """FailSafeStateMachine for N simulations at once (src/core/state_machine.py:29-278)."""

    def __init__(self, n: int, k: Dict[str, float]):
        self.k = k
        self.state = np.zeros(n, dtype=np.int64)
        self.failures = np.zeros(n, dtype=np.int64)
        combined = _knob(k, "ego_radius", 1.0) + _knob(k, "ped_radius", 0.2)                       # :44-46
        rc, re = k.get("state_machine_recover_clearance_caution"), k.get("state_machine_recover_clearance_emergency")
        nan = lambda v: v is None or (isinstance(v, float) and math.isnan(v))
        self.clearance_caution = _knob(k, "state_machine_safe_distance_caution", 2.0) - combined if nan(rc) else rc
        self.clearance_emergency = _knob(k, "state_machine_safe_distance_emergency", 3.0) - combined if nan(re) else re
        self.trigger_clearance = _knob(k, "state_machine_trigger_clearance_caution", 0.0)
        self.trigger_headway = _knob(k, "state_machine_trigger_time_headway", 0.0)
        self.envelope_decel = _knob(k, "state_machine_envelope_decel", 0.0)
        self.envelope_standoff = _knob(k, "state_machine_envelope_standoff", 0.5)
        self.last_clearance = np.full(n, np.inf)
        self.last_clearance_ahead = np.full(n, np.inf)

    def update(self, idx, found, clearance, clearance_ahead, ego_speed):
        """state_machine.py:116-179 for the simulations `idx`."""
        self.last_clearance[idx] = clearance
        self.last_clearance_ahead[idx] = clearance_ahead
        for j, i in enumerate(idx):
            trig = self.trigger_clearance + self.trigger_headway * max(float(ego_speed[j]), 0.0)
            st, ok, cl = self.state[i], bool(found[j]), float(clearance[j])
            if st == NORMAL:
                if not ok:
                    self.state[i] = CAUTION
                    self.failures[i] += 1
                elif trig > 0.0 and cl < trig:
                    self.state[i] = CAUTION
                    self.failures[i] = 0
                else:
                    self.failures[i] = 0
            elif st == CAUTION:
                if ok and self.failures[i] == 0:
                    if cl > max(self.clearance_caution, trig):
                        self.state[i] = NORMAL
                elif not ok:
                    self.state[i] = EMERGENCY
                    self.failures[i] += 1
                else:
                    self.failures[i] = 0
            else:
                if ok and cl > self.clearance_emergency:
                    self.state[i] = CAUTION

    def planner_config(self, idx):
        """state_machine.py:181-278: (target speed, limits [n,4], max_stop_distance (NaN = none))."""
        k = self.k
        v_target = float(k["ego_target_speed"])
        base = np.array([k["ego_max_speed"], k["ego_max_accel"], k["ego_max_curvature"], _knob(k, "ego_max_lat_accel", 3.0)])
        target = np.empty(len(idx))
        limits = np.tile(base, (len(idx), 1))
        msd = np.full(len(idx), np.nan)
        for j, i in enumerate(idx):
            ca = float(self.last_clearance_ahead[i])
            v_env = None                                                                             # :252-266
            if self.envelope_decel > 0.0 and math.isfinite(ca):
                v_env = math.sqrt(2.0 * self.envelope_decel * max(ca - self.envelope_standoff, 0.0))
            room = max(ca - 0.2, 0.05) if math.isfinite(ca) else None                                # :268-278
            st = self.state[i]
            if st == NORMAL:
                target[j] = v_env if (v_env is not None and v_env < v_target) else v_target
            elif st == CAUTION:
                speed_mult = _knob(k, "state_machine_caution_speed_multiplier", 0.8)
                t = v_target * speed_mult
                if v_env is not None:
                    t = min(t, v_env)
                    if v_env <= 0.0 and room is not None:
                        msd[j] = room
                target[j] = t
                limits[j, 1] = k["ego_max_accel"] * _knob(k, "state_machine_caution_accel_multiplier", 1.5)
                limits[j, 0] = k["ego_max_speed"] * speed_mult
            else:
                target[j] = 0.0
                limits[j, 1] = k["ego_max_accel"] * _knob(k, "state_machine_emergency_accel_multiplier", 3.0)
                limits[j, 3] = _knob(k, "ego_max_lat_accel", 3.0) * _knob(k, "state_machine_emergency_lat_accel_multiplier", 2.0)
                if self.envelope_decel > 0.0 and room is not None:
                    msd[j] = room
        return target, limits, msd


class BatchedClosedLoop:
    """N closed-loop simulations of the reference's planning stack, advanced in lock-step.

    waypoints_x / waypoints_y : the reference path (shared)
    knobs        : SimulationConfig fields by name (dt, obs_len, ego_*, planner and state-machine knobs)
    ped_tracks   : [N, T_frames, P, 2] replayed pedestrian positions, one frame per dt
    ego0         : [N, 5] initial (x, y, yaw, v, a)
    """

    def __init__(self, waypoints_x, waypoints_y, knobs: Dict[str, float], ped_tracks: np.ndarray, ego0: np.ndarray,
                 device: int = 0):
        k = {key: (None if (isinstance(v, float) and math.isnan(v)) else v) for key, v in knobs.items()}
        self.k = k
        self.dt = float(k["dt"])
        self.tracks = np.ascontiguousarray(ped_tracks, dtype=np.float64)
        self.n, self.n_frames, self.P, _ = self.tracks.shape
        self.ego = np.array(ego0, dtype=np.float64).reshape(self.n, 5).copy()
        self.spline = CubicSpline2D(list(waypoints_x), list(waypoints_y))
        self.ego_radius, self.ped_radius = _knob(k, "ego_radius", 1.0), _knob(k, "ped_radius", 0.3)
        self.planner = BatchFrenetPlanner(
            self.spline, max_speed=k["ego_max_speed"], max_accel=k["ego_max_accel"], max_curvature=k["ego_max_curvature"],
            max_lat_accel=_knob(k, "ego_max_lat_accel", 3.0), dt=self.dt, d_road_w=k["d_road_w"],
            max_road_width=k["max_road_width"], robot_radius=self.ego_radius,
            obstacle_radius=_knob(k, "obstacle_radius", self.ped_radius), min_t=_knob(k, "min_t", 4.0),
            max_t=_knob(k, "max_t", 5.0), d_t_s=_knob(k, "d_t_s", 5.0 / 3.6), k_j=k["k_j"], k_t=k["k_t"], k_d=k["k_d"],
            k_s_dot=k["k_s_dot"], k_lat=k["k_lat"], k_lon=k["k_lon"], chance_epsilon=_knob(k, "chance_epsilon", 0.0),
            collision_margin_inflation=_knob(k, "collision_margin_inflation", 1.0), device=device)
        self.device = device
        self.post = DevicePredictionPostprocessor(pred_len=int(_knob(k, "pred_len", 12)), sgan_dt=SGAN_DT, sim_dt=self.dt,
                                                  plan_horizon=_knob(k, "max_t", 5.0), device=device)
        self.fsm = _StateMachines(self.n, k)
        # per-simulation planner state: ego curvature cache and the two nearest-point caches (the planner's
        # converter and the simulator's own goal-check converter are separate objects in the reference)
        self.last_kappa = np.zeros(self.n)
        self.plan_conv = BatchCoordinateConverter(self.spline, self.n)
        self.goal_conv = BatchCoordinateConverter(self.spline, self.n)
        # replayed pedestrians (replay_source.py:31-111): forward-difference velocities, shared clock
        vel = np.zeros_like(self.tracks)
        if self.n_frames >= 2:
            vel[:, :-1] = (self.tracks[:, 1:] - self.tracks[:, :-1]) / self.dt
            vel[:, -1] = vel[:, -2]
        self.vel = vel
        self.frame, self.ped_time = 0, 0.0
        # observer (observer.py:26-86), identical timing for every simulation
        self.obs_len = int(k["obs_len"])
        self.hist, self.hist_t, self.obs_acc, self.obs_last_t = [], [], 0.0, None
        self.time = 0.0
        self.active = np.ones(self.n, dtype=bool)
        self.reason = np.array(["timeout"] * self.n, dtype=object)
        self.n_plan_calls = 0
        self.timers = {"frenet_host": 0.0, "sweep": 0.0, "prediction": 0.0, "metrics": 0.0, "total": 0.0}

    # -- pedestrians + observer -----------------------------------------------------------------
    def _ped_step(self):
        if self.frame < self.n_frames - 1:
            self.frame += 1
        self.ped_time += self.dt
        delta = self.dt if self.obs_last_t is None else max(self.ped_time - self.obs_last_t, 0.0)   # observer.py:58-66
        self.obs_last_t = self.ped_time
        self.obs_acc += delta
        if self.obs_acc + 1e-9 >= SGAN_DT:
            self.hist.append(self.tracks[:, self.frame].copy())
            self.hist_t.append(self.ped_time)
            self.hist, self.hist_t = self.hist[-self.obs_len:], self.hist_t[-self.obs_len:]
            self.obs_acc = max(self.obs_acc - SGAN_DT, 0.0)

    def warmup(self):
        """integrated_simulator.py:406-422."""
        for _ in range(int(self.obs_len * SGAN_DT / self.dt)):
            self._ped_step()

    # -- one planning attempt for the simulations `idx` -------------------------------------------
    def _plan(self, idx, target, limits, msd, dyn_dev):
        import torch
        t0 = time.perf_counter()
        # ok False = conversion failure: plan() returns None (frenet_planner.py:346-374)
        frenet, ok = ego_to_frenet_many(self.plan_conv, idx, self.ego[idx], self.last_kappa[idx])
        self.n_plan_calls += len(idx)
        t1 = time.perf_counter()
        sel = torch.as_tensor(np.asarray(idx), device=dyn_dev.device)
        batch = DeviceBatch(self.planner, frenet, target, dyn_dev.index_select(0, sel), _lib.FOT_DYN_SINGLE,
                            limits=limits, max_stop_distance=msd)
        batch.launch(None)
        best = batch.out["best_idx"].cpu().numpy()
        wlen = batch.out["winner_len"].cpu().numpy()
        win = batch.out["winner"][:, 9:15, :2].cpu().numpy()          # x y yaw c v a, first two samples
        found = ok & (best >= 0)
        self.timers["frenet_host"] += t1 - t0
        self.timers["sweep"] += time.perf_counter() - t1
        return found, wlen, win

    # -- one simulation step for every active simulation -------------------------------------------
    def step(self):
        k, dt = self.k, self.dt
        t_step = time.perf_counter()
        idx = np.nonzero(self.active)[0]
        self._ped_step()
        pos, vel = self.tracks[:, self.frame], self.vel[:, self.frame]
        ts = self.ped_time
        # prediction (integrated_simulator.py:424-527): CV from the observer's last two float32 samples, or the
        # current positions alone while the observer is still filling
        if len(self.hist) >= self.obs_len:
            stale = max(ts - self.hist_t[-1], 0.0)
            dyn = self.post.predict_cv(self.hist[-1], self.hist[-2], stale, pos, obs_float32=True)
        else:
            import torch
            dyn = torch.from_numpy(np.ascontiguousarray(pos[:, None, :, None, :])).to(self.post._dev)
        t_pred = time.perf_counter()
        m = safety_metrics(self.ego, pos, vel, self.ego_radius, self.ped_radius, device=self.device)
        clearance, ahead = m["clearance"].cpu().numpy(), m["clearance_ahead"].cpu().numpy()
        self.timers["prediction"] += t_pred - t_step
        self.timers["metrics"] += time.perf_counter() - t_pred
        last_clearance = ahead.copy()                                  # :566-567 feeds the emergency stop

        # planning cycle with escalation retries (:529-653)
        state_before = self.fsm.state[idx].copy()
        target, limits, msd = self.fsm.planner_config(idx)
        found, wlen, win = self._plan(idx, target, limits, msd, dyn)
        self.fsm.update(idx, found, clearance[idx], ahead[idx], self.ego[idx, 3])
        attempts = np.zeros(len(idx), dtype=np.int64)
        calls = np.ones(len(idx), dtype=np.int64)
        while True:
            retry = np.nonzero(~found & (self.fsm.state[idx] != state_before) & (attempts < 3))[0]
            if len(retry) == 0:
                break
            attempts[retry] += 1
            calls[retry] += 1
            sub = idx[retry]
            t2, l2, m2 = self.fsm.planner_config(sub)
            f2, w2, win2 = self._plan(sub, t2, l2, m2, dyn)
            found[retry], wlen[retry], win[retry] = f2, w2, win2
            failed = retry[~f2]
            state_before[failed] = self.fsm.state[idx[failed]]
            if len(failed):
                self.fsm.update(idx[failed], np.zeros(len(failed), bool), clearance[idx[failed]], ahead[idx[failed]],
                                self.ego[idx[failed], 3])

        # ego update (:655-676) or adaptive emergency stop (:749-802)
        for j, i in enumerate(idx):
            if found[j] and wlen[j] >= 2:
                x, y, yaw, c, v, a = win[j, :, 1]
                self.ego[i] = [x, y, yaw, v, a]
                self.last_kappa[i] = float(c)                          # frenet_planner.py:301-302
            else:
                cap = k.get("ego_emergency_decel") or k["ego_max_accel"] * 2.0
                cl = float(last_clearance[i])
                x, y, yaw, v, a = self.ego[i]
                required = v ** 2 / (2.0 * max(cl - 0.2, 0.05)) if math.isfinite(cl) else cap
                max_dec = float(np.clip(required, k["ego_max_accel"], cap))
                x += v * np.cos(yaw) * dt
                y += v * np.sin(yaw) * dt
                v = max(0.0, v - max_dec * dt)
                self.ego[i] = [x, y, yaw, v, -max_dec if v > 0 else 0.0]
                self.last_kappa[i] = 0.0                               # planner.reset_ego_curvature()
        # termination (:870-886): collision of the NEW ego state with the same pedestrian frame, then the goal
        t_m = time.perf_counter()
        m2 = safety_metrics(self.ego, pos, vel, self.ego_radius, self.ped_radius, device=self.device)
        collided = m2["collision"].cpu().numpy()
        self.timers["metrics"] += time.perf_counter() - t_m
        hit = idx[collided[idx]]
        self.active[hit], self.reason[hit] = False, "collision"
        alive = idx[~collided[idx]]
        if len(alive):
            s_now = self.goal_conv.nearest_s(alive, self.ego[alive, 0], self.ego[alive, 1])
            done = alive[self.spline.s[-1] - s_now < 2.0]
            self.active[done], self.reason[done] = False, "goal"
        self.time += dt
        self.timers["total"] += time.perf_counter() - t_step
        return idx, found, calls

    def run(self, n_steps: Optional[int] = None):
        """IntegratedSimulator.run (:842-892) for every simulation; returns per-simulation histories."""
        if n_steps is None:
            n_steps = int(self.k["total_time"] / self.dt)
        ego_hist = np.full((self.n, n_steps, 5), np.nan)
        fsm_hist = np.full((self.n, n_steps), -1, dtype=np.int64)
        found_hist = np.zeros((self.n, n_steps), dtype=bool)
        calls_hist = np.zeros((self.n, n_steps), dtype=np.int64)
        steps = np.zeros(self.n, dtype=np.int64)
        for s in range(n_steps):
            if not self.active.any():
                break
            idx, found, calls = self.step()
            ego_hist[idx, s] = self.ego[idx]
            fsm_hist[idx, s] = self.fsm.state[idx]
            found_hist[idx, s] = found
            calls_hist[idx, s] = calls
            steps[idx] += 1
        return {"ego": ego_hist, "fsm": fsm_hist, "found": found_hist, "calls": calls_hist, "steps": steps,
                "reason": self.reason.copy()}
