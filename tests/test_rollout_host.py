"""Host logic of the batched roll-out driver (no GPU): the array form of the fail-safe state machine against a
scalar restatement of reference `src/core/state_machine.py:116-278`, one simulation at a time, over random
sequences of planning results and clearances; the static-obstacle expansion."""
import math

import numpy as np

from integrated_path_planning_b200.rollout import CAUTION, EMERGENCY, NORMAL, _StateMachines, expand_static_obstacles

KNOBS = dict(ego_radius=1.0, ped_radius=0.3, ego_target_speed=6.0, ego_max_speed=10.0, ego_max_accel=2.0,
             ego_max_curvature=0.2, ego_max_lat_accel=3.0, state_machine_safe_distance_caution=2.5,
             state_machine_safe_distance_emergency=3.5, state_machine_trigger_clearance_caution=0.8,
             state_machine_trigger_time_headway=0.3, state_machine_recover_clearance_caution=None,
             state_machine_recover_clearance_emergency=None, state_machine_caution_speed_multiplier=0.6,
             state_machine_caution_accel_multiplier=1.5, state_machine_emergency_accel_multiplier=3.0,
             state_machine_emergency_lat_accel_multiplier=2.0, state_machine_envelope_decel=1.2,
             state_machine_envelope_standoff=0.5)


class ScalarMachine:
    """state_machine.py:29-278, one vehicle."""

    def __init__(self, k):
        self.k, self.state, self.failures = k, NORMAL, 0
        combined = k["ego_radius"] + k["ped_radius"]
        rc, re = k["state_machine_recover_clearance_caution"], k["state_machine_recover_clearance_emergency"]
        self.cc = rc if rc is not None else k["state_machine_safe_distance_caution"] - combined
        self.ce = re if re is not None else k["state_machine_safe_distance_emergency"] - combined
        self.ca = float("inf")

    def update(self, ok, clearance, ahead, v):
        k = self.k
        self.ca = ahead
        trig = k["state_machine_trigger_clearance_caution"] + k["state_machine_trigger_time_headway"] * max(v, 0.0)
        if self.state == NORMAL:
            if not ok:
                self.state, self.failures = CAUTION, self.failures + 1
            elif trig > 0.0 and clearance < trig:
                self.state, self.failures = CAUTION, 0
            else:
                self.failures = 0
        elif self.state == CAUTION:
            if ok and self.failures == 0:
                if clearance > max(self.cc, trig):
                    self.state = NORMAL
            elif not ok:
                self.state, self.failures = EMERGENCY, self.failures + 1
            else:
                self.failures = 0
        elif ok and clearance > self.ce:
            self.state = CAUTION

    def config(self):
        k = self.k
        dec = k["state_machine_envelope_decel"]
        v_env = None
        if dec > 0.0 and math.isfinite(self.ca):
            v_env = math.sqrt(2.0 * dec * max(self.ca - k["state_machine_envelope_standoff"], 0.0))
        room = max(self.ca - 0.2, 0.05) if math.isfinite(self.ca) else None
        lim = [k["ego_max_speed"], k["ego_max_accel"], k["ego_max_curvature"], k["ego_max_lat_accel"]]
        msd = None
        if self.state == NORMAL:
            target = v_env if (v_env is not None and v_env < k["ego_target_speed"]) else k["ego_target_speed"]
        elif self.state == CAUTION:
            target = k["ego_target_speed"] * k["state_machine_caution_speed_multiplier"]
            if v_env is not None:
                target = min(target, v_env)
                if v_env <= 0.0:
                    msd = room
            lim[1] = k["ego_max_accel"] * k["state_machine_caution_accel_multiplier"]
            lim[0] = k["ego_max_speed"] * k["state_machine_caution_speed_multiplier"]
        else:
            target = 0.0
            lim[1] = k["ego_max_accel"] * k["state_machine_emergency_accel_multiplier"]
            lim[3] = k["ego_max_lat_accel"] * k["state_machine_emergency_lat_accel_multiplier"]
            msd = room if dec > 0.0 else None
        return target, lim, msd


def _run(knobs, seed, n=40, steps=300):
    rng = np.random.default_rng(seed)
    batch, scalars = _StateMachines(n, knobs), [ScalarMachine(knobs) for _ in range(n)]
    for _ in range(steps):
        idx = np.nonzero(rng.random(n) < 0.85)[0]
        target, limits, msd = batch.planner_config(idx)
        for j, i in enumerate(idx):
            t, lim, m = scalars[i].config()
            assert target[j] == t and limits[j].tolist() == lim, (i, target[j], t)
            assert (np.isnan(msd[j]) and m is None) or msd[j] == m
        ok = rng.random(len(idx)) < 0.7
        clearance = np.where(rng.random(len(idx)) < 0.1, np.inf, rng.uniform(-0.2, 4.0, len(idx)))
        ahead = np.where(rng.random(len(idx)) < 0.2, np.inf, clearance + rng.uniform(0, 1.0, len(idx)))
        speed = rng.uniform(-0.5, 8.0, len(idx))
        batch.update(idx, ok, clearance, ahead, speed)
        for j, i in enumerate(idx):
            scalars[i].update(bool(ok[j]), float(clearance[j]), float(ahead[j]), float(speed[j]))
        assert batch.state.tolist() == [s.state for s in scalars]
        assert batch.failures.tolist() == [s.failures for s in scalars]
    assert set(batch.state.tolist()) == {NORMAL, CAUTION, EMERGENCY} or steps < 50


def test_state_machines_match_the_scalar_restatement():
    _run(KNOBS, 1)


def test_state_machines_without_envelope_and_trigger():
    _run(dict(KNOBS, state_machine_envelope_decel=0.0, state_machine_trigger_clearance_caution=0.0,
              state_machine_trigger_time_headway=0.0, state_machine_recover_clearance_caution=1.0,
              state_machine_recover_clearance_emergency=2.0), 2)


def test_static_rectangles_expand_to_boundary_points():
    walls = [[-5.0, 55.0, -6.0, -5.0], [-5.0, 55.0, 5.0, 6.0]]
    pts = expand_static_obstacles(walls)
    assert pts.shape == (488, 2)                              # 2 x (121 x 2 + 3 x 2 - 4 shared corners)
    assert np.array_equal(pts, np.unique(pts, axis=0))        # sorted rows, no duplicates
    assert pts[:, 0].min() == -5.0 and pts[:, 0].max() == 55.0 and set(np.unique(pts[:, 1])) == {-6.0, -5.5, -5.0, 5.0, 5.5, 6.0}
    assert expand_static_obstacles(None).shape == (0, 2) and expand_static_obstacles(pts) is not None
    assert np.array_equal(expand_static_obstacles(pts), pts)
