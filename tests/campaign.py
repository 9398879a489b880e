"""Seeded inputs of the differential campaign: shared by the fixture generator
(tests/golden/make_golden_campaign.py, which runs the unmodified reference on them) and by the GPU tests, which
rebuild the same inputs from the seeds and hold the CUDA path to the recorded answers.  Nothing here reads the
reference.

  A  "config4": 5000 BASELINE-config-4 style queries (scenario_01 grid, 50 constant-velocity pedestrians, random ego)
  B  "limits" : 2000 queries whose speed / acceleration / curvature / lateral-acceleration limits and road width are
                drawn right through the range the candidates take, on a straight, an S-curve and an arc
  C  "config5": the 500-step NORMAL -> CAUTION -> EMERGENCY relaxation rollout (3 plan() calls per step, stateful)
  D  "config3": the dense 65 x 32 x 32 (+ brake) grid against 200 pedestrians x 20 samples, epsilon = 0
"""
from __future__ import annotations

import zlib

import numpy as np

from tests import scenarios

N_A, N_B, N_C_STEPS = 5000, 2000, 500
SEED_A, SEED_B, SEED_C = 1_000_000, 2_000_000, 3_000_017
TARGET_SPEED = 6.0
PATHS = {"straight": scenarios.STRAIGHT_60, "s_curve": scenarios.s_curve_waypoints(), "arc": scenarios.arc_waypoints()}
PATH_NAMES = ("straight", "s_curve", "arc")


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# ---- A: config-4 style ------------------------------------------------------------------------------------
def query_a(i: int):
    """-> (ego x,y,yaw,v,a ; dyn [50,51,2]) -- the same draw order as bench.make_queries."""
    rng = np.random.default_rng(SEED_A + i)
    dyn = scenarios.pedestrian_field(rng, 50, 51, scenarios.S1_KNOBS["dt"])
    ego = (rng.uniform(2, 20), rng.uniform(-1, 1), rng.normal(0, 0.05), rng.uniform(0, 8), rng.uniform(-1, 1))
    return tuple(float(v) for v in ego), dyn


# ---- B: limits through the candidates' range ----------------------------------------------------------------
def query_b(i: int):
    """-> dict(path, road, fs[6], target, overrides, dyn [6,51,2]).  Frenet state given directly (the sweep is what
    is under test here; ego -> Frenet has its own tests)."""
    rng = np.random.default_rng(SEED_B + i)
    path = PATH_NAMES[i % 3]
    road = (2.7, 1.1)[(i // 3) % 2]
    over = dict(max_speed=float(rng.uniform(2.0, 9.0)), max_accel=float(rng.uniform(0.3, 3.0)),
                max_curvature=float(rng.uniform(0.01, 0.3)), max_lat_accel=float(rng.uniform(0.05, 2.0)))
    fs = np.array([rng.uniform(3, 25), rng.uniform(0.0, 9.0), rng.uniform(-1.5, 1.5), rng.uniform(-2.9, 2.9),
                   rng.uniform(-1.0, 1.0), rng.uniform(-0.5, 0.5)])
    target = float(rng.uniform(0.5, 9.0))
    if i % 7 == 0:                                   # the EMERGENCY shape: stop target + stop-distance directive
        target = 0.0
    msd = float(rng.uniform(2.0, 12.0)) if i % 7 == 0 else None
    dyn = scenarios.pedestrian_field(rng, 6)
    return dict(path=path, road=road, fs=fs, target=target, overrides=over, dyn=dyn, msd=msd)


def knobs_b(road: float):
    return dict(scenarios.S1_KNOBS, max_road_width=road)


# ---- C: config 5 ---------------------------------------------------------------------------------------------
C_PATH = scenarios.s_curve_waypoints(length=260.0, amp=3.0, n=60)


def plans_c():
    """The three knob sets the reference's fail-safe state machine hands out (state_machine.py:183-248)."""
    k = scenarios.S1_KNOBS
    return [(6.0, None, None),
            (3.6, {"max_accel": k["max_accel"] * 1.5, "max_speed": k["max_speed"] * 0.6}, None),
            (0.0, {"max_accel": k["max_accel"] * 3.0, "max_lat_accel": k["max_lat_accel"] * 2.0}, 5.0)]


def rollout_c(n_steps: int = N_C_STEPS):
    """Yields (step, ego[5], dyn [10,51,2], static or None) of the synthetic rollout: the ego follows the S-curve at a
    speed that wanders between 0 and 8 m/s, with a fresh pedestrian field ahead of it every step and, every fifth
    step, a wall across the lane (the field that makes NORMAL fail and the relaxation levels matter)."""
    rng = np.random.default_rng(SEED_C)
    wx, wy = np.asarray(C_PATH[0]), np.asarray(C_PATH[1])
    x, v = 3.0, 5.0
    for step in range(n_steps):
        y = float(np.interp(x, wx, wy)) + float(rng.normal(0, 0.15))
        slope = float(np.interp(x + 0.5, wx, wy) - np.interp(x - 0.5, wx, wy))
        yaw = float(np.arctan2(slope, 1.0)) + float(rng.normal(0, 0.02))
        ego = (x, y, yaw, v, float(rng.uniform(-1.0, 1.0)))
        dyn = scenarios.pedestrian_field(rng, 10, x_range=(x + 3.0, x + 30.0), y_range=(y - 6.0, y + 6.0))
        static = None
        if step % 5 == 4:
            xw = x + float(rng.uniform(8.0, 25.0))
            yw = float(np.interp(xw, wx, wy))
            static = np.stack([np.full(9, xw), np.linspace(yw - 1.0, yw + 1.0, 9)], axis=1)
        yield step, ego, dyn, static
        x += 0.1 * max(v, 0.5)
        v = float(np.clip(v + rng.normal(0, 0.35), 0.0, 8.0))


# ---- D: config 3 ---------------------------------------------------------------------------------------------
def config3(variant: int = 0):
    """BASELINE config 3: 65 d x 32 T x 32 v (+ 3 brake) = 66 563 candidates against 200 pedestrians x 20 samples.
    variant 0: pedestrians everywhere (nearly every candidate collides); variant 1: the lane itself mostly kept free,
    so that thousands of candidates survive all 20 samples."""
    rng = np.random.default_rng(33 + variant)
    knobs = dict(scenarios.S1_KNOBS, d_road_w=0.1, max_road_width=3.2, min_t=1.9, max_t=5.0, d_t_s=0.2)
    wp = (np.linspace(0.0, 80.0, 9).tolist(), [0.0] * 9)
    base = scenarios.pedestrian_field(rng, 200 if variant == 0 else 600, x_range=(5.0, 65.0), vel_clip=2.5)
    if variant == 1:
        keep = np.abs(base[:, :, 1]).min(axis=1) > 3.6
        base = base[keep][:200]
        assert base.shape[0] == 200
    dist = scenarios.sample_distribution(rng, base, 20)
    fs = np.array([5.0, 5.0, 0.0, 0.0, 0.0, 0.0])
    return dict(knobs=knobs, waypoints=wp, fs=fs, target=6.2, dist=dist)
