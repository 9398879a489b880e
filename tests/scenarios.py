"""Seeded synthetic planning queries shared by the parity tests, the golden generator and bench.py.

Shapes follow SURVEY.md section 8(d).  Nothing here reads /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

# scenario_01 knobs (reference scenarios/scenario_01_cv.yaml:7-65 + SimulationConfig defaults)
S1_KNOBS = dict(max_speed=10.0, max_accel=2.0, max_curvature=0.2, dt=0.1, d_road_w=0.3, max_road_width=2.7,
                robot_radius=1.0, obstacle_radius=0.2, min_t=4.0, max_t=5.0, d_t_s=5.0 / 3.6, n_s_sample=1,
                max_lat_accel=3.0, k_j=1.0, k_t=1.0, k_d=1.0, k_s_dot=1.0, k_lat=1.0, k_lon=1.0)
STRAIGHT_60 = ([0.0, 10.0, 20.0, 30.0, 40.0, 50.0, 60.0], [0.0] * 7)


def arc_waypoints(radius=15.0, span=1.2 * np.pi, n=40):
    th = np.linspace(0.0, span, n)
    return (radius * np.sin(th)).tolist(), (radius * (1.0 - np.cos(th))).tolist()


def s_curve_waypoints(length=80.0, amp=4.0, n=30):
    x = np.linspace(0.0, length, n)
    return x.tolist(), (amp * np.sin(2 * np.pi * x / length)).tolist()


def pedestrian_field(rng, n_peds, n_steps=51, dt=0.1, x_range=(5.0, 45.0), y_range=(-10.0, 10.0),
                     vel_clip=None):
    """Constant-velocity tracks dyn[p, k] = p0 + vel * k * dt  (SURVEY.md 8d config 2)."""
    p0 = np.stack([rng.uniform(*x_range, n_peds), rng.uniform(*y_range, n_peds)], axis=1)
    vel = rng.normal(size=(n_peds, 2))
    if vel_clip is not None:
        vel = np.clip(vel, -vel_clip, vel_clip)
    t = np.arange(n_steps) * dt
    return p0[:, None, :] + vel[:, None, :] * t[None, :, None]


def sample_distribution(rng, dyn, n_samples, sigma=0.05):
    """SGAN-style samples: the CV track plus a per-sample random walk, zero at k=0 (config 3)."""
    walk = np.cumsum(rng.normal(0.0, sigma, size=(n_samples,) + dyn.shape), axis=2)
    walk[:, :, 0] = 0.0
    return dyn[None] + walk


@dataclass
class Query:
    """Everything one plan() call needs."""
    name: str
    knobs: Dict[str, float]
    waypoints: tuple
    ego: tuple                       # x, y, yaw, v, a
    target_speed: float
    last_kappa: float = 0.0
    static: Optional[np.ndarray] = None
    dyn: Optional[np.ndarray] = None         # [P, T, 2]
    dist: Optional[np.ndarray] = None        # [S, P, T, 2]
    overrides: Optional[Dict[str, float]] = None
    max_stop_distance: Optional[float] = None
    footprint: Optional[tuple] = None        # (vehicle_length, vehicle_width, n_circles)
    extra: dict = field(default_factory=dict)


def wall(x=24.0, half=8.0, n=33):
    ys = np.linspace(-half, half, n)
    return np.stack([np.full_like(ys, x), ys], axis=1)


def standard_queries():
    """A spread of cases covering every branch of the priority chain and both collision modes."""
    out = []
    rng = np.random.default_rng
    k = S1_KNOBS
    out.append(Query("s1_cfg2_seed0", k, STRAIGHT_60, (5.0, 0.0, 0.0, 5.0, 0.0), 6.0,
                     dyn=pedestrian_field(rng(0), 50)))
    out.append(Query("s1_sparse", k, STRAIGHT_60, (5.0, 0.2, 0.03, 5.0, 0.1), 6.0,
                     dyn=pedestrian_field(rng(1), 8)))
    out.append(Query("s1_caution", k, STRAIGHT_60, (5.0, 0.3, 0.05, 5.0, 0.2), 3.6,
                     dyn=pedestrian_field(rng(2), 6), overrides={"max_accel": 3.0, "max_speed": 6.0}))
    out.append(Query("s1_emergency_stop", k, STRAIGHT_60, (5.0, 0.3, 0.05, 5.0, 0.2), 0.0,
                     dyn=pedestrian_field(rng(3), 4), overrides={"max_accel": 6.0, "max_lat_accel": 6.0},
                     max_stop_distance=6.0))
    out.append(Query("s1_near_end_truncation", k, STRAIGHT_60, (40.0, -0.3, -0.05, 5.0, 0.2), 6.0,
                     dyn=pedestrian_field(rng(4), 3)))
    out.append(Query("s1_no_obstacles", k, STRAIGHT_60, (12.0, 0.1, 0.0, 4.0, 0.0), 6.0))
    out.append(Query("s1_standstill", k, STRAIGHT_60, (5.0, 0.0, 0.0, 0.0, 0.0), 6.0,
                     dyn=pedestrian_field(rng(5), 5)))
    out.append(Query("s1_wall", k, STRAIGHT_60, (5.0, 0.0, 0.0, 5.0, 0.0), 6.0, static=wall(),
                     dyn=pedestrian_field(rng(6), 5)))
    d6 = pedestrian_field(rng(7), 12)
    out.append(Query("s1_distribution_eps0", k, STRAIGHT_60, (5.0, 0.3, 0.05, 4.0, 0.2), 6.0,
                     dyn=d6, dist=sample_distribution(rng(8), d6, 6)))
    k2 = dict(k, chance_epsilon=0.2, collision_margin_inflation=1.3)
    d7 = pedestrian_field(rng(9), 12)
    out.append(Query("s1_distribution_eps02", k2, STRAIGHT_60, (5.0, 0.3, 0.05, 4.0, 0.2), 6.0,
                     dyn=d7, dist=sample_distribution(rng(10), d7, 10)))
    out.append(Query("s1_inflated_single", k2, STRAIGHT_60, (5.0, 0.3, 0.05, 4.0, 0.2), 6.0,
                     dyn=pedestrian_field(rng(11), 10)))
    out.append(Query("arc_curved", k, arc_waypoints(), (1.0, 0.1, 0.1, 4.0, 0.0), 6.0,
                     dyn=pedestrian_field(rng(12), 10, x_range=(0.0, 20.0), y_range=(-5.0, 25.0))))
    out.append(Query("arc_footprint", k, arc_waypoints(), (1.0, 0.1, 0.1, 4.0, 0.0), 6.0,
                     dyn=pedestrian_field(rng(13), 10, x_range=(0.0, 20.0), y_range=(-5.0, 25.0)),
                     footprint=(4.5, 2.0, 3)))
    k3 = dict(k, max_speed=13.9, max_accel=8.0, max_curvature=10.0, d_road_w=0.5, max_road_width=7.0, d_t_s=1.39)
    th = np.linspace(0.0, 1.5 * np.pi, 60)
    out.append(Query("tight_arc_singularity", k3, ((5 * np.sin(th)).tolist(), (5 * (1 - np.cos(th))).tolist()),
                     (0.5, 0.0, 0.1, 3.0, 0.0), 3.0))
    kd = dict(k, d_road_w=0.5, max_road_width=7.0)    # SimulationConfig defaults, config 2b
    out.append(Query("cfg2b_defaults", kd, s_curve_waypoints(), (3.0, 0.2, 0.3, 6.0, 0.3), 8.33,
                     dyn=pedestrian_field(rng(14), 20, x_range=(5.0, 60.0), y_range=(-8.0, 8.0)), last_kappa=0.01))
    out.append(Query("short_obstacle_horizon", k, STRAIGHT_60, (5.0, 0.0, 0.0, 5.0, 0.0), 6.0,
                     dyn=pedestrian_field(rng(15), 6, n_steps=13)))
    out.append(Query("nan_pedestrian", k, STRAIGHT_60, (5.0, 0.0, 0.0, 5.0, 0.0), 6.0,
                     dyn=_with_nan(pedestrian_field(rng(16), 6))))
    return out


def _with_nan(dyn):
    dyn = dyn.copy()
    dyn[0, 40:, :] = np.nan      # NaN tail: the reference's AABB prefilter drops this pedestrian entirely
    return dyn
