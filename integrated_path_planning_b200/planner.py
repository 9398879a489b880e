"""Drop-in `FrenetPlanner`: the reference planner's Python API over the B200 sweep.

Same constructor, `plan()` signature, attributes and side effects as reference
`src/planning/frenet_planner.py:125-332`, so it can be handed to the
reference's `IntegratedSimulator` in place of the NumPy planner
(`integrated_simulator.py:342-366, 576-584, 622-630, 800-802`).  What runs where:

  host (this file, O(1) per call)   ego -> Frenet state, constraint overrides, grids, FrenetPath
  GPU  (csrc/, via include/fot.h)   candidate generation, Frenet->global, validity chain,
                                    collision tests, cost, arg-min, winner regeneration

There is no CPU path for the sweep: without libfot.so or a CUDA device `plan()` raises.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

from . import _lib
from .engine import SweepEngine, SweepResult
from .frenet_host import CoordinateConverter, ego_to_frenet
from .types import FrenetPath, FrenetState, SERIES

# module defaults of the reference (frenet_planner.py:25-43, :91)
MAX_SPEED = 50.0 / 3.6
MAX_ACCEL = 2.0
MAX_CURVATURE = 1.0
MAX_ROAD_WIDTH = 7.0
D_ROAD_W = 0.5
DT = 0.2
MAX_T = 5.0
MIN_T = 4.0
TARGET_SPEED = 30.0 / 3.6
D_T_S = 5.0 / 3.6
N_S_SAMPLE = 1
K_J = 0.1
K_T = 0.1
K_D = 1.0
K_S_DOT = 1.0
K_LAT = 1.0
K_LON = 1.0
ROBOT_RADIUS = 2.0
MAX_LAT_ACCEL = 3.0

_REF_STAT_ORDER = ("max_speed_error", "max_accel_error", "max_curvature_error", "max_lat_accel_error",
                   "road_bound_error", "collision_error", "ok")


def _default_device() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def classify_dynamic(dynamic_obstacles, distribution):
    """Which dynamic-obstacle tensor the reference would use (frenet_planner.py:1043-1047,
    :1205-1208) -> (mode, array [1,S,P,T,2] or None)."""
    if distribution is not None and np.size(distribution) > 0:
        dist = np.asarray(distribution, dtype=np.float64)
        if dist.ndim == 4 and dist.shape[-1] == 2 and dist.shape[1] > 0 and dist.shape[2] > 0:
            return _lib.FOT_DYN_DISTRIBUTION, dist[None]
        return _lib.FOT_DYN_NONE, None       # every per-sample test returns False (:1205-1208)
    if dynamic_obstacles is not None and np.size(dynamic_obstacles) > 0:
        dyn = np.asarray(dynamic_obstacles, dtype=np.float64)
        if dyn.ndim == 3 and dyn.shape[-1] == 2:
            return _lib.FOT_DYN_SINGLE, dyn[None, None]
    return _lib.FOT_DYN_NONE, None


class FrenetPlanner:
    """Frenet optimal-trajectory planner (Werling et al. 2010) -- B200 candidate sweep."""

    SINGULARITY_EPS = 0.05
    EPS_S_DOT = 1e-3

    def __init__(self, reference_path, max_speed: float = MAX_SPEED, max_accel: float = MAX_ACCEL,
                 max_curvature: float = MAX_CURVATURE, dt: float = DT, d_road_w: float = D_ROAD_W,
                 max_road_width: float = MAX_ROAD_WIDTH, robot_radius: float = ROBOT_RADIUS,
                 obstacle_radius: float = 0.3, min_t: float = MIN_T, max_t: float = MAX_T,
                 d_t_s: float = D_T_S, n_s_sample: int = N_S_SAMPLE, **kwargs):
        self.csp = reference_path
        self.max_speed = max_speed
        self.max_accel = max_accel
        self.max_curvature = max_curvature
        self.max_lat_accel = float(kwargs.get("max_lat_accel", MAX_LAT_ACCEL))
        self.dt = dt
        self.d_road_w = d_road_w
        self.max_road_width = max_road_width
        self.converter = CoordinateConverter(reference_path)
        self.robot_radius = robot_radius
        self.obstacle_radius = obstacle_radius
        self.min_t = min_t
        self.max_t = max_t
        self.d_t_s = d_t_s
        self.n_s_sample = n_s_sample            # accepted, unused (frenet_planner.py:401-403)
        self.k_j = kwargs.get("k_j", K_J)
        self.k_t = kwargs.get("k_t", K_T)
        self.k_d = kwargs.get("k_d", K_D)
        self.k_s_dot = kwargs.get("k_s_dot", K_S_DOT)
        self.k_lat = kwargs.get("k_lat", K_LAT)
        self.k_lon = kwargs.get("k_lon", K_LON)
        self.chance_epsilon = float(kwargs.get("chance_epsilon", 0.0))
        self.collision_margin_inflation = float(kwargs.get("collision_margin_inflation", 1.0))
        self.footprint = kwargs.get("footprint", None)
        self.device = int(kwargs.get("device", _default_device()))
        self._last_kappa = 0.0
        self.last_check_stats: Optional[Dict[str, int]] = None
        self.last_result: Optional[SweepResult] = None   # diagnostics: raw C-ABI output of the last call
        self._engine: Optional[SweepEngine] = None

    # -- engine ----------------------------------------------------------------------------
    @property
    def engine(self) -> SweepEngine:
        """The device handle, created on first use (so the host-side API can be exercised on a
        machine without a GPU; the sweep itself cannot)."""
        if self._engine is None:
            self._engine = SweepEngine(
                self.csp, max_speed=self.max_speed, dt=self.dt, d_road_w=self.d_road_w,
                max_road_width=self.max_road_width, robot_radius=self.robot_radius,
                obstacle_radius=self.obstacle_radius, min_t=self.min_t, max_t=self.max_t, d_t_s=self.d_t_s,
                k_j=self.k_j, k_t=self.k_t, k_d=self.k_d, k_s_dot=self.k_s_dot, k_lat=self.k_lat,
                k_lon=self.k_lon, chance_epsilon=self.chance_epsilon,
                collision_margin_inflation=self.collision_margin_inflation, footprint=self.footprint,
                device=self.device)
        return self._engine

    def resolve_limits(self, constraint_overrides) -> np.ndarray:
        """[max_speed, max_accel, max_curvature, max_lat_accel] after overrides (frenet_planner.py:921-930)."""
        lim = [self.max_speed, self.max_accel, self.max_curvature, self.max_lat_accel]
        if constraint_overrides:
            for i, key in enumerate(("max_speed", "max_accel", "max_curvature", "max_lat_accel")):
                lim[i] = constraint_overrides.get(key, lim[i])
        return np.array(lim, dtype=np.float64)

    # -- API -------------------------------------------------------------------------------
    def plan(self, ego_state, static_obstacles, dynamic_obstacles=None, target_speed: float = TARGET_SPEED,
             constraint_overrides: Optional[Dict[str, float]] = None, dynamic_obstacles_distribution=None,
             max_stop_distance: Optional[float] = None, _want_candidates: bool = False) -> Optional[FrenetPath]:
        """Best trajectory from the current ego state, or None (frenet_planner.py:227-304)."""
        self.last_check_stats = None
        self.last_result = None
        fs = self._cartesian_to_frenet_state(ego_state)
        if fs is None:
            return None
        return self.plan_from_frenet(fs.to_array(), static_obstacles, dynamic_obstacles, target_speed,
                                     constraint_overrides, dynamic_obstacles_distribution, max_stop_distance,
                                     _want_candidates)

    def plan_from_frenet(self, frenet_state, static_obstacles, dynamic_obstacles=None,
                         target_speed: float = TARGET_SPEED, constraint_overrides=None,
                         dynamic_obstacles_distribution=None, max_stop_distance=None,
                         _want_candidates: bool = False) -> Optional[FrenetPath]:
        """frenet_planner.py:271-304 for a given Frenet state [s, s_d, s_dd, d, d_d, d_dd]."""
        mode, dyn = classify_dynamic(dynamic_obstacles, dynamic_obstacles_distribution)
        static = None
        if static_obstacles is not None and len(static_obstacles) > 0:
            static = np.asarray(static_obstacles, dtype=np.float64).reshape(-1, 2)
        stop = np.nan if max_stop_distance is None else float(max_stop_distance)
        res = self.engine.run_host(
            np.asarray(frenet_state, dtype=np.float64).reshape(1, 6), float(target_speed),
            self.resolve_limits(constraint_overrides), stop, static, dyn, mode,
            want_candidates=_want_candidates)
        self.last_result = res
        stats = {key: int(res.stats[0, _lib.STAT_KEYS.index(key)]) for key in _REF_STAT_ORDER}
        if max_stop_distance is not None:
            stats["stop_distance_error"] = int(res.stats[0, _lib.STAT_KEYS.index("stop_distance_error")])
        self.last_check_stats = stats
        series = res.series(0)
        if series is None:
            return None
        path = FrenetPath(cost=np.float64(res.best_cost[0]))
        for name in SERIES[:9]:
            setattr(path, name, series[name])
        for name in SERIES[9:]:
            setattr(path, name, series[name].tolist())
        if len(path.c) > 1:
            self._last_kappa = float(path.c[1])
        return path

    def reset_ego_curvature(self):
        """frenet_planner.py:326-332."""
        self._last_kappa = 0.0

    def _cartesian_to_frenet_state(self, ego_state) -> Optional[FrenetState]:
        """frenet_planner.py:334-374."""
        arr = ego_to_frenet(self.converter, ego_state, self._last_kappa)
        return None if arr is None else FrenetState(*arr)
