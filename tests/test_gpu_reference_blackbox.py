"""GPU: the reference's own black-box `plan()` tests, restated against the drop-in planner.

SURVEY.md section 8(c) lists the reference tests that exercise this path through `plan()` only
(tests/test_frenet_conventions.py, test_planner_guards.py, test_smooth_braking.py,
test_frenet_planner.py::test_plan_end_to_end).  Each test below states the same property for
`integrated_path_planning_b200.FrenetPlanner` with the same planner set-up, and additionally checks
the call against the oracle (same winner, same category histogram), which the reference's tests
cannot do.  Citations give the reference test each one follows.
"""
import numpy as np
import pytest

from oracle import frenet_oracle as O
from tests import runners

pytestmark = pytest.mark.gpu

NO_OBS = np.empty((0, 2))


def _wrap(a):
    return np.angle(np.exp(1j * np.asarray(a)))


class _Case:
    """A planner pair (CUDA drop-in + oracle) on the same spline and knobs."""

    def __init__(self, xs, ys, **knobs):
        from integrated_path_planning_b200 import CubicSpline2D, FrenetPlanner
        self.knobs = knobs
        self.cuda = FrenetPlanner(CubicSpline2D(xs, ys), **knobs)
        self.oracle = O.OraclePlanner(O.Spline2D(xs, ys), O.Knobs(**knobs))

    def plan(self, ego, static=NO_OBS, dyn=None, target_speed=5.0, overrides=None, msd=None):
        """plan() on the drop-in, cross-checked against the oracle; returns the FrenetPath or None."""
        self.oracle.last_kappa = self.cuda._last_kappa
        self.oracle.search.prev_s = getattr(self.cuda.converter, "_prev_s", None)
        ref = self.oracle.plan(tuple(ego), static, dyn, target_speed, overrides, None, msd)
        path = self.cuda.plan(runners._Ego(*ego), static, dyn, target_speed, overrides, None, msd)
        assert (path is None) == (ref.best_index < 0)
        assert int(self.cuda.last_result.best_idx[0]) == ref.best_index
        want = {k: v for k, v in ref.stats.items()}
        assert self.cuda.last_check_stats == want, (self.cuda.last_check_stats, want)
        if path is not None:
            assert float(path.cost) == float(ref.cost)
        return path


def straight_conventions(length=120.0, **kw):
    """test_frenet_conventions.py:24-42 make_straight_planner."""
    n = int(length / 10) + 1
    knobs = dict(max_speed=10.0, max_accel=2.0, max_curvature=1.0, dt=0.1, d_road_w=1.0, max_road_width=7.0,
                 robot_radius=1.0, obstacle_radius=0.3)
    knobs.update(kw)
    return _Case([10.0 * i for i in range(n)], [0.0] * n, **knobs)


def straight_braking(**kw):
    """test_smooth_braking.py:21-29 make_straight_planner."""
    knobs = dict(max_speed=10.0, max_accel=2.0, max_curvature=0.2, dt=0.1, d_road_w=0.5, max_road_width=3.0,
                 robot_radius=1.0, obstacle_radius=0.2, min_t=4.0, max_t=5.0, d_t_s=1.39, n_s_sample=1)
    knobs.update(kw)
    return _Case(np.linspace(0, 80, 30).tolist(), [0.0] * 30, **knobs)


# ---- test_frenet_conventions.py ------------------------------------------------------------------
def test_yaw_matches_polyline_tangent():
    """:60-71 -- the stored yaw agrees with the tangent of the converted polyline."""
    path = straight_conventions().plan((20.0, 0.0, np.deg2rad(15.0), 5.0, 0.0))
    assert path is not None
    x, y, yaw = np.asarray(path.x), np.asarray(path.y), np.asarray(path.yaw)
    tangent = np.arctan2(np.diff(y), np.diff(x))
    assert np.max(np.abs(_wrap(yaw[:-1] - tangent))) < np.deg2rad(5.0)


def test_initial_speed_continuity():
    """:73-79 -- the converted speed at index 0 equals the ego speed."""
    path = straight_conventions().plan((20.0, 0.0, np.deg2rad(15.0), 5.0, 0.0))
    assert path is not None and np.isclose(path.v[0], 5.0, atol=1e-6)


def test_plan_from_standstill_is_finite():
    """:81-93 -- s_dot ~ 0 must not blow up the spatial-derivative conversion."""
    path = straight_conventions().plan((20.0, 0.0, np.deg2rad(10.0), 0.0, 0.0))
    assert path is not None
    for arr in (path.x, path.y, path.yaw, path.v, path.a, path.c):
        assert np.all(np.isfinite(arr))


def test_grid_contains_zero_and_is_symmetric():
    """:97-105 -- on an empty straight road the cheapest candidate ends exactly on d = 0."""
    path = straight_conventions(d_road_w=0.3, max_road_width=7.0).plan((20.0, 0.0, 0.0, 5.0, 0.0))
    assert path is not None and np.isclose(path.d[-1], 0.0, atol=1e-9)


def test_ti_range_includes_max_t():
    """:125-133 -- the horizon grid handed to the device ends on max_t."""
    case = straight_conventions(min_t=4.0, max_t=5.0)
    assert case.plan((20.0, 0.0, 0.0, 5.0, 0.0)) is not None
    T = case.cuda.engine.T
    assert np.isclose(T[0], 4.0) and np.isclose(T[-1], 5.0) and len(T) == 11


def test_collision_checked_at_horizon_endpoint_only_at_the_same_time():
    """:135-152 through plan(): a pedestrian standing where the ego will be at t = 5.0 blocks the
    candidates that are there at that time; the same place at another time does not."""
    case = straight_conventions(min_t=5.0, max_t=5.0, d_road_w=7.0)        # one lateral offset: d = 0 (+-7 leave the road)
    free = case.plan((20.0, 0.0, 0.0, 5.0, 0.0))
    assert free is not None
    x_end = free.x[-1]
    dyn = np.full((1, 51, 2), 1000.0)
    dyn[0, 50] = [x_end, 0.0]
    blocked = case.plan((20.0, 0.0, 0.0, 5.0, 0.0), dyn=dyn)
    assert case.cuda.last_check_stats["collision_error"] > 0
    assert blocked is None or abs(blocked.x[-1] - x_end) > 1.0
    dyn_other = np.full((1, 51, 2), 1000.0)
    dyn_other[0, 10] = [x_end, 0.0]                                        # same place, wrong time
    again = case.plan((20.0, 0.0, 0.0, 5.0, 0.0), dyn=dyn_other)
    assert again is not None and np.isclose(again.x[-1], x_end)


def test_truncated_path_arrays_stay_in_lockstep():
    """:156-170 -- paths leaving the spline domain are truncated across all 15 arrays."""
    case = straight_conventions(length=60.0)
    path = case.plan((45.0, 0.0, 0.0, 5.0, 0.0))
    assert path is not None
    n = len(path.x)
    assert n < 41                                                          # shorter than the shortest horizon
    for arr in (path.y, path.yaw, path.c, path.v, path.a, path.t, path.s, path.s_d, path.s_dd, path.s_ddd,
                path.d, path.d_d, path.d_dd, path.d_ddd):
        assert len(arr) == n


def test_ego_curvature_cache_updates_on_success_and_survives_failure():
    """:186-207."""
    case = straight_conventions()
    pl = case.cuda
    assert pl._last_kappa == 0.0
    ego = (20.0, 0.0, 0.0, 5.0, 0.0)
    path = case.plan(ego)
    assert path is not None and pl._last_kappa == float(path.c[1])
    kept = pl._last_kappa
    wall_y = np.linspace(-8.0, 8.0, 33)
    wall = np.stack([np.full_like(wall_y, 24.0), wall_y], axis=1)
    assert case.plan(ego, static=wall) is None
    assert pl._last_kappa == kept
    pl.reset_ego_curvature()
    assert pl._last_kappa == 0.0


# ---- test_planner_guards.py ----------------------------------------------------------------------
def test_out_of_domain_paths_are_truncated_not_dropped():
    """:116-133 -- candidates that overrun the spline end keep their valid prefix."""
    case = _Case(np.linspace(0, 60, 25).tolist(), [0.0] * 25, max_speed=10.0, max_accel=8.0, max_curvature=10.0,
                 dt=0.1, d_road_w=0.5, max_road_width=7.0, robot_radius=1.0, min_t=4.0, max_t=5.0, d_t_s=1.39,
                 n_s_sample=1)
    path = case.plan((45.0, 0.0, 0.0, 6.0, 0.0), dyn=np.empty((0, 0, 2)), target_speed=6.0)
    assert path is not None and len(path.x) >= 2


def test_straight_reference_unaffected_by_singularity_guard():
    """:135-147."""
    case = _Case(np.linspace(0, 50, 20).tolist(), [0.0] * 20, max_speed=13.9, max_accel=8.0, max_curvature=10.0,
                 dt=0.1, d_road_w=0.5, max_road_width=7.0, robot_radius=1.0, min_t=4.0, max_t=5.0, d_t_s=1.39,
                 n_s_sample=1)
    path = case.plan((5.0, 0.0, 0.0, 5.0, 0.0), dyn=np.empty((0, 0, 2)))
    assert path is not None and len(path.x) > 1


def test_candidates_beyond_curvature_center_are_dropped_silently():
    """:86-114 through plan(): on a radius-5 arc the candidates whose offset crosses the curvature centre
    are dropped and counted in no category (they are missing from the histogram's total)."""
    theta = np.linspace(0.0, 1.5 * np.pi, 60)
    case = _Case((5 * np.sin(theta)).tolist(), (5 * (1 - np.cos(theta))).tolist(), max_speed=13.9, max_accel=8.0,
                 max_curvature=10.0, dt=0.1, d_road_w=0.5, max_road_width=7.0, robot_radius=1.0, min_t=4.0, max_t=5.0,
                 d_t_s=1.39, n_s_sample=1)
    case.plan((0.5, 0.0, 0.1, 3.0, 0.0), target_speed=3.0)
    res = case.cuda.last_result
    assert sum(case.cuda.last_check_stats.values()) < int(res.n_cand[0])


# ---- test_smooth_braking.py ----------------------------------------------------------------------
def test_plan_yields_short_stop_when_wall_inside_min_t_distance():
    """:76-90 -- every grid candidate hits the wall 6 m ahead; a brake-ladder stop is returned."""
    case = straight_braking(max_accel=8.0)
    ys = np.arange(-3.5, 3.6, 0.25)
    wall = np.stack([np.full_like(ys, 16.0), ys], axis=1)
    path = case.plan((10.0, 0.0, 0.0, 5.0, 0.0), static=wall, dyn=np.empty((0, 0, 2)))
    assert path is not None
    assert path.v[-1] == pytest.approx(0.0, abs=0.05)
    assert max(path.x) < 16.0 - 1.0
    assert int(case.cuda.last_result.best_idx[0]) >= int(case.cuda.last_result.n_cand[0]) - case.cuda.engine.n_B


def test_stop_distance_filter_keeps_only_short_stops():
    """:94-114."""
    case = straight_braking(max_accel=8.0)
    ego = (10.0, 0.0, 0.0, 3.0, 0.0)
    lazy = case.plan(ego, dyn=np.empty((0, 0, 2)), target_speed=0.0)
    committed = case.plan(ego, dyn=np.empty((0, 0, 2)), target_speed=0.0, msd=2.5)
    assert lazy is not None and committed is not None
    assert lazy.s[-1] - lazy.s[0] > 4.0
    assert committed.s[-1] - committed.s[0] <= 2.5 + 1e-6
    assert abs(committed.v[-1]) < 0.15
    assert case.cuda.last_check_stats["stop_distance_error"] > 0


def test_stop_distance_infeasible_room_fails_plan():
    """:116-124."""
    case = straight_braking(max_accel=8.0)
    assert case.plan((10.0, 0.0, 0.0, 5.0, 0.0), dyn=np.empty((0, 0, 2)), target_speed=0.0, msd=0.05) is None


def test_stop_distance_hold_in_place_when_already_stopped():
    """:126-134."""
    case = straight_braking(max_accel=8.0)
    path = case.plan((10.0, 0.0, 0.0, 0.05, 0.0), dyn=np.empty((0, 0, 2)), target_speed=0.0, msd=0.3)
    assert path is not None and (path.s[-1] - path.s[0]) <= 0.3 + 1e-6


# ---- test_frenet_planner.py ----------------------------------------------------------------------
def test_plan_end_to_end_rejects_a_mock_spline():
    """:472-487 uses a MagicMock spline; the device evaluates the spline itself, so an object without
    coefficients is rejected loudly instead of being sampled through Python callbacks."""
    from unittest.mock import MagicMock
    from integrated_path_planning_b200 import FrenetPlanner
    pl = FrenetPlanner(MagicMock(), max_speed=10.0, max_accel=2.0, max_curvature=1.0, dt=0.1, d_road_w=1.0,
                       max_road_width=7.0, robot_radius=1.0)
    with pytest.raises(TypeError):
        pl.engine
