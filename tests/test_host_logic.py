"""CPU: host-side logic of the package (grids, spline, ego->Frenet, API surface, C-ABI exports).
No compute entry point is called here -- there is no GPU in this tier."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import frenet_oracle as O
from tests import scenarios

import integrated_path_planning_b200 as ipp
from integrated_path_planning_b200 import _lib, engine
from integrated_path_planning_b200.frenet_host import CoordinateConverter, ego_to_frenet
from integrated_path_planning_b200.planner import classify_dynamic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/fot.h <-> libfot.so <-> ctypes table: same set of functions."""
    header = open(os.path.join(ROOT, "include", "fot.h")).read()
    declared = set(re.findall(r"^\s*(?:int|float|const char\*)\s+(fot_\w+)\s*\(", header, flags=re.M))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fot_abi_version() == _lib.FOT_ABI_VERSION


def test_struct_layouts_match_header_sizes():
    # fot_config_t: 12 doubles + 8 circle offsets + 6 int32 ; fot_batch_t / fot_result_t as declared
    assert ctypes.sizeof(_lib.FotConfig) == (12 + 8) * 8 + 6 * 4
    assert ctypes.sizeof(_lib.FotTables) == 18 * 8
    assert ctypes.sizeof(_lib.FotBatch) == 2 * 4 + 6 * 8 + 8 + 2 * 4 + 8 + 4 * 4
    assert ctypes.sizeof(_lib.FotResult) == 7 * 8 + 2 * 4


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pl = ipp.FrenetPlanner(ipp.CubicSpline2D(*scenarios.STRAIGHT_60), **scenarios.S1_KNOBS)
    with pytest.raises(_lib.FotError):
        pl.plan(ipp.EgoVehicleState(5.0, 0.0, 0.0, 5.0, 0.0), np.empty((0, 2)), target_speed=6.0)
    assert pl.last_check_stats is None


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "integrated_path_planning_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in text.replace("# oracle", ""), fn


@pytest.mark.parametrize("target", [6.0, 3.6, 0.0, 8.33, 5.0 / 3.6, 1e-10, 0.5])
def test_speed_grid_matches_reference_formula(target):
    k = O.Knobs(**scenarios.S1_KNOBS)
    want = O.speed_grid(k, target)
    got = engine.speed_grid(target, k.d_t_s)
    assert np.array_equal(got, want)
    grid, n_v = engine.speed_grid_batch(np.array([target, 6.0, 0.0]), k.d_t_s)
    assert n_v[0] == len(want) and np.array_equal(grid[0, :n_v[0]], want)
    assert n_v[2] == 1 and grid[2, 0] == 0.0


def test_grids_and_time_tables_match_oracle():
    k = O.Knobs(**scenarios.S1_KNOBS)
    assert np.array_equal(engine.horizon_grid(k.min_t, k.max_t, k.dt), O.horizon_grid(k))
    assert np.array_equal(engine.lateral_grid(k.d_road_w, k.max_road_width), O.lateral_grid(k))
    for T in O.horizon_grid(k):
        n, i4, i5 = engine.time_table(T, k.dt)
        tt = O.time_table(T, k.dt)
        assert n == len(tt.t) - 1 and np.array_equal(i4, tt.inv4) and np.array_equal(i5, tt.inv5)


@pytest.mark.parametrize("wp", [scenarios.STRAIGHT_60, scenarios.arc_waypoints(), scenarios.s_curve_waypoints()])
def test_spline_coefficients_and_eval_match_oracle(wp):
    ours, ref = ipp.CubicSpline2D(*wp), O.Spline2D(*wp)
    for a, b in ((ours.sx, ref.sx), (ours.sy, ref.sy)):
        for name in "abcd":
            assert np.array_equal(getattr(a, name), getattr(b, name))
    s = np.linspace(-1.0, float(ours.s[-1]) + 1.0, 57)
    for f, g in ((ours.calc_yaw, ref.yaw), (ours.calc_curvature, ref.curvature), (ours.calc_curvature_rate, ref.curvature_rate)):
        np.testing.assert_array_equal(f(s), g(s))
    np.testing.assert_array_equal(ours.calc_position(s)[0], ref.position(s)[0])
    tabs = ipp.spline.spline_tables(ours)
    assert tabs["knots"].dtype == np.float64 and tabs["xb"].shape == (len(wp[0]) - 1,)


def test_spline_tables_rejects_objects_without_coefficients():
    from unittest.mock import MagicMock
    with pytest.raises(TypeError):
        ipp.spline.spline_tables(MagicMock())


def test_ego_to_frenet_matches_oracle_including_cache():
    wp = scenarios.s_curve_waypoints()
    conv, search = CoordinateConverter(ipp.CubicSpline2D(*wp)), O.NearestPointSearch(O.Spline2D(*wp))
    rng = np.random.default_rng(3)
    x = 2.0
    for step in range(12):                      # a drive along the path exercises the _prev_s window
        x += rng.uniform(0.2, 6.0)
        ego = (x, rng.uniform(-2, 2), rng.normal(0, 0.2), rng.uniform(0, 8), rng.uniform(-1, 1))
        kappa = rng.normal(0, 0.01)
        got = ego_to_frenet(conv, ipp.EgoVehicleState(*ego), kappa)
        want = O.ego_to_frenet(search, ego, kappa)
        np.testing.assert_array_equal(got, np.array(want))
        assert conv._prev_s == search.prev_s


def test_frenet_path_contract():
    p = ipp.FrenetPath()
    assert len(p) == 0 and p.cost == float("inf")
    p = ipp.FrenetPath(t=[0.0, 0.1], x=[1.0, 2.0], y=[0.0, 0.0], yaw=[0.0, 0.1], v=[1.0, 1.0], a=[0.0, 0.0], c=[0.0, 0.2])
    assert len(p) == 2
    st = p.get_state_at_index(1)
    assert (st.x, st.yaw, st.timestamp) == (2.0, 0.1, 0.1)
    with pytest.raises(IndexError):
        p.get_state_at_index(2)


def test_planner_constructor_mirrors_reference_defaults():
    pl = ipp.FrenetPlanner(ipp.CubicSpline2D(*scenarios.STRAIGHT_60))
    assert pl.dt == 0.2 and pl.k_j == 0.1 and pl.max_lat_accel == 3.0 and pl.robot_radius == 2.0
    assert pl.obstacle_radius == 0.3 and pl.chance_epsilon == 0.0 and pl.footprint is None
    assert pl._last_kappa == 0.0 and pl.last_check_stats is None
    pl._last_kappa = 0.3
    pl.reset_ego_curvature()
    assert pl._last_kappa == 0.0
    lim = pl.resolve_limits({"max_accel": 6.0, "max_lat_accel": 6.0})
    assert lim.tolist() == [pl.max_speed, 6.0, pl.max_curvature, 6.0]


def test_classify_dynamic_follows_reference_precedence():
    dyn = np.zeros((3, 5, 2))
    dist = np.zeros((4, 3, 5, 2))
    assert classify_dynamic(None, None)[0] == _lib.FOT_DYN_NONE
    assert classify_dynamic(np.empty((0, 0, 2)), None)[0] == _lib.FOT_DYN_NONE
    mode, arr = classify_dynamic(dyn, None)
    assert mode == _lib.FOT_DYN_SINGLE and arr.shape == (1, 1, 3, 5, 2)
    mode, arr = classify_dynamic(dyn, dist)            # distribution wins when non-empty (fp.py:1043)
    assert mode == _lib.FOT_DYN_DISTRIBUTION and arr.shape == (1, 4, 3, 5, 2)
    assert classify_dynamic(dyn, np.empty((0, 3, 5, 2)))[0] == _lib.FOT_DYN_SINGLE
    assert classify_dynamic(dyn[:, :1], None)[1].shape == (1, 1, 3, 1, 2)   # (P,1,2) accepted


def test_shard_bounds_partition():
    for n_q, world in ((4096, 8), (10, 4), (3, 8), (1, 1)):
        spans = [ipp.shard_bounds(n_q, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n_q
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
