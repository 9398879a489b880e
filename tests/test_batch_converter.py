"""BatchCoordinateConverter == N scalar CoordinateConverters, bit for bit (host logic, no GPU).

The scalar converter is pinned to the reference by the golden closed-loop calls (tests/test_oracle_golden.py,
tests/test_gpu_rollout.py); the batched one has to return exactly the same numbers and keep exactly the same
caches, through cached windows, stale-window global searches, path ends and first calls.
"""
import os

import numpy as np

from integrated_path_planning_b200.frenet_host import (BatchCoordinateConverter, CoordinateConverter, ego_to_frenet,
                                                       ego_to_frenet_many)
from integrated_path_planning_b200.spline import CubicSpline2D
from integrated_path_planning_b200.types import EgoVehicleState

HERE = os.path.dirname(os.path.abspath(__file__))


def _path():
    z = np.load(os.path.join(HERE, "golden", "rollout_s01.npz"))
    return CubicSpline2D(list(z["v0/wx"]), list(z["v0/wy"]))


def _curved_path():
    s = np.linspace(0, 60, 13)
    return CubicSpline2D(list(s), list(6.0 * np.sin(s / 9.0)))


def _walks(path, n, steps, rng):
    """Ego tracks along the path with lateral noise, a jump now and then (stale window -> global search)."""
    length = path.s[-1]
    s = rng.uniform(0, 0.3 * length, n)
    out = np.zeros((steps, n, 5))
    for t in range(steps):
        s = np.minimum(s + rng.uniform(0.0, 0.9, n), length)
        jump = rng.random(n) < 0.03
        s = np.where(jump, rng.uniform(0, length, n), s)
        px, py = path.calc_position(s)
        yaw = np.array([path.calc_yaw(v) for v in s])
        d = rng.normal(0, 0.8, n)
        out[t, :, 0] = px - d * np.sin(yaw)
        out[t, :, 1] = py + d * np.cos(yaw)
        out[t, :, 2] = yaw + rng.normal(0, 0.1, n)
        out[t, :, 3] = rng.uniform(0, 9, n)
        out[t, :, 4] = rng.normal(0, 1, n)
    return out


def _check(path, n=24, steps=60, seed=0):
    rng = np.random.default_rng(seed)
    ego = _walks(path, n, steps, rng)
    kappa = rng.normal(0, 0.05, (steps, n))
    scalar = [CoordinateConverter(path) for _ in range(n)]
    batch = BatchCoordinateConverter(path, n)
    for t in range(steps):
        idx = np.nonzero(rng.random(n) < 0.8)[0]              # a changing subset, as simulations finish or retry
        if len(idx) == 0:
            continue
        got, ok = ego_to_frenet_many(batch, idx, ego[t, idx], kappa[t, idx])
        for j, i in enumerate(idx):
            want = ego_to_frenet(scalar[i], EgoVehicleState(*ego[t, i]), float(kappa[t, i]))
            assert ok[j] == (want is not None)
            if want is not None:
                assert np.array_equal(got[j], want), (t, i, got[j] - want)
            assert batch.prev_s[i] == scalar[i]._prev_s


def test_straight_path_bit_identical():
    _check(_path(), seed=1)


def test_curved_path_bit_identical():
    _check(_curved_path(), seed=2)


def test_nearest_s_only_matches_goal_check():
    path = _curved_path()
    rng = np.random.default_rng(3)
    ego = _walks(path, 16, 40, rng)
    scalar = [CoordinateConverter(path) for _ in range(16)]
    batch = BatchCoordinateConverter(path, 16)
    for t in range(40):
        s = batch.nearest_s(np.arange(16), ego[t, :, 0], ego[t, :, 1])
        for i in range(16):
            assert s[i] == scalar[i].find_nearest_point_on_path(ego[t, i, 0], ego[t, i, 1])[0]


def test_near_tie_rows_use_the_reference_arithmetic():
    """A point exactly on the perpendicular bisector of two window samples: the two distances are equal or an
    ulp apart, the row is recomputed with math.hypot and the first minimum wins as in the reference loop."""
    path = _path()
    scalar, batch = CoordinateConverter(path), BatchCoordinateConverter(path, 1)
    for cv in (scalar,):
        cv._prev_s = 10.0
    batch.prev_s[0] = 10.0
    grid = np.linspace(0.0, 20.0, 100)
    mid = 0.5 * (grid[40] + grid[41])
    px, py = path.calc_position(mid)
    x, y = float(np.asarray(px).reshape(-1)[0]), float(np.asarray(py).reshape(-1)[0]) + 0.7
    want = scalar.find_nearest_point_on_path(x, y)[0]
    got = batch.nearest_s([0], [x], [y])[0]
    assert got == want


def test_lean_three_point_evaluator_is_the_reference_point_by_point():
    """The refinement steps evaluate their three points through one [2, 3] array expression; a converter that calls
    the path's calc_position once per point, as the reference does (coordinate_converter.py:253-280), must see
    exactly the same nearest points, Frenet states and caches -- straight path, curved path, segment borders."""
    for path, seed in ((_path(), 5), (_curved_path(), 6)):
        rng = np.random.default_rng(seed)
        ego = _walks(path, 6, 80, rng)
        knots = np.asarray(path.s if not hasattr(path, "sx") else path.sx.x)
        for i in range(6):
            lean, literal = CoordinateConverter(path), CoordinateConverter(path)
            literal._batched = False
            for t in range(80):
                e = ego[t, i].copy()
                if t % 9 == 0:                                 # park the ego next to a knot: the three points straddle segments
                    kx, ky = path.calc_position(float(knots[1 + (t // 9) % (len(knots) - 2)]))
                    e[0], e[1] = float(np.asarray(kx).reshape(-1)[0]), float(np.asarray(ky).reshape(-1)[0]) + 0.4
                a = ego_to_frenet(lean, EgoVehicleState(*e), 0.01)
                b = ego_to_frenet(literal, EgoVehicleState(*e), 0.01)
                assert (a is None) == (b is None)
                if a is not None:
                    assert np.array_equal(a, b), (i, t, a - b)
                assert lean._prev_s == literal._prev_s
