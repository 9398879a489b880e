"""GPU: batched entry points, unusual shapes, and full-size properties of the sweep."""
import numpy as np
import pytest

from oracle import frenet_oracle as O
from tests import runners, scenarios

pytestmark = pytest.mark.gpu


def _planner(knobs=None, wp=scenarios.STRAIGHT_60, **kw):
    from integrated_path_planning_b200 import BatchFrenetPlanner, CubicSpline2D
    return BatchFrenetPlanner(CubicSpline2D(*wp), **dict(knobs or scenarios.S1_KNOBS, **kw))


def _random_frenet(rng, n):
    return np.stack([rng.uniform(2, 20, n), rng.uniform(0, 8, n), rng.uniform(-1, 1, n),
                     rng.uniform(-1, 1, n), rng.normal(0, 0.3, n), rng.normal(0, 0.05, n)], axis=1)


def _oracle_batch(knobs, wp, frenet, target, dyn=None, dist=None, static=None, overrides=None, msd=None):
    pl = O.OraclePlanner(O.Spline2D(*wp), O.Knobs(**knobs))
    out = []
    for i in range(len(frenet)):
        t = target if np.isscalar(target) else target[i]
        m = None if msd is None else (None if np.isnan(msd[i]) else float(msd[i]))
        out.append(pl.plan_frenet(tuple(frenet[i]), static, None if dyn is None else dyn[i], t, overrides,
                                  None if dist is None else dist[i], m))
    return out


def _check_batch(res, refs):
    for i, ref in enumerate(refs):
        n_c = len(ref.categories)
        assert int(res.n_cand[i]) == n_c
        assert np.array_equal(res.cand_cat[i, :n_c].astype(np.int8), ref.categories), i
        np.testing.assert_allclose(res.cand_cost[i, :n_c], ref.costs, rtol=runners.RTOL)
        assert int(res.best_idx[i]) == ref.best_index, (i, int(res.best_idx[i]), ref.best_index)
        want = [ref.stats.get(k, 0) for k in runners.GOLDEN_STAT_KEYS]
        assert res.stats[i].tolist() == want, (i, res.stats[i].tolist(), want)
        if ref.best_index >= 0:
            series = res.series(i)
            for name in runners.SERIES:
                np.testing.assert_allclose(series[name], ref.arrays[name], rtol=runners.RTOL, atol=runners.ATOL)


def test_batch_matches_oracle_with_per_query_speed_grids_and_stop_distance():
    rng = np.random.default_rng(11)
    n = 24
    frenet = _random_frenet(rng, n)
    frenet[3, 1] = 0.05                                   # below BRAKE_MIN_SPEED: no brake ladder for this query
    target = np.array([6.0, 3.6, 0.0, 6.0] * 6)            # n_v = 6, 4, 1, 6 ...
    msd = np.where(target == 0.0, 6.0, np.nan)
    dyn = np.stack([scenarios.pedestrian_field(np.random.default_rng(100 + i), 9) for i in range(n)])
    pl = _planner()
    res = pl.plan_batch(frenet, target, dynamic_obstacles=dyn, max_stop_distance=msd, want_candidates=True)
    _check_batch(res, _oracle_batch(scenarios.S1_KNOBS, scenarios.STRAIGHT_60, frenet, target, dyn=dyn, msd=msd))


def test_batch_equals_single_calls():
    rng = np.random.default_rng(5)
    n = 16
    frenet = _random_frenet(rng, n)
    dyn = np.stack([scenarios.pedestrian_field(np.random.default_rng(7 + i), 12) for i in range(n)])
    pl = _planner()
    batch = pl.plan_batch(frenet, 6.0, dynamic_obstacles=dyn)
    best = batch.best_idx.copy()
    cost = batch.best_cost.copy()
    for i in range(n):
        path = pl.plan_from_frenet(frenet[i], None, dyn[i], 6.0)
        assert int(pl.last_result.best_idx[0]) == int(best[i])
        if path is not None:
            assert float(path.cost) == float(cost[i])


def test_static_per_query_and_shared():
    rng = np.random.default_rng(2)
    n = 6
    frenet = _random_frenet(rng, n)
    wall = scenarios.wall(x=30.0)
    pl = _planner()
    shared = pl.plan_batch(frenet, 6.0, static_obstacles=wall, want_candidates=True)
    _check_batch(shared, _oracle_batch(scenarios.S1_KNOBS, scenarios.STRAIGHT_60, frenet, 6.0, static=wall))
    per_q = np.stack([scenarios.wall(x=26.0 + 2 * i) for i in range(n)])
    res = pl.plan_batch(frenet, 6.0, static_obstacles=per_q, want_candidates=True)
    refs = [_oracle_batch(scenarios.S1_KNOBS, scenarios.STRAIGHT_60, frenet[i:i + 1], 6.0, static=per_q[i])[0]
            for i in range(n)]
    _check_batch(res, refs)


def test_many_static_points_use_chunked_tiles():
    """More static points than one ring stage holds (> 512): the chunked-tile path."""
    ys = np.linspace(-9.0, 9.0, 700)
    wall = np.stack([np.full_like(ys, 33.0) + 0.3 * np.sin(ys), ys], axis=1)
    gap = np.abs(wall[:, 1] - 1.5) > 1.3                   # leave a gap so that some candidates survive
    wall = wall[gap]
    assert len(wall) > 512
    q = scenarios.Query("static_chunks", scenarios.S1_KNOBS, scenarios.STRAIGHT_60, (5.0, 0.0, 0.0, 5.0, 0.0), 6.0,
                        static=wall, dyn=scenarios.pedestrian_field(np.random.default_rng(8), 4))
    ref = runners.run_oracle(q)
    pl, path = runners.run_cuda(q)
    runners.assert_matches_oracle(q, ref, pl, path)
    assert ref.stats["collision_error"] > 0 and ref.stats["ok"] > 0


def test_large_distribution_uses_chunked_planes():
    """S*P > 512 entries per time plane (config 3 style, reduced): chunked planes, epsilon = 0."""
    rng = np.random.default_rng(21)
    knobs = dict(scenarios.S1_KNOBS, d_road_w=0.9, d_t_s=3.0)          # 7 d x 11 T x 3 v: keep the oracle fast
    base = scenarios.pedestrian_field(rng, 60, x_range=(5.0, 55.0), vel_clip=2.5)
    keep = np.abs(base[:, :, 1]).min(axis=1) > 2.2                    # keep the lane itself mostly free
    base = base[keep][:40]
    dist = scenarios.sample_distribution(rng, base, 16, sigma=0.02)  # 16 x 40 = 640 entries per plane
    q = scenarios.Query("dist_chunks", knobs, scenarios.STRAIGHT_60, (5.0, 0.2, 0.0, 5.0, 0.0), 6.0,
                        dyn=base, dist=dist)
    ref = runners.run_oracle(q)
    pl, path = runners.run_cuda(q)
    runners.assert_matches_oracle(q, ref, pl, path)
    assert ref.stats["collision_error"] > 0


def test_long_time_grid_pairwise_sum_split():
    """dt = 0.02 -> 201..251 samples per candidate: np.sum's >128-element halving path in the cost."""
    knobs = dict(scenarios.S1_KNOBS, dt=0.02, d_road_w=0.9, d_t_s=3.0, min_t=4.0, max_t=5.0)
    knobs["max_t"] = 4.06                                              # 4 horizons: keep the oracle fast
    q = scenarios.Query("long_grid", knobs, scenarios.STRAIGHT_60, (5.0, 0.1, 0.02, 5.0, 0.1), 6.0,
                        dyn=scenarios.pedestrian_field(np.random.default_rng(4), 5, n_steps=260, dt=0.02))
    ref = runners.run_oracle(q)
    pl, path = runners.run_cuda(q)
    runners.assert_matches_oracle(q, ref, pl, path)
    rep = runners.bit_exact_report(ref, pl, path)
    assert rep["cost_bit_exact_frac"] == 1.0


def test_state_machine_relaxation_rollout():
    """Config 5: every step plans three times with the knobs the reference's fail-safe state machine
    hands out (state_machine.py:183-248): NORMAL, CAUTION (max_accel x1.5, max_speed x0.6, target x0.6),
    EMERGENCY (max_accel x3, max_lat_accel x2, target 0, stop-distance directive); curvature never relaxed."""
    k = scenarios.S1_KNOBS
    rng = np.random.default_rng(17)
    plans = [(6.0, None, None),
             (3.6, {"max_accel": k["max_accel"] * 1.5, "max_speed": k["max_speed"] * 0.6}, None),
             (0.0, {"max_accel": k["max_accel"] * 3.0, "max_lat_accel": k["max_lat_accel"] * 2.0}, 5.0)]
    from integrated_path_planning_b200 import CubicSpline2D, FrenetPlanner
    cu = FrenetPlanner(CubicSpline2D(*scenarios.STRAIGHT_60), **k)
    orc = O.OraclePlanner(O.Spline2D(*scenarios.STRAIGHT_60), O.Knobs(**k))
    ego = np.array([3.0, 0.1, 0.0, 5.0, 0.0])
    counts = []
    for step in range(12):
        dyn = scenarios.pedestrian_field(rng, 10, x_range=(ego[0] + 3, ego[0] + 30), y_range=(-6, 6))
        for target, ovr, msd in plans:
            ref = orc.plan(tuple(ego), np.empty((0, 2)), dyn, target, ovr, None, msd)
            path = cu.plan(runners._Ego(*ego), np.empty((0, 2)), dyn, target, ovr, None, msd, _want_candidates=True)
            q = scenarios.Query(f"step{step}_t{target}", k, scenarios.STRAIGHT_60, tuple(ego), target)
            runners.assert_matches_oracle(q, ref, cu, path)
            counts.append(len(ref.categories))
            assert cu._last_kappa == orc.last_kappa
        ego = ego + np.array([0.45, rng.normal(0, 0.02), rng.normal(0, 0.005), rng.normal(0, 0.1), 0.0])
    assert {1261, 843, 216} <= set(counts)


def test_full_size_batch_properties():
    """BASELINE config 4 at full size (4096 queries): properties that do not need the oracle --
    duplicate queries give identical answers wherever they sit in the batch, category counts add up
    to the candidates generated, winners are categorised OK and carry the minimum cost, and a
    sample of queries agrees with the oracle."""
    import bench
    spline, frenet, dyn = bench.make_queries(0, 256)
    reps = 16
    frenet_b = np.tile(frenet, (reps, 1))
    dyn_b = np.tile(dyn, (reps, 1, 1, 1, 1))
    perm = np.random.default_rng(0).permutation(len(frenet_b))
    pl = _planner()
    res = pl.plan_batch(frenet_b[perm], 6.0, dynamic_obstacles=dyn_b[perm][:, 0], want_candidates=True)
    assert len(res.best_idx) == 4096
    inv = np.argsort(perm)
    best = res.best_idx[inv].reshape(reps, -1)
    cost = res.best_cost[inv].reshape(reps, -1)
    stats = res.stats[inv].reshape(reps, 256, -1)
    assert np.all(best == best[0]) and np.all(cost == cost[0]) and np.all(stats == stats[0])
    dropped = (res.cand_cat == 8).sum(axis=1)           # unused slots beyond n_cand hold 9
    assert np.array_equal(res.stats.sum(axis=1) + dropped, res.n_cand)
    ok = res.best_idx >= 0
    rows = np.nonzero(ok)[0]
    assert np.all(res.cand_cat[rows, res.best_idx[rows]] == 0)
    masked = np.where(res.cand_cat == 0, res.cand_cost, np.inf)
    assert np.array_equal(masked.min(axis=1)[rows], res.best_cost[rows])
    assert np.array_equal(masked.argmin(axis=1)[rows], res.best_idx[rows])
    assert np.all(np.isinf(res.best_cost[~ok]))
    sample = [0, 17, 101, 255]
    refs = _oracle_batch(scenarios.S1_KNOBS, scenarios.STRAIGHT_60, frenet[sample], 6.0, dyn=dyn[sample][:, 0])
    first = inv.reshape(reps, -1)[0]
    for j, i in enumerate(sample):
        assert int(res.best_idx[first[i]]) == refs[j].best_index


def test_device_batch_to_host_compact_read_back():
    """DeviceBatch.launch_to_host(winner_samples=k): the first k samples of every winner series, equal to the head of the
    full read-back; indices, costs and histograms unchanged."""
    from integrated_path_planning_b200 import _lib
    from integrated_path_planning_b200.batch import DeviceBatch
    rng = np.random.default_rng(41)
    n = 300
    frenet = _random_frenet(rng, n)
    dyn = np.stack([scenarios.pedestrian_field(rng, 20) for _ in range(n)])[:, None]
    pl = _planner()
    db = DeviceBatch(pl, frenet, 6.0, dyn, _lib.FOT_DYN_SINGLE)
    full = {k: np.zeros(tuple(v.shape), dtype=v.cpu().numpy().dtype) for k, v in db.out.items()}
    db.launch_to_host(full)
    k_head = 2
    head = {k: (np.zeros((n, v.shape[1], k_head)) if k == "winner" else np.zeros_like(v)) for k, v in full.items()}
    db.launch_to_host(head, winner_samples=k_head)
    for key in ("best_idx", "stats", "winner_len"):
        assert np.array_equal(full[key], head[key]), key
    assert np.array_equal(full["best_cost"].view(np.uint64), head["best_cost"].view(np.uint64))
    has = full["best_idx"] >= 0
    assert has.sum() > 20
    assert np.array_equal(head["winner"][has].view(np.uint64), full["winner"][has][:, :, :k_head].view(np.uint64))
