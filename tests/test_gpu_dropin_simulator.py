"""The drop-in, executed: the CUDA planner under the reference's OWN IntegratedSimulator.

`sim_mod.FrenetPlanner = integrated_path_planning_b200.FrenetPlanner` before `IntegratedSimulator(config)`
(integrated_simulator.py:23, :342-366); the simulator then calls it at :576-584 / :622-630, reads
`last_check_stats` at :732 and resets the curvature cache at :800-802 -- all unmodified reference code, imported
from oracle/_ref (the byte-for-byte copy that travels to the GPU box; oracle/make_ref.py).

Every plan() call of the closed loop is also given, on identical inputs and identical planner state, to the
reference's NumPy planner (stock `plan()`), and must select the same candidate index.
"""
import numpy as np
import pytest

from oracle import ref_loader
from tests import runners

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")]

RTOL = 1e-9     # north_star: trajectory points and costs within 1e-9 relative (fp64)
ATOL = 1e-9


def _lockstep(sim, shadow):
    """Wrap sim.planner.plan (the CUDA planner): each call is mirrored to `shadow` (reference planner) first."""
    ours = sim.planner
    rec = ref_loader.IndexRecorder(shadow)
    cuda_plan = ours.plan
    log = []

    def plan(ego_state, static_obstacles, dynamic_obstacles=None, target_speed=None, constraint_overrides=None,
             dynamic_obstacles_distribution=None, max_stop_distance=None):
        # identical planner state on both sides: ego curvature cache and the nearest-point window cache
        state_equal = (shadow._last_kappa == ours._last_kappa and
                       getattr(shadow.converter, "_prev_s", None) == getattr(ours.converter, "_prev_s", None))
        shadow._last_kappa = ours._last_kappa
        if hasattr(ours.converter, "_prev_s"):
            shadow.converter._prev_s = ours.converter._prev_s
        elif hasattr(shadow.converter, "_prev_s"):
            del shadow.converter._prev_s
        kw = dict(target_speed=target_speed, constraint_overrides=constraint_overrides,
                  dynamic_obstacles_distribution=dynamic_obstacles_distribution, max_stop_distance=max_stop_distance)
        static = None if static_obstacles is None else np.array(static_obstacles, copy=True)
        dyn = None if dynamic_obstacles is None else np.array(dynamic_obstacles, copy=True)
        want = rec.plan(ego_state, static, dyn, **kw)
        got = cuda_plan(ego_state, static_obstacles, dynamic_obstacles, **kw)
        n = len(log)
        idx = int(ours.last_result.best_idx[0]) if ours.last_result is not None else -1
        assert idx == rec.last_index, (n, idx, rec.last_index)
        assert ours.last_check_stats == shadow.last_check_stats, (n, ours.last_check_stats, shadow.last_check_stats)
        assert (got is None) == (want is None), n
        if got is not None:
            np.testing.assert_allclose(float(got.cost), float(want.cost), rtol=RTOL, atol=0, err_msg=f"call {n} cost")
            for name in runners.SERIES:
                a, b = np.asarray(getattr(got, name), dtype=float), np.asarray(getattr(want, name), dtype=float)
                assert a.shape == b.shape, (n, name, a.shape, b.shape)
                np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, err_msg=f"call {n} {name}")
            # the types the simulator consumes (data_structures.py:149-220)
            assert isinstance(got.x, list) and isinstance(got.s, np.ndarray) and len(got) == len(want)
        assert ours._last_kappa == pytest.approx(shadow._last_kappa, rel=RTOL, abs=1e-15)
        log.append(dict(index=idx, n_cand=rec.last_n_candidates, state_equal=state_equal,
                        overrides=constraint_overrides, target=target_speed))
        return got

    ours.plan = plan
    return log


@pytest.mark.parametrize("scenario,footprint,max_steps", [("scenario_01_cv", False, 400), ("scenario_03_cv", False, 120),
                                                          ("scenario_02_cv", True, 60)])
def test_cuda_planner_drives_the_reference_simulator(scenario, footprint, max_steps):
    import integrated_path_planning_b200 as b200
    sim, cfg, sim_mod = ref_loader.scenario_simulator(scenario, footprint, planner_cls=b200.FrenetPlanner)
    assert type(sim.planner) is b200.FrenetPlanner                       # the reference constructed OUR class
    shadow_sim, _, _ = ref_loader.scenario_simulator(scenario, footprint)
    shadow = shadow_sim.planner
    assert type(shadow).__module__ == "src.planning.frenet_planner"
    log = _lockstep(sim, shadow)
    sim.warmup()
    states, egos, reason = [], [], "timeout"
    for _ in range(max_steps):
        res = sim.step()
        states.append(sim.state_machine.current_state.name)
        egos.append([sim.ego_state.x, sim.ego_state.y, sim.ego_state.yaw, sim.ego_state.v, sim.ego_state.a])
        if res.metrics.get("collision", False):
            reason = "collision"
            break
        e = sim.ego_state
        if sim.reference_path.s[-1] - sim.coord_converter.find_nearest_point_on_path(e.x, e.y)[0] < 2.0:
            reason = "goal"
            break
    n_calls = len(log)
    assert n_calls >= len(states)
    retries = sum(1 for c in log if c["overrides"])
    print(f"\n{scenario} footprint={footprint}: {len(states)} steps ({reason}), {n_calls} plan() calls, {retries} with "
          f"constraint overrides, winners identical on all; planner state bit-equal before {sum(c['state_equal'] for c in log)} "
          f"calls; states {dict((s, states.count(s)) for s in set(states))}")
    if scenario == "scenario_01_cv":
        # the run the reference itself makes (tests/golden/rollout_s01.npz variant 0: 274 steps to the goal through all
        # three fail-safe states) -- same length, same state sequence, same ego track
        z = runners.load_golden("rollout_s01.npz")
        order = {"NORMAL": 0, "CAUTION": 1, "EMERGENCY": 2}
        assert reason == str(z["v0/reason"]) and len(states) == len(z["v0/fsm"])
        assert np.array_equal([order[s] for s in states], z["v0/fsm"])
        np.testing.assert_allclose(np.array(egos), z["v0/ego"], rtol=1e-7, atol=1e-7)
        assert retries > 0 and n_calls == int(z["v0/calls"].sum())
