"""CPU: the oracle against the golden fixtures recorded from the unmodified reference."""
import numpy as np
import pytest

from oracle import frenet_oracle as O
from tests import runners, scenarios

QUERIES = scenarios.standard_queries()


@pytest.fixture(scope="module")
def golden():
    return runners.load_golden("standard_queries.npz")


@pytest.mark.parametrize("q", QUERIES, ids=[q.name for q in QUERIES])
def test_oracle_reproduces_reference(q, golden):
    g = runners.golden_case(golden, q.name)
    res = runners.run_oracle(q)
    runners.assert_oracle_matches_golden(q.name, res, g)


def test_oracle_closed_loop_calls():
    """scenario_01_cv closed loop (config 1): host conversion with the recorded _prev_s / _last_kappa,
    then the sweep, for a spread of the recorded plan() calls incl. CAUTION / EMERGENCY retries."""
    store = runners.load_golden("closed_loop_s01.npz")
    sp = O.Spline2D(store["waypoints_x"], store["waypoints_y"])
    kn = O.Knobs(**scenarios.S1_KNOBS)
    n = len(store["kept"])
    picked = sorted(set(np.linspace(0, n - 1, 14).astype(int)))
    seen_override = False
    for j in picked:
        g = runners.golden_case(store, f"c{j}")
        pl = O.OraclePlanner(sp, kn)
        pl.last_kappa = float(g["last_kappa"])
        pl.search.prev_s = None if np.isnan(g["prev_s"]) else float(g["prev_s"])
        ovr = {k: float(v) for k, v in zip(("max_speed", "max_accel", "max_curvature", "max_lat_accel"), g["ovr"])
               if not np.isnan(v)} or None
        seen_override |= ovr is not None
        msd = None if np.isnan(g["msd"]) else float(g["msd"])
        res = pl.plan(tuple(g["ego"]), np.empty((0, 2)), g["dyn"], float(g["target"]), ovr, None, msd)
        runners.assert_oracle_matches_golden(f"call {j}", res, g)
    assert seen_override


def _campaign_row(res, g, i, tag):
    from tests import campaign
    assert len(res.categories) == int(g["n_cand"][i]), (tag, i)
    assert res.best_index == int(g["best"][i]), (tag, i, res.best_index, int(g["best"][i]))
    assert campaign.crc(res.categories.astype(np.uint8)) == int(g["crc"][i]), (tag, i)
    want = g["stats"][i]
    got = [res.stats.get(k, 0) for k in runners.GOLDEN_STAT_KEYS]
    assert got[:7] == want[:7].tolist() and (want[7] < 0 or got[7] == want[7]), (tag, i, got, want)
    if res.best_index >= 0:
        np.testing.assert_allclose(res.cost, float(g["cost"][i]), rtol=1e-12)
    np.testing.assert_allclose(np.array(res.frenet_state), g["fs"][i], rtol=1e-12, atol=1e-15)


def test_oracle_reproduces_campaign_samples():
    """The port against a sample of the differential campaign recorded from the reference
    (tests/golden/make_golden_campaign.py): config-4 queries, near-limit queries, the first steps of the config-5
    rollout (stateful: nearest-point cache and ego curvature fed back)."""
    from tests import campaign
    ga, gb, gc = (runners.load_golden(f"campaign_{x}.npz") for x in "abc")
    orc = O.OraclePlanner(O.Spline2D(*scenarios.STRAIGHT_60), O.Knobs(**scenarios.S1_KNOBS))
    for i in (0, 1, 2, 3, 1234, 4999):
        ego, dyn = campaign.query_a(i)
        orc.reset_ego_curvature()
        orc.search = O.NearestPointSearch(orc.sp)
        _campaign_row(orc.plan(ego, np.empty((0, 2)), dyn, campaign.TARGET_SPEED), ga, i, "A")
    for i in list(range(0, 36)) + [777, 1999]:
        q = campaign.query_b(i)
        ob = O.OraclePlanner(O.Spline2D(*campaign.PATHS[q["path"]]), O.Knobs(**campaign.knobs_b(q["road"])))
        _campaign_row(ob.plan_frenet(tuple(q["fs"]), np.empty((0, 2)), q["dyn"], q["target"], q["overrides"], None, q["msd"]),
                      gb, i, "B")
    oc = O.OraclePlanner(O.Spline2D(*campaign.C_PATH), O.Knobs(**scenarios.S1_KNOBS))
    n = 0
    for step, ego, dyn, static in campaign.rollout_c(6):
        for target, ovr, msd in campaign.plans_c():
            st = np.empty((0, 2)) if static is None else static
            assert oc.last_kappa == float(gc["kappa_in"][n])
            _campaign_row(oc.plan(ego, st, dyn, target, ovr, None, msd), gc, n, "C")
            n += 1
