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
