"""GPU parity: the CUDA sweep (through the C ABI) against the oracle on seeded queries."""
import numpy as np
import pytest

from tests import runners, scenarios

pytestmark = pytest.mark.gpu

QUERIES = scenarios.standard_queries()


@pytest.mark.parametrize("q", QUERIES, ids=[q.name for q in QUERIES])
def test_plan_matches_oracle(q):
    ref = runners.run_oracle(q)
    pl, path = runners.run_cuda(q)
    runners.assert_matches_oracle(q, ref, pl, path)
    # side effects of plan(): _last_kappa only moves on success (frenet_planner.py:301-302)
    if path is not None and len(path.c) > 1:
        assert pl._last_kappa == float(path.c[1])
    else:
        assert pl._last_kappa == q.last_kappa


def test_cost_bit_exactness_report(capsys):
    q = QUERIES[1]
    ref = runners.run_oracle(q)
    pl, path = runners.run_cuda(q)
    rep = runners.bit_exact_report(ref, pl, path)
    with capsys.disabled():
        print("\nbit-exact fractions vs oracle:", rep)
    assert rep["cost_bit_exact_frac"] > 0.5
