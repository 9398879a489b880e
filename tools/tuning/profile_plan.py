import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import runners, scenarios
from integrated_path_planning_b200 import CubicSpline2D, FrenetPlanner
pl = FrenetPlanner(CubicSpline2D(*scenarios.STRAIGHT_60), **scenarios.S1_KNOBS)
dyn = scenarios.pedestrian_field(np.random.default_rng(1), 50)
ego = runners._Ego(5.0, 0.0, 0.0, 5.0, 0.0)
st = np.empty((0, 2))
for _ in range(5): pl.plan(ego, st, dyn, 6.0)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): pl.plan(ego, st, dyn, 6.0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
