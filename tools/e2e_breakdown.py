import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tests import scenarios
from integrated_path_planning_b200 import BatchFrenetPlanner, _lib
from integrated_path_planning_b200 import engine as E
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spline, frenet, dyn = bench.make_queries(0, Q)
pl = BatchFrenetPlanner(spline, **scenarios.S1_KNOBS)
dynp = torch.from_numpy(dyn).pin_memory().numpy()
for _ in range(3): r = pl.plan_batch(frenet, 6.0, dynamic_obstacles=dynp[:, 0])
orig = pl.engine.lib.fot_plan_batch_host
tc = []
def timed(*a):
    t0 = time.perf_counter(); rc = orig(*a); tc.append(time.perf_counter() - t0); return rc
pl.engine.lib.fot_plan_batch_host = timed
tt = []
for _ in range(10):
    t0 = time.perf_counter(); r = pl.plan_batch(frenet, 6.0, dynamic_obstacles=dynp[:, 0]); tt.append(time.perf_counter() - t0)
print(f"plan_batch total {1e3*np.median(tt):.2f} ms; C call {1e3*np.median(tc):.2f} ms; python overhead {1e3*(np.median(tt)-np.median(tc)):.2f} ms")
t0 = time.perf_counter(); E.speed_grid_batch(np.full(Q, 6.0), 5/3.6); print("speed_grid_batch ms", 1e3*(time.perf_counter()-t0))
st = np.array([pl.engine.launch_stage_ms(b) for b in range(8)])
print("per-chunk stage ms (prepass, sweep, winner):\n", np.round(st[::-1], 3), "\nsum", st.sum())
