"""cProfile of BatchedClosedLoop.step (256 simulations of the recorded scenario_01 variants with jitter)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from integrated_path_planning_b200.rollout import BatchedClosedLoop
z = np.load(os.path.join(ROOT, "tests", "golden", "rollout_s01.npz"))
knobs = {k[5:]: float(z[k]) for k in z.files if k.startswith("knob/")}
rng = np.random.default_rng(7)
n, n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(z["n_variants"])
tracks = np.stack([z[f"v{i % n_var}/traj"] + rng.normal(0.0, 0.3, (1, z["v0/traj"].shape[1], 2)) for i in range(n)])
ego0 = np.stack([z[f"v{i % n_var}/ego0"] for i in range(n)])
sim = BatchedClosedLoop(z["v0/wx"], z["v0/wy"], knobs, tracks, ego0)
sim.warmup()
for _ in range(5):
    sim.step()
t0 = time.perf_counter()
for _ in range(30):
    sim.step()
print("ms per lock-step", (time.perf_counter() - t0) / 30 * 1e3, {k: round(v, 3) for k, v in sim.timers.items()})
pr = cProfile.Profile(); pr.enable()
for _ in range(30):
    sim.step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
