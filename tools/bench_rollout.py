#!/usr/bin/env python
"""Closed-loop campaign throughput: N simulations advanced in lock-step by BatchedClosedLoop.

    python tools/bench_rollout.py --sims 256 --steps 60

The four recorded scenario_01 variants (tests/golden/rollout_s01.npz) are tiled to N simulations, each with
its own jitter on the pedestrian tracks and the ego start speed, and stepped K times.  Prints one JSON line:
simulation steps per second (all simulations), plan() calls per second, and where the wall time went
(host Frenet conversion, sweep launches incl. result read-back, prediction, safety metrics).  The reference
runs the same campaign as a sequential loop at one plan() call per ~0.2 s of CPU (bench.py cpu_baseline).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sims", type=int, default=256)
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    from integrated_path_planning_b200.rollout import BatchedClosedLoop
    z = np.load(os.path.join(ROOT, "tests", "golden", "rollout_s01.npz"))
    knobs = {k[5:]: float(z[k]) for k in z.files if k.startswith("knob/")}
    rng = np.random.default_rng(7)
    n_var = int(z["n_variants"])
    tracks, ego0 = [], []
    for i in range(a.sims):
        v = i % n_var
        t = z[f"v{v}/traj"]
        tracks.append(t + rng.normal(0.0, 0.3, (1, t.shape[1], 2)))
        e = z[f"v{v}/ego0"].copy()
        e[3] = max(0.5, e[3] + rng.normal(0.0, 0.3))
        ego0.append(e)
    sim = BatchedClosedLoop(z["v0/wx"], z["v0/wy"], knobs, np.stack(tracks), np.stack(ego0))
    sim.warmup()
    for _ in range(3):
        sim.step()
    for k in sim.timers:
        sim.timers[k] = 0.0
    calls0, t0, sim_steps = sim.n_plan_calls, time.perf_counter(), 0
    for _ in range(a.steps):
        sim_steps += int(sim.active.sum())
        sim.step()
    wall = time.perf_counter() - t0
    print(json.dumps({"metric": "closed_loop_sim_steps_per_s", "value": sim_steps / wall, "sims": a.sims, "steps": a.steps,
                      "plan_calls_per_s": (sim.n_plan_calls - calls0) / wall, "still_active": int(sim.active.sum()),
                      "ms_per_lockstep": 1e3 * wall / a.steps,
                      "share": {k: round(v / wall, 3) for k, v in sim.timers.items()}}))


if __name__ == "__main__":
    main()
