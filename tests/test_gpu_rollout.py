"""Batched closed-loop roll-outs against the reference simulator's recorded runs.

tests/golden/rollout_s01.npz holds four closed-loop runs of the UNMODIFIED reference IntegratedSimulator
(scenario_01_cv and three perturbed variants; tests/golden/make_golden_rollout.py): per step the ego state,
the fail-safe state, whether a path was found and how many plan() calls the step made, plus how the run
ended.  The batched driver advances all four in lock-step, one sweep launch per planning attempt, and has
to reproduce every one of them -- through the fail-safe escalations, emergency stops and the collision
terminations of variants 1 and 2 -- with the feedback loop closed over ITS OWN outputs.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
# ego states feed back into the next step's query for hundreds of steps; the sweep's points agree with the
# reference to ~1e-12 per call (tests/test_gpu_golden.py), so the closed loop is held to 1e-7 absolute.
EGO_TOL = 1e-7


def _load(name):
    z = np.load(os.path.join(HERE, "golden", name))
    knobs = {k[5:]: float(z[k]) for k in z.files if k.startswith("knob/")}
    n = int(z["n_variants"])
    runs = [{name: z[f"v{i}/{name}"] for name in ("traj", "ego0", "ego", "fsm", "found", "calls", "wx", "wy", "reason")}
            for i in range(n)]
    static = z["static_points"] if "static_points" in z.files else None
    return knobs, runs, static


@pytest.fixture(scope="module")
def golden():
    return _load("rollout_s01.npz")[:2]


def _drive(knobs, runs, static=None):
    from integrated_path_planning_b200.rollout import BatchedClosedLoop
    tracks = np.stack([r["traj"] for r in runs])
    ego0 = np.stack([r["ego0"] for r in runs])
    sim = BatchedClosedLoop(runs[0]["wx"], runs[0]["wy"], knobs, tracks, ego0, static_obstacles=static)
    sim.warmup()
    return sim, sim.run()


@pytest.mark.parametrize("name", ["rollout_s01.npz", "rollout_s02.npz", "rollout_s02fp.npz", "rollout_s03.npz"])
def test_batch_reproduces_reference_rollouts(name):
    """s01: open road, 14 pedestrians (four variants); s02: corridor between two static walls; s02fp: the same with
    the three-circle footprint (planner collision test, safety metrics and state-machine radii); s03: right turn."""
    knobs, runs, static = _load(name)
    sim, out = _drive(knobs, runs, static)
    for i, r in enumerate(runs):
        n = len(r["ego"])
        assert out["steps"][i] == n, f"variant {i}: {out['steps'][i]} steps, reference {n}"
        assert out["reason"][i] == str(r["reason"]), f"variant {i}"
        np.testing.assert_array_equal(out["found"][i, :n], r["found"], err_msg=f"variant {i} found")
        np.testing.assert_array_equal(out["fsm"][i, :n], r["fsm"], err_msg=f"variant {i} fail-safe state")
        np.testing.assert_array_equal(out["calls"][i, :n], r["calls"], err_msg=f"variant {i} plan() calls")
        np.testing.assert_allclose(out["ego"][i, :n], r["ego"], rtol=0, atol=EGO_TOL, err_msg=f"variant {i} ego")
    assert sim.n_plan_calls == sum(int(r["calls"].sum()) for r in runs)


def test_single_rollout_equals_its_row_in_the_batch(golden):
    """Simulations in a batch are independent: variant 1 alone gives what it gives inside the batch of four."""
    knobs, runs = golden
    _, alone = _drive(knobs, runs[1:2])
    _, batch = _drive(knobs, runs)
    n = int(alone["steps"][0])
    assert n == int(batch["steps"][1])
    np.testing.assert_array_equal(alone["ego"][0, :n], batch["ego"][1, :n])
    np.testing.assert_array_equal(alone["fsm"][0, :n], batch["fsm"][1, :n])


def test_result_file_matches_the_reference_writer(tmp_path):
    """trajectory.npz as IntegratedSimulator.save_results writes it (integrated_simulator.py:906-985): 80 steps of
    scenario_01_cv through the reference and its own writer are the golden file; the batched driver's writer must
    produce the same keys, dtypes and shapes and the same values (wall-clock timing columns aside)."""
    g = np.load(os.path.join(HERE, "golden", "rollout_s01_trajectory.npz"))
    knobs, runs, _ = _load("rollout_s01.npz")
    from integrated_path_planning_b200.rollout import BatchedClosedLoop
    sim = BatchedClosedLoop(runs[0]["wx"], runs[0]["wy"], knobs, runs[0]["traj"][None], runs[0]["ego0"][None], record=True)
    sim.warmup()
    n = len(g["times"])
    for _ in range(n):
        sim.step()
    z = np.load(sim.save_results(0, str(tmp_path)), allow_pickle=True)
    assert sorted(z.files) == g["keys"].tolist()
    for key in z.files:
        got = z[key]
        if key in ("proc_prediction", "proc_planning"):
            assert got.shape == (n,)
            continue
        if key == "ego_state":
            assert got.dtype.kind == "U" and got.tolist() == g[key].tolist()
            continue
        if got.dtype != object:
            assert got.shape == g[key].shape and got.dtype == g[key].dtype, key
            fin = np.isfinite(g[key])
            assert np.array_equal(np.isfinite(got), fin), key
            np.testing.assert_allclose(got[fin], g[key][fin], rtol=1e-9, atol=EGO_TOL, err_msg=key)
            continue
        # np.array(list_of_equal_arrays, dtype=object) keeps the full shape, ragged rows give a 1-D object array --
        # as recorded by tests/golden/make_golden_rollout.py from the reference's file
        top = {"ped_positions": (n, 14, 2), "ped_velocities": (n, 14, 2), "ped_goals": (n, 14, 2),
               "predicted_trajectories": (n, 14, 50, 2)}.get(key, (n,))
        assert got.shape == top, (key, got.shape)
        rows = [np.asarray(r, dtype=float) for r in got]
        shapes, want = g[key + "/shape"], g[key]
        for i in range(len(want)):                      # (predictions: the golden file keeps the first few steps)
            shape = tuple(int(v) for v in shapes[i][:rows[i].ndim])
            assert rows[i].shape == shape, (key, i, rows[i].shape, shape)
            np.testing.assert_allclose(rows[i].reshape(-1), want[i][:rows[i].size], rtol=1e-9, atol=EGO_TOL, err_msg=f"{key}[{i}]")
        assert len(rows) == n
