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


def _sgan_knobs(z):
    return {k[5:]: float(z[k]) for k in z.files if k.startswith("sgan/")}


@pytest.mark.parametrize("name", ["rollout_s01_dist.npz", "rollout_s01_best.npz"])
def test_batch_reproduces_sample_set_rollouts(name):
    """Distribution-aware planning in the batched driver.  The reference ran scenario_01 with a seeded
    stand-in generator in its unmodified predictor (tests/stub_sampler.py): S raw samples per step -> process_prediction
    -> closest-to-mean sample -> t = 0 column (trajectory_predictor.py:233-353, integrated_simulator.py:503-525), and
    planned against the WHOLE sample set under chance_epsilon = 0.2 (`_dist`, 6 samples) or against the representative
    sample (`_best`, 4 samples).  The driver gets the same raw samples from `sampler=` and does everything behind the
    generator on the device; it must reproduce each run step for step."""
    from integrated_path_planning_b200.rollout import BatchedClosedLoop
    from tests.stub_sampler import batched_sampler
    z = np.load(os.path.join(HERE, "golden", name))
    knobs, runs, _ = _load(name)
    sg = _sgan_knobs(z)
    assert int(knobs["num_samples"]) == int(sg["num_samples"]) and bool(knobs["distribution_aware_planning"]) == bool(sg["distribution_aware"])
    seeds = [int(sg["gen_seed"]) + 1000 * i for i in range(len(runs))]
    sampler = batched_sampler(seeds, int(sg["num_samples"]), int(knobs["pred_len"]), float(sg["sigma"]))
    tracks = np.stack([r["traj"] for r in runs])
    ego0 = np.stack([r["ego0"] for r in runs])
    sim = BatchedClosedLoop(runs[0]["wx"], runs[0]["wy"], knobs, tracks, ego0, sampler=sampler)
    assert sim.distribution_aware == bool(sg["distribution_aware"])
    sim.warmup()
    out = sim.run()
    for i, r in enumerate(runs):
        n = len(r["ego"])
        assert out["steps"][i] == n, f"variant {i}: {out['steps'][i]} steps, reference {n}"
        assert out["reason"][i] == str(r["reason"]), f"variant {i}"
        np.testing.assert_array_equal(out["found"][i, :n], r["found"], err_msg=f"variant {i} found")
        np.testing.assert_array_equal(out["fsm"][i, :n], r["fsm"], err_msg=f"variant {i} fail-safe state")
        np.testing.assert_array_equal(out["calls"][i, :n], r["calls"], err_msg=f"variant {i} plan() calls")
        np.testing.assert_allclose(out["ego"][i, :n], r["ego"], rtol=0, atol=EGO_TOL, err_msg=f"variant {i} ego")


def _metric_rows(csv_text):
    import csv
    import io
    rows = list(csv.reader(io.StringIO(csv_text)))
    assert len(rows) == 2
    return rows[0], dict(zip(rows[0], rows[1]))


@pytest.mark.parametrize("name,with_sampler", [("rollout_s01_trajectory.npz", False), ("rollout_s01_dist_trajectory.npz", True)])
def test_metrics_files_match_the_reference_writer(tmp_path, name, with_sampler):
    """metrics_summary.csv / metrics_report.txt (integrated_simulator.py:1019-1065; aggregate metrics of
    src/core/metrics.py:287-333): 80 steps of scenario_01 through the reference and its own writer are the golden
    text -- once with the CV predictor (20 identical samples: nll undefined), once with the seeded stand-in generator
    (6 samples, distribution-aware planning: KDE NLL and best-of-N ADE / FDE over real sample sets).  Same columns in
    the same order, same values; the four wall-clock columns are this driver's own."""
    from integrated_path_planning_b200.rollout import BatchedClosedLoop
    g = np.load(os.path.join(HERE, "golden", name))
    knobs, runs, _ = _load("rollout_s01.npz")
    sampler = None
    if with_sampler:
        from tests.stub_sampler import batched_sampler
        sg = _sgan_knobs(g)
        knobs = dict(knobs, num_samples=sg["num_samples"], distribution_aware_planning=sg["distribution_aware"],
                     chance_epsilon=sg["chance_epsilon"], pred_len=12)
        sampler = batched_sampler([int(sg["gen_seed"])], int(sg["num_samples"]), 12, float(sg["sigma"]))
    else:
        knobs = dict(knobs, num_samples=20)                        # scenario_01_cv.yaml: the CV forecast replicated 20 times
    sim = BatchedClosedLoop(runs[0]["wx"], runs[0]["wy"], knobs, runs[0]["traj"][None], runs[0]["ego0"][None], record=True,
                            sampler=sampler)
    sim.warmup()
    for _ in range(len(g["times"])):
        sim.step()
    sim.save_results(0, str(tmp_path), context={"prediction_method": "cv"})    # (the reference's config said cv in both runs)
    want_cols, want = _metric_rows(str(g["metrics_csv"]))
    got_cols, got = _metric_rows(open(os.path.join(str(tmp_path), "metrics_summary.csv")).read())
    assert got_cols == want_cols
    clock = {"avg_prediction_time", "max_prediction_time", "avg_planning_time", "max_planning_time"}
    for col in want_cols:
        if col in clock:
            assert float(got[col]) >= 0.0
            continue
        try:
            w, v = float(want[col]), float(got[col])
        except ValueError:
            assert got[col] == want[col], (col, got[col], want[col])
            continue
        if np.isnan(w):
            assert np.isnan(v), col
        else:
            assert v == pytest.approx(w, rel=1e-6, abs=1e-9), (col, v, w)
    # the text report: same lines, same order, numbers as above
    want_lines = str(g["metrics_txt"]).splitlines()
    got_lines = open(os.path.join(str(tmp_path), "metrics_report.txt")).read().splitlines()
    assert len(got_lines) == len(want_lines)
    for a, b in zip(got_lines, want_lines):
        assert a.split(":")[0] == b.split(":")[0], (a, b)
    if with_sampler:
        assert int(got["pred_samples"]) == 6 and not np.isnan(float(got["nll"])) and int(got["nll_eval_count"]) > 0
