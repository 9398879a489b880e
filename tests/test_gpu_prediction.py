"""GPU: the prediction post-processing kernels against the golden vectors
recorded from the reference and against the oracle on batched random inputs -- bit for bit -- and the
whole device path "observations -> obstacle tensor -> sweep" against the host path."""
import os

import numpy as np
import pytest

from oracle import prediction_oracle as PO
from tests import scenarios

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prediction.npz"))
NAMES = [str(n) for n in G["names"]]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _post(pred_len=12, **kw):
    from integrated_path_planning_b200.prediction import DevicePredictionPostprocessor
    return DevicePredictionPostprocessor(pred_len=pred_len, **kw)


def test_golden_resampling_and_cv():
    for n in NAMES:
        pred_len, stale = int(G[n + "_meta"][0]), float(G[n + "_meta"][1])
        pp = _post(pred_len)
        anchor = G[n + "_anchor"][None] if G[n + "_anchor"].size else None
        dense = pp.process_prediction(G[n + "_raw"][None, None], anchor, stale).cpu().numpy()[0, 0]
        assert np.array_equal(_bits(dense), _bits(G[n + "_dense"])), n
        obs = G[n + "_obs"]
        cv = pp.predict_cv(obs[-1][None], obs[-2][None], stale).cpu().numpy()[0, 0]
        assert np.array_equal(_bits(cv), _bits(G[n + "_cv"])), n
        cv1 = pp.predict_cv(obs[-1][None], None, stale).cpu().numpy()[0, 0]
        assert np.array_equal(_bits(cv1), _bits(G[n + "_cv1"])), n
        cv32 = pp.predict_cv(obs[-1][None], obs[-2][None], stale, obs_float32=True).cpu().numpy()[0, 0]
        assert np.array_equal(_bits(cv32), _bits(G[n + "_cv32"])), n


def test_golden_selection():
    pp = _post()
    for j in range(3):
        s = G[f"sel{j}_samples"]
        best, dist = pp.select_best(s[None])
        assert int(best[0]) == int(G[f"sel{j}_best"][0])
        assert np.array_equal(_bits(dist.cpu().numpy()[0]), _bits(PO.select_best(s)[1]))


def test_batched_pipeline_matches_oracle():
    """64 queries x 8 samples: resample, select, both prepend rules (incl. queries whose prediction already
    starts at the current positions), per-query staleness."""
    rng = np.random.default_rng(5)
    n_q, S, P, L = 64, 8, 13, 12
    pp = _post(L)
    anchor = rng.uniform(-10, 10, (n_q, P, 2))
    vel = rng.normal(0, 1.0, (n_q, 1, 1, P, 2))
    steps = (np.arange(1, L + 1) * 0.4)[None, None, :, None, None]
    pred = anchor[:, None, None] + vel * steps + rng.normal(0, 0.03, (n_q, S, L, P, 2)).cumsum(axis=2)
    pred[::7] = anchor[::7, None, None]                       # standing crowds: constant fill + skipped prepend
    stale = rng.choice([0.0, 0.1, 0.2, 0.3], n_q)
    cur = anchor + (vel[:, 0, 0] * stale[:, None, None])
    cur[::7] = anchor[::7]
    single, dist, best = pp.obstacles_from_samples(pred, anchor, stale, cur)
    single, dist, best = single.cpu().numpy(), dist.cpu().numpy(), best.cpu().numpy()
    skipped = 0
    for q in range(n_q):
        dense = np.stack([PO.process_prediction(pred[q, s], 0.4, pp.time_target_host, anchor[q], float(stale[q]))
                          for s in range(S)])
        idx, _ = PO.select_best(dense)
        assert int(best[q]) == idx, q
        one, many = PO.prepend_current(dense[idx], cur[q], dense)
        assert np.array_equal(_bits(dist[q]), _bits(many)), q
        if one.shape[1] == dense.shape[2]:                   # prepend skipped: the device tensor repeats the last step
            skipped += 1
            one = np.concatenate([one, one[:, -1:]], axis=1)
        assert np.array_equal(_bits(single[q, 0]), _bits(one)), q
    assert skipped > 0


def test_cv_on_device_feeds_the_sweep_like_the_host_path():
    """Observations -> CV tensor -> fot_plan_batch_device, all on the device, against the host route
    (oracle CV tensor with the reference's prepend, uploaded through plan_batch)."""
    import torch
    import bench
    from integrated_path_planning_b200 import BatchFrenetPlanner, DeviceBatch, _lib
    n_q, P = 96, 50
    spline, frenet, _ = bench.make_queries(300, n_q)
    rng = np.random.default_rng(9)
    p_curr = np.stack([rng.uniform(5, 45, (n_q, P)), rng.uniform(-10, 10, (n_q, P))], axis=-1)
    vel = rng.normal(0, 1.0, (n_q, P, 2))
    vel[::5] = 0.0                                            # standing crowds: the skipped-prepend branch
    p_prev = p_curr - vel * 0.4
    stale = rng.choice([0.0, 0.1, 0.3], n_q)
    cur = p_curr + vel * stale[:, None, None]
    pp = _post()
    dyn_dev = pp.predict_cv(p_curr, p_prev, stale, cur)                     # [n_q, 1, P, 51, 2]
    assert dyn_dev.shape == (n_q, 1, P, 51, 2)
    planner = BatchFrenetPlanner(spline, **scenarios.S1_KNOBS)
    batch = DeviceBatch(planner, frenet, 6.0, dyn_dev, _lib.FOT_DYN_SINGLE)
    batch.launch(None)
    got_idx, got_cost = batch.out["best_idx"].cpu().numpy(), batch.out["best_cost"].cpu().numpy()
    host = np.empty((n_q, 1, P, 51, 2))
    for q in range(n_q):
        cv = PO.predict_cv(np.stack([p_prev[q], p_curr[q]]), 0.4, pp.time_target_host, float(stale[q]))
        one, _ = PO.prepend_current(cv, cur[q])
        if one.shape[1] == 50:
            one = np.concatenate([one, one[:, -1:]], axis=1)  # what the sweep's time-index clamp sees
        host[q, 0] = one
    assert np.array_equal(_bits(dyn_dev.cpu().numpy()), _bits(host))
    ref = planner.plan_batch(frenet, 6.0, dynamic_obstacles=host[:, 0])
    assert np.array_equal(got_idx, ref.best_idx) and np.array_equal(_bits(got_cost), _bits(ref.best_cost))
    assert (got_idx >= 0).any()


def test_safety_metrics_match_reference_golden():
    """compute_safety_metrics_static on the device against values recorded from the unmodified reference:
    distances and clearances bit for bit in single-circle mode; the quantities that pass through cos / sin /
    a dot product (ttc, footprint centres) within 1e-12 relative."""
    from integrated_path_planning_b200.prediction import safety_metrics
    from tests import runners
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "safety.npz"))
    fp = runners.make_footprint((4.5, 2.0, 3))
    assert np.array_equal(fp.offsets, g["fp_offsets"]) and fp.radius == float(g["fp_radius"][0])
    for footprint, want in ((None, g["single"]), (fp, g["footprint"])):
        m = safety_metrics(g["ego"], g["pos"], g["vel"], 1.0, 0.2, footprint, g["n_peds"])
        got = np.stack([m["min_distance"].cpu().numpy(), m["collision"].cpu().numpy().astype(float), m["ttc"].cpu().numpy(),
                        m["clearance"].cpu().numpy(), m["clearance_ahead"].cpu().numpy()], axis=1)
        assert np.array_equal(got[:, 1], want[:, 1])                                   # collision flags
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12, atol=1e-12)
        if footprint is None:
            for col in (0, 3, 4):
                assert np.array_equal(_bits(got[:, col]), _bits(want[:, col]))
        assert want[3, 1] == 1.0 and np.isinf(want[0, 0])                             # a collision and an empty scene are covered
