"""CPU: the prediction-post-processing oracle against vectors recorded from the unmodified reference
(tests/golden/make_golden_prediction.py), bit for bit."""
import os

import numpy as np

from oracle import prediction_oracle as PO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prediction.npz"))
NAMES = [str(n) for n in G["names"]]


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_process_prediction_and_cv_match_reference():
    for n in NAMES:
        pred_len, stale = int(G[n + "_meta"][0]), float(G[n + "_meta"][1])
        tt = PO.time_grid(0.1, 5.0, pred_len, 0.4)
        anchor = G[n + "_anchor"] if G[n + "_anchor"].size else None
        assert _same(PO.process_prediction(G[n + "_raw"], 0.4, tt, anchor, stale), G[n + "_dense"]), n
        assert _same(PO.predict_cv(G[n + "_obs"], 0.4, tt, stale), G[n + "_cv"]), n
        assert _same(PO.predict_cv(G[n + "_obs"][-1:], 0.4, tt, stale), G[n + "_cv1"]), n
        assert _same(PO.predict_cv(G[n + "_obs"].astype(np.float32), 0.4, tt, stale), G[n + "_cv32"]), n


def test_select_best_matches_reference():
    for j in range(3):
        idx, _ = PO.select_best(G[f"sel{j}_samples"])
        assert idx == int(G[f"sel{j}_best"][0])


def test_prepend_rules():
    rng = np.random.default_rng(0)
    cur = rng.normal(size=(4, 2))
    pred = cur[:, None, :] + rng.normal(size=(4, 10, 2))
    dist = rng.normal(size=(3, 4, 10, 2))
    one, many = PO.prepend_current(pred, cur, dist)
    assert one.shape == (4, 11, 2) and many.shape == (3, 4, 11, 2)
    assert np.array_equal(one[:, 0], cur) and np.array_equal(many[:, :, 0], np.broadcast_to(cur, (3, 4, 2)))
    pred[:, 0] = cur + 1e-9                                  # already starts at the current positions
    one, _ = PO.prepend_current(pred, cur)
    assert one.shape == (4, 10, 2)


def test_safety_metrics_match_reference():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "safety.npz"))
    keys = ("min_distance", "collision", "ttc", "clearance", "clearance_ahead")
    for use_fp, want in ((False, g["single"]), (True, g["footprint"])):
        for q in range(len(g["ego"])):
            k = int(g["n_peds"][q])
            m = PO.safety_metrics(g["ego"][q], g["pos"][q, :k], g["vel"][q, :k], 1.0, 0.2,
                                  g["fp_offsets"] if use_fp else None, float(g["fp_radius"][0]))
            got = np.array([float(m[key]) for key in keys])
            assert np.array_equal(got.view(np.uint64), want[q].view(np.uint64)), (use_fp, q, got, want[q])
