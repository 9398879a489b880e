"""Record golden vectors for the prediction post-processing from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_prediction.py
Writes tests/golden/prediction.npz.  The reference predictor is used as shipped: TrajectoryPredictor
(method='cv', no weights) for predict_cv / process_prediction, and predict_single_best with `predict`
replaced by a replay of fixed samples (the SGAN weights are not in the container).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
from loguru import logger  # noqa: E402

logger.remove()
from src.prediction.trajectory_predictor import TrajectoryPredictor  # noqa: E402


def cases():
    rng = np.random.default_rng(2026)
    out = []
    for i, (P, pred_len, stale, with_anchor) in enumerate([(7, 12, 0.0, True), (5, 12, 0.3, True), (9, 8, 0.1, False),
                                                           (3, 12, 0.2, True), (4, 20, 0.0, True), (6, 1, 0.0, False)]):
        start = rng.uniform(-5, 5, (P, 2))
        vel = rng.normal(0, 1.2, (P, 2))
        steps = np.arange(1, pred_len + 1)[:, None, None] * 0.4
        raw = start[None] + vel[None] * steps + rng.normal(0, 0.05, (pred_len, P, 2)).cumsum(axis=0)
        raw[:, 0, :] = start[0]                    # a standing pedestrian: the constant-fill rule
        if P > 2:
            raw[:, 1, 0] = 0.0                     # an all-zero axis (warm-up rows)
        if P > 2 and pred_len >= 4:
            raw[-3:, 2, 1] = raw[-4, 2, 1] + 4.0 * np.arange(1, 4)    # runs away at 10 m/s: the tail clamp
        out.append(dict(name=f"p{i}", raw=raw, anchor=start if with_anchor else None, staleness=stale,
                        pred_len=pred_len, obs=np.stack([start - vel * 0.4, start])))
    return out


def main():
    store = {}
    names = []
    for c in cases():
        tp = TrajectoryPredictor(model_path=None, pred_len=c["pred_len"], num_samples=1, device="cpu", sgan_dt=0.4,
                                 sim_dt=0.1, plan_horizon=5.0, method="cv")
        n = c["name"]
        names.append(n)
        store[n + "_raw"], store[n + "_obs"] = c["raw"], c["obs"]
        store[n + "_anchor"] = c["anchor"] if c["anchor"] is not None else np.zeros((0, 2))
        store[n + "_meta"] = np.array([c["pred_len"], c["staleness"]])
        store[n + "_dense"] = tp.process_prediction(c["raw"].copy(), anchor_pos=c["anchor"], staleness=c["staleness"])
        store[n + "_cv"] = tp.predict_cv(torch.from_numpy(c["obs"]), staleness=c["staleness"])
        store[n + "_cv1"] = tp.predict_cv(torch.from_numpy(c["obs"][-1:]), staleness=c["staleness"])
        # the simulator's real data flow: float32 observation tensors (observer.py:131-132)
        store[n + "_cv32"] = tp.predict_cv(torch.from_numpy(c["obs"]).float(), staleness=c["staleness"])
    # closest-to-mean selection through predict_single_best with replayed samples
    rng = np.random.default_rng(7)
    for j, (S, P, T) in enumerate([(6, 5, 50), (20, 11, 50), (3, 2, 60)]):
        samples = rng.normal(0, 1, (S, P, T, 2)).cumsum(axis=2) * 0.1
        tp = TrajectoryPredictor(model_path=None, pred_len=12, num_samples=S, device="cpu", method="cv")
        it = iter(samples)
        tp.predict = lambda *a, **k: next(it)
        best, dist = tp.predict_single_best(None, None, None)
        idx = int(np.nonzero([np.array_equal(best, s) for s in samples])[0][0])
        store[f"sel{j}_samples"], store[f"sel{j}_best"] = samples, np.array([idx])
        assert np.array_equal(dist, samples)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "prediction.npz"), **store)
    print("wrote prediction.npz:", len(names), "resampling cases, 3 selection cases")


if __name__ == "__main__":
    main()
