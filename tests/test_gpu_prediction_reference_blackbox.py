"""GPU: the reference's own predictor tests (tests/test_prediction_anchor.py, the post-processing part of
tests/test_trajectory_predictor.py), restated against the device post-processor.  Same set-ups and
tolerances; citations give the reference test each one follows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIM_DT, SGAN_DT, PLAN_HORIZON = 0.1, 0.4, 5.0


def _post(pred_len=12):
    from integrated_path_planning_b200.prediction import DevicePredictionPostprocessor
    return DevicePredictionPostprocessor(pred_len=pred_len, sgan_dt=SGAN_DT, sim_dt=SIM_DT, plan_horizon=PLAN_HORIZON)


def _cv_raw(p0, v):
    """Raw SGAN-like prediction of a constant-velocity pedestrian: point k is k * sgan_dt after the anchor."""
    return np.stack([p0 + v * (k * SGAN_DT) for k in range(1, 13)], axis=0)[:, None, :]     # (12, 1, 2)


def _dense(staleness, anchor=True):
    p0, v = np.array([2.0, -1.0]), np.array([1.2, 0.5])
    out = _post().process_prediction(_cv_raw(p0, v)[None, None], p0[None, None] if anchor else None, staleness)
    return out.cpu().numpy()[0, 0], p0, v


def test_reanchored_grid_matches_true_future():
    """test_prediction_anchor.py:48-66 -- dense[k] is the position at current time + (k+1) sim_dt for every
    observation phase."""
    for j in range(4):
        staleness = j * SIM_DT
        dense, p0, v = _dense(staleness)
        support_end = 12 * SGAN_DT - staleness
        for k in range(dense.shape[1]):
            t = (k + 1) * SIM_DT
            if t > support_end:
                break
            np.testing.assert_allclose(dense[0, k], p0 + v * (t + staleness), atol=1e-9)


def test_no_left_clamp_with_anchor():
    """:68-74 -- the first sim steps interpolate from the anchor."""
    dense, p0, v = _dense(0.0)
    np.testing.assert_allclose(dense[0, 0], p0 + v * 0.1, atol=1e-9)
    np.testing.assert_allclose(dense[0, 2], p0 + v * 0.3, atol=1e-9)


def test_tail_extrapolation_continues_velocity():
    """:76-85 -- beyond the shifted prediction support the tail continues at the clamped tail velocity."""
    dense, p0, v = _dense(0.3)
    k_last = dense.shape[1] - 1
    np.testing.assert_allclose(dense[0, k_last], p0 + v * ((k_last + 1) * SIM_DT + 0.3), atol=1e-9)


def test_zero_staleness_no_anchor_backward_compatible():
    """:87-97."""
    p0, v = np.array([0.0, 0.0]), np.array([1.0, 0.0])
    dense = _post().process_prediction(_cv_raw(p0, v)[None, None]).cpu().numpy()[0, 0]
    np.testing.assert_allclose(dense[0, 3], p0 + v * 0.4, atol=1e-9)


def test_cv_origin_shifted_by_staleness():
    """:101-117 (float32 observation tensors there; the tolerance is the reference's)."""
    pp = _post()
    p_prev = np.array([[0.0, 0.0]], dtype=np.float32).astype(np.float64)
    p_curr = np.array([[0.48, 0.0]], dtype=np.float32).astype(np.float64)
    for j in range(4):
        staleness = j * SIM_DT
        dense = pp.predict_cv(p_curr[None], p_prev[None], staleness).cpu().numpy()[0, 0]
        for k in (0, 9, 49):
            t = (k + 1) * SIM_DT
            np.testing.assert_allclose(dense[0, k], np.array([0.48, 0.0]) + np.array([1.2, 0.0]) * (t + staleness), atol=1e-6)


def test_dense_prediction_matches_truth_at_all_phases():
    """:126-164 -- an observer sampling every 0.4 s driven at the 0.1 s sim cadence: for a constant-velocity
    pedestrian the prediction matches the true future at every step, whatever the sampling phase."""
    speed = np.array([1.2, -0.4])
    pos = lambda t: np.array([[speed[0] * t, speed[1] * t]])
    pp = _post()
    samples, sample_t, acc, t = [], [], 0.0, 0.0
    for step in range(40):                                 # the observer's sampling rule (observer.py:52-86)
        t = round(t + SIM_DT, 9)
        acc += SIM_DT
        if acc + 1e-9 >= SGAN_DT:
            samples.append(pos(t)); sample_t.append(t)
            acc = max(acc - SGAN_DT, 0.0)
        if step < 32:
            continue
        staleness = t - sample_t[-1]
        dense = pp.predict_cv(samples[-1][None], samples[-2][None], staleness).cpu().numpy()[0, 0]
        for k in (0, 3, 19, 39):
            np.testing.assert_allclose(dense[0, k], pos(t + (k + 1) * SIM_DT)[0], atol=1e-4)


def test_process_prediction_shapes_and_horizon():
    """tests/test_trajectory_predictor.py: the dense grid covers max(plan_horizon, pred_len * sgan_dt) at sim_dt
    whatever the staleness (the length must not depend on it)."""
    for pred_len, want in ((12, 50), (8, 50), (20, 80)):
        pp = _post(pred_len)
        assert pp.n_steps == want
        raw = np.zeros((1, 1, pred_len, 3, 2)) + np.arange(pred_len)[None, None, :, None, None]
        for stale in (0.0, 0.3):
            assert tuple(pp.process_prediction(raw, None, stale).shape) == (1, 1, 3, want, 2)
