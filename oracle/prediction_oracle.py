"""Oracle for the predictor's post-processing -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference code between "raw predictor output / last two observations" and
the obstacle tensor the planner receives:

  predict_cv          src/prediction/trajectory_predictor.py:188-231
  process_prediction  src/prediction/trajectory_predictor.py:233-313
  select_best         src/prediction/trajectory_predictor.py:343-351 (inside predict_single_best)
  prepend_current     src/simulation/integrated_simulator.py:503-525 (inside _update_prediction)

Parity pinned: `tests/golden/make_golden_prediction.py` runs the unmodified reference predictor on
seeded inputs and `tests/test_prediction_oracle_golden.py` checks this file against those fixtures bit
for bit.  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import it; the product path
(integrated_path_planning_b200.prediction) calls the CUDA kernels and never this module.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

MAX_WALKING_SPEED = 2.5                                                     # :289


def time_grid(sim_dt: float, plan_horizon: float, pred_len: int, sgan_dt: float) -> np.ndarray:
    """:214-215 / :283-284."""
    target_horizon = max(plan_horizon, pred_len * sgan_dt)
    return np.arange(sim_dt, target_horizon + 1e-9, sim_dt)


def predict_cv(obs_traj: np.ndarray, sgan_dt: float, time_target: np.ndarray, staleness: float = 0.0) -> np.ndarray:
    """obs_traj [obs_len, P, 2] -> [P, n_steps, 2]  (:188-231).  A float32 obs_traj reproduces the simulator's
    data flow (the observer hands over float32 tensors): NumPy then keeps the velocity in float32 and promotes
    the extrapolation to float64, exactly as in the reference."""
    if obs_traj.shape[0] < 2:
        current_pos = obs_traj[-1]
        velocities = np.zeros((current_pos.shape[0], 2))
    else:
        p_curr, p_prev = obs_traj[-1], obs_traj[-2]
        velocities = (p_curr - p_prev) / sgan_dt
        current_pos = p_curr
    out = np.zeros((current_pos.shape[0], len(time_target), 2))
    for i in range(len(time_target)):
        t = time_target[i] + staleness
        out[:, i, :] = current_pos + velocities * t
    return out


def process_prediction(pred_traj: np.ndarray, sgan_dt: float, time_target: np.ndarray,
                       anchor_pos: Optional[np.ndarray] = None, staleness: float = 0.0) -> np.ndarray:
    """pred_traj [pred_len, P, 2] -> [P, n_steps, 2]  (:233-313)."""
    pred_len, n_peds, _ = pred_traj.shape
    time_src = np.arange(1, pred_len + 1) * sgan_dt - staleness
    if anchor_pos is not None:
        time_src = np.concatenate(([-staleness], time_src))
        pred_traj = np.concatenate((anchor_pos[None, ...], pred_traj), axis=0)
    dense = np.zeros((n_peds, len(time_target), 2), dtype=float)
    for ped in range(n_peds):
        traj = pred_traj[:, ped, :]
        for axis in range(2):
            coords = traj[:, axis]
            if np.allclose(coords, coords[0]) or np.allclose(coords, 0.0):
                dense[ped, :, axis] = coords[-1]
                continue
            vals = np.interp(time_target, time_src, coords)
            if len(coords) >= 2:
                lookback = min(3, len(coords))
                v_tail = (coords[-1] - coords[-lookback]) / ((lookback - 1) * sgan_dt)
                v_tail = max(min(v_tail, MAX_WALKING_SPEED), -MAX_WALKING_SPEED)
                tail = time_target > time_src[-1]
                if tail.any():
                    vals[tail] = coords[-1] + v_tail * (time_target[tail] - time_src[-1])
            dense[ped, :, axis] = vals
    return dense


def select_best(samples: np.ndarray) -> Tuple[int, np.ndarray]:
    """samples [S, P, T, 2] -> (index of the sample closest to the mean, distances [S])  (:346-349)."""
    mean_traj = samples.mean(axis=0)
    distances = np.linalg.norm(samples - mean_traj[None, ...], axis=-1).sum(axis=(1, 2))
    return int(np.argmin(distances)), distances


def prepend_current(dynamic_obstacles: np.ndarray, current_positions: np.ndarray,
                    distribution: Optional[np.ndarray] = None):
    """integrated_simulator.py:503-525: t = 0 column for the representative sample [P, T, 2] (skipped when it
    already starts at the current positions) and, unconditionally, for every distribution sample [S, P, T, 2]."""
    cur = current_positions[:, None, :]
    if dynamic_obstacles.size == 0:
        dynamic_obstacles = cur
    else:
        has = dynamic_obstacles.shape[1] >= 1 and np.allclose(dynamic_obstacles[:, 0, :], cur[:, 0, :])
        if not has:
            dynamic_obstacles = np.concatenate([cur, dynamic_obstacles], axis=1)
    if distribution is not None and distribution.size > 0:
        cur_dist = np.broadcast_to(cur[None, ...], (distribution.shape[0],) + cur.shape)
        distribution = np.concatenate([cur_dist, distribution], axis=2)
    return dynamic_obstacles, distribution


def safety_metrics(ego_xyyva, ped_pos: np.ndarray, ped_vel: np.ndarray, ego_radius: float, ped_radius: float,
                   footprint_offsets: Optional[np.ndarray] = None, footprint_radius: float = 0.0) -> dict:
    """compute_safety_metrics_static, src/core/data_structures.py:301-388."""
    x, y, yaw, v = ego_xyyva[0], ego_xyyva[1], ego_xyyva[2], ego_xyyva[3]
    if footprint_offsets is None:
        centers = np.array([[x, y]])
        combined = ego_radius + ped_radius
    else:
        direction = np.array([np.cos(yaw), np.sin(yaw)])                    # footprint.py:44-45
        centers = np.array([x, y]) + np.asarray(footprint_offsets)[:, None] * direction
        combined = footprint_radius + ped_radius
    if len(ped_pos) > 0:
        dist_matrix = np.linalg.norm(ped_pos[None, :, :] - centers[:, None, :], axis=2)
        min_distance = float(np.min(dist_matrix))
    else:
        dist_matrix = np.empty((len(centers), 0))
        min_distance = float("inf")
    collision = min_distance < combined
    ttc = float("inf")
    if len(ped_pos) > 0:
        ego_vel = np.array([v * np.cos(yaw), v * np.sin(yaw)])
        for ci, center in enumerate(centers):
            for pi, (pos, vel) in enumerate(zip(ped_pos, ped_vel)):
                rel_pos = pos - center
                rel_vel = vel - ego_vel
                along = -np.dot(rel_pos, rel_vel) / (np.linalg.norm(rel_pos) + 1e-8)
                if along > 1e-5:
                    t = (dist_matrix[ci, pi] - combined) / along
                    if t >= 0:
                        ttc = min(ttc, t)
    clearance_ahead = float("inf")
    if len(ped_pos) > 0:
        heading = np.array([np.cos(yaw), np.sin(yaw)])
        ahead = (ped_pos - np.array([x, y])) @ heading > 0.0
        if np.any(ahead):
            clearance_ahead = float(np.min(dist_matrix[:, ahead])) - combined
    return {"min_distance": min_distance, "collision": collision, "ttc": ttc, "clearance": min_distance - combined,
            "clearance_ahead": clearance_ahead}
