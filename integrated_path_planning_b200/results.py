"""Result files of a closed-loop run: `metrics_summary.csv` and `metrics_report.txt`
as `IntegratedSimulator.save_results` writes them (src/simulation/integrated_simulator.py:987-1065), so that the
reference's downstream aggregation / plotting scripts keep working on runs of the batched driver.

The aggregate metrics restate `calculate_aggregate_metrics` (src/core/metrics.py:287-333) on the driver's per-step
records: safety extrema, jerk / acceleration statistics, the fixed-horizon best-of-N ADE / FDE at the predictor
cadence, scene-level and per-agent (:33-118), the rolling planner-resolution ADE / FDE (:222-268) and the KDE
negative log-likelihood of the ground truth under the sample set (:121-189).  Host NumPy: this is bookkeeping
over a finished run, not part of the planning path.
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Optional

import numpy as np

KDE_BANDWIDTH_FLOOR = 0.05       # metrics.py:14
KDE_NLL_LOG_P_FLOOR = -20.0      # metrics.py:19


def _steps_for_interval(interval: float, dt: float) -> int:
    """metrics.py:22-28."""
    ratio = interval / dt
    rounded = int(round(ratio))
    if rounded <= 0 or not np.isclose(ratio, rounded):
        raise ValueError(f"Evaluation interval {interval} must be a multiple of dt={dt}")
    return rounded


def _samples_of(rec) -> Optional[np.ndarray]:
    """[S, P, T, 2] sample set of one step: the distribution when the predictor produced one, else the single
    forecast (metrics.py:60-76)."""
    dist, single = rec.get("dist"), rec.get("pred")
    if dist is not None and np.size(dist) > 0:
        return np.asarray(dist)
    if single is not None and np.size(single) > 0:
        return np.asarray(single)[None, ...]
    return None


def standard_ade_fde(history: List[dict], dt: float, prediction_dt: float, prediction_steps: int):
    """metrics.py:33-118 -> (ade, fde, ade_per_agent, fde_per_agent, max_samples, count)."""
    stride = _steps_for_interval(prediction_dt, dt)
    pred_indices = stride * np.arange(1, prediction_steps + 1) - 1
    future_offsets = stride * np.arange(1, prediction_steps + 1)
    total_ade = total_fde = total_ade_pa = total_fde_pa = 0.0
    count = max_samples = 0
    for i, rec in enumerate(history):
        samples = _samples_of(rec)
        if samples is None:
            continue
        n_samples, n_peds, dense_steps, _ = samples.shape
        if dense_steps <= pred_indices[-1] or i + future_offsets[-1] >= len(history):
            continue
        gt = np.stack([history[i + off]["ped_pos"] for off in future_offsets], axis=1)
        if gt.shape != (n_peds, prediction_steps, 2):
            continue
        disp = np.linalg.norm(samples[:, :, pred_indices, :] - gt[None, ...], axis=3)
        ade_s = np.mean(disp, axis=(1, 2))
        fde_s = np.mean(disp[:, :, -1], axis=1)
        total_ade += float(np.min(ade_s)) * n_peds
        total_fde += float(np.min(fde_s)) * n_peds
        total_ade_pa += float(np.sum(np.min(np.mean(disp, axis=2), axis=0)))
        total_fde_pa += float(np.sum(np.min(disp[:, :, -1], axis=0)))
        count += n_peds
        max_samples = max(max_samples, int(rec.get("n_samples", n_samples)))
    if count == 0:
        return float("nan"), float("nan"), float("nan"), float("nan"), 0, 0
    return total_ade / count, total_fde / count, total_ade_pa / count, total_fde_pa / count, max_samples, count


def planning_ade_fde(history: List[dict]):
    """metrics.py:222-268."""
    total_ade = total_fde = 0.0
    count = 0
    for i, rec in enumerate(history):
        pred = rec.get("pred")
        if pred is None or np.size(pred) == 0:
            continue
        n_peds, n_steps, _ = pred.shape
        eval_steps = min(n_steps, len(history) - (i + 1))
        if eval_steps == 0:
            continue
        gt = np.stack([history[i + 1 + k]["ped_pos"] for k in range(eval_steps)], axis=1)
        if gt.shape != (n_peds, eval_steps, 2):
            continue
        disp = np.linalg.norm(pred[:, :eval_steps, :] - gt, axis=2)
        total_ade += float(np.sum(np.mean(disp, axis=1)))
        total_fde += float(np.sum(disp[:, -1]))
        count += n_peds
    if count == 0:
        return float("nan"), float("nan"), 0
    return total_ade / count, total_fde / count, count


def kde_nll(history: List[dict], dt: float, prediction_dt: float, prediction_steps: int):
    """metrics.py:121-189."""
    stride = _steps_for_interval(prediction_dt, dt)
    pred_indices = stride * np.arange(1, prediction_steps + 1) - 1
    future_offsets = stride * np.arange(1, prediction_steps + 1)
    total, count = 0.0, 0
    for i, rec in enumerate(history):
        dist = rec.get("dist")
        if dist is None or np.size(dist) == 0 or dist.shape[0] < 2:
            continue
        n_samples, n_peds, dense_steps, _ = dist.shape
        if dense_steps <= pred_indices[-1] or i + future_offsets[-1] >= len(history):
            continue
        gt = np.stack([history[i + off]["ped_pos"] for off in future_offsets], axis=1)
        if gt.shape != (n_peds, prediction_steps, 2):
            continue
        samples = dist[:, :, pred_indices, :]
        if not np.any(np.ptp(samples, axis=0) > 0):
            continue
        scott = n_samples ** (-1.0 / 6.0)
        bandwidth = np.maximum(samples.std(axis=0, ddof=1) * scott, KDE_BANDWIDTH_FLOOR)
        scaled = (samples - gt[None, ...]) / bandwidth[None, ...]
        log_kernel = -0.5 * np.sum(scaled ** 2, axis=3) - np.log(2.0 * np.pi * bandwidth[..., 0] * bandwidth[..., 1])[None, ...]
        peak = log_kernel.max(axis=0)
        log_p = peak + np.log(np.mean(np.exp(log_kernel - peak[None, ...]), axis=0))
        log_p = np.maximum(log_p, KDE_NLL_LOG_P_FLOOR)
        total += float(log_p.sum())
        count += log_p.size
    if count == 0:
        return float("nan"), 0
    return -total / count, count


def aggregate_metrics(history: List[dict], dt: float, prediction_dt: float = 0.4, prediction_steps: int = 12) -> Dict[str, float]:
    """calculate_aggregate_metrics (metrics.py:287-333), same keys in the same order.

    history: one record per step with `min_distance`, `collision`, `ttc`, `ego` (x, y, yaw, v, a), `jerk`, `ped_pos`
    [P, 2], `pred` [P, T, 2] or None, `dist` [S, P, T, 2] or None, `n_samples`."""
    min_d = [r["min_distance"] for r in history]
    ttc_valid = [r["ttc"] for r in history if r["ttc"] > 0 and r["ttc"] != float("inf")]
    jerks = [abs(r["jerk"]) for r in history]
    accels = [abs(r["ego"][4]) for r in history]
    ade, fde, ade_pa, fde_pa, n_samples, ade_n = standard_ade_fde(history, dt, prediction_dt, prediction_steps)
    p_ade, p_fde, p_n = planning_ade_fde(history)
    nll, nll_n = kde_nll(history, dt, prediction_dt, prediction_steps)
    return {
        "min_dist": min(min_d) if min_d else 0.0,
        "collision_count": sum(1 for r in history if r.get("collision", False)),
        "min_ttc": min(ttc_valid) if ttc_valid else float("inf"),
        "max_jerk": max(jerks) if jerks else 0.0,
        "mean_jerk": np.mean(jerks) if jerks else 0.0,
        "rms_jerk": float(np.sqrt(np.mean(np.square(jerks)))) if jerks else 0.0,
        "max_accel": max(accels) if accels else 0.0,
        "mean_accel": np.mean(accels) if accels else 0.0,
        "ade": ade, "fde": fde, "ade_per_agent": ade_pa, "fde_per_agent": fde_pa,
        "pred_samples": n_samples, "ade_eval_count": ade_n,
        "planning_ade": p_ade, "planning_fde": p_fde, "planning_eval_count": p_n,
        "nll": nll, "nll_eval_count": nll_n,
    }


def write_metrics_files(output_dir: str, context: Dict[str, object], metrics: Dict[str, object], collision: bool):
    """metrics_summary.csv (one header row, one data row: context, metrics, `collision`) and metrics_report.txt,
    laid out as integrated_simulator.py:1019-1065 writes them."""
    os.makedirs(output_dir, exist_ok=True)
    row = dict(context)
    row.update(metrics)
    if "collision" not in row:
        row["collision"] = bool(collision)
    csv_path = os.path.join(output_dir, "metrics_summary.csv")
    with open(csv_path, "w", newline="") as f:
        writer = csv.DictWriter(f, fieldnames=row.keys())
        writer.writeheader()
        writer.writerow(row)
    txt_path = os.path.join(output_dir, "metrics_report.txt")
    with open(txt_path, "w") as f:
        f.write("=" * 40 + "\n")
        f.write("       SIMULATION REPORT\n")
        f.write("=" * 40 + "\n\n")
        f.write("--- Configuration ---\n")
        for k, v in context.items():
            f.write(f"{k}: {v}\n")
        f.write("\n")
        f.write("--- Metrics ---\n")
        for k, v in metrics.items():
            f.write(f"{k}: {v}\n")
        f.write("\n")
        if not metrics:
            f.write("No detailed metrics available.\n")
        f.write("=" * 40 + "\n")
    return csv_path, txt_path
