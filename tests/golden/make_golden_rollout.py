#!/usr/bin/env python
"""Record closed-loop roll-outs of the UNMODIFIED reference simulator (BASELINE config 1 and variants).

Build container only (imports /root/reference read-only):   python tests/golden/make_golden_rollout.py

scenarios/scenario_01_cv.yaml is run through the reference's IntegratedSimulator exactly as
tests/golden/make_golden.py does (pysocialforce stubbed, pedestrians replayed at constant velocity through
ReplayPedestrianSource, CV predictor), for the scenario itself and three perturbed variants (pedestrian
start positions / speeds, ego start speed).  Per step the ego state, the fail-safe state, whether a path was
found and how many plan() calls the step made are stored in tests/golden/rollout_s01.npz, together with the
inputs (pedestrian tracks, knobs) the batched driver needs to repeat the run.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FOT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("pysocialforce", types.ModuleType("pysocialforce"))

from loguru import logger  # noqa: E402

logger.remove()

KNOB_KEYS = ("num_samples", "distribution_aware_planning", "pred_len", "dt", "total_time", "obs_len", "ego_target_speed", "ego_max_speed", "ego_max_accel", "ego_max_curvature",
             "ego_max_lat_accel", "ego_radius", "ped_radius", "obstacle_radius", "d_road_w", "max_road_width", "min_t",
             "max_t", "d_t_s", "k_j", "k_t", "k_d", "k_s_dot", "k_lat", "k_lon",
             "state_machine_trigger_clearance_caution", "state_machine_trigger_time_headway",
             "state_machine_recover_clearance_caution", "state_machine_recover_clearance_emergency",
             "state_machine_safe_distance_caution", "state_machine_safe_distance_emergency",
             "state_machine_caution_speed_multiplier", "state_machine_caution_accel_multiplier",
             "state_machine_emergency_accel_multiplier", "state_machine_emergency_lat_accel_multiplier",
             "state_machine_envelope_decel", "state_machine_envelope_standoff", "ego_emergency_decel",
             "chance_epsilon", "collision_margin_inflation", "vehicle_length", "vehicle_width", "ego_footprint_n_circles")


def _install_stub(sim, sgan):
    """Seeded stand-in generator in the unmodified predictor: predictor.predict() then runs its SGAN branch
    (trajectory_predictor.py:164-186: generator -> relative_to_abs -> process_prediction) and predict_single_best its
    multi-sample branch (:338-353)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests.stub_sampler import StubGenerator
    sim.predictor.method = "sgan"
    sim.predictor.generator = StubGenerator(sgan["gen_seed"], sim.predictor.pred_len, sgan.get("sigma", 0.05))
    assert sim.predictor.num_samples == sgan["num_samples"]
    assert sim.distribution_aware_planning == bool(sgan["distribution_aware"])


def run_variant(seed, scenario="scenario_01_cv", footprint=False, sgan=None):
    import yaml
    from src.config import SimulationConfig, validate_config
    from src.core.data_structures import VehicleState
    from src.simulation.integrated_simulator import IntegratedSimulator
    from src.simulation.replay_source import ReplayPedestrianSource

    with open(os.path.join(REF, "scenarios", scenario + ".yaml")) as f:
        d = yaml.safe_load(f)
    peds = np.array(d.pop("ped_initial_states"), dtype=float)
    d.pop("ped_groups", None)
    d["sgan_model_path"] = None
    d["visualization_enabled"] = False
    d["prediction_method"] = "cv"
    if footprint:
        d["ego_footprint"] = "multi_circle"
    if sgan:
        d["num_samples"] = sgan["num_samples"]
        d["distribution_aware_planning"] = bool(sgan["distribution_aware"])
        d["chance_epsilon"] = sgan["chance_epsilon"]
    if seed > 0:
        rng = np.random.default_rng(seed)
        peds[:, 0:2] += rng.normal(0, 1.5, peds[:, 0:2].shape)
        peds[:, 2:4] *= rng.uniform(0.6, 1.4, (len(peds), 1))
        e0 = list(d["ego_initial_state"])
        if scenario == "scenario_01_cv":         # (kept as first recorded: rollout_s01.npz)
            d["ego_initial_state"] = [0.0, float(rng.uniform(-0.3, 0.3)), 0.0, float(rng.uniform(2.0, 6.0)), 0.0]
        else:
            d["ego_initial_state"] = [e0[0], e0[1] + float(rng.uniform(-0.3, 0.3)), e0[2], float(rng.uniform(2.0, 6.0)), 0.0]
    cfg = SimulationConfig(**d)
    validate_config(cfg)
    sim = IntegratedSimulator(cfg)
    assert (sim.ego_footprint is not None) == bool(footprint)
    if sgan:
        _install_stub(sim, sgan)
    n_frames = int(cfg.total_time / cfg.dt) + 200
    t = np.arange(n_frames)[:, None, None] * cfg.dt
    traj = peds[None, :, 0:2] + peds[None, :, 2:4] * t
    sim.pedestrian_sim = ReplayPedestrianSource(traj, dt=cfg.dt)

    n_calls = [0]
    original = sim.planner.plan

    def counting_plan(*a, **k):
        n_calls[0] += 1
        return original(*a, **k)

    sim.planner.plan = counting_plan
    sim.warmup()
    order = {VehicleState.NORMAL: 0, VehicleState.CAUTION: 1, VehicleState.EMERGENCY: 2}
    ego, fsm, found, calls = [], [], [], []
    n_steps = int(cfg.total_time / cfg.dt)
    reason = "timeout"
    for i in range(n_steps):
        n_calls[0] = 0
        res = sim.step()
        e = sim.ego_state
        ego.append([e.x, e.y, e.yaw, e.v, e.a])
        fsm.append(order[sim.state_machine.current_state])
        found.append(res.planned_path is not None)
        calls.append(n_calls[0])
        if res.metrics.get("collision", False):
            reason = "collision"
            break
        s_now = sim.coord_converter.find_nearest_point_on_path(e.x, e.y)[0]
        if sim.reference_path.s[-1] - s_now < 2.0:
            reason = "goal"
            break
    knobs = {k: getattr(cfg, k, None) for k in KNOB_KEYS}
    knobs["ego_footprint_multi_circle"] = 1.0 if footprint else 0.0
    return dict(traj=traj, ego0=np.array(cfg.ego_initial_state, dtype=float), ego=np.array(ego), fsm=np.array(fsm),
                found=np.array(found), calls=np.array(calls), reason=reason, knobs=knobs,
                static_points=np.asarray(sim.static_obstacle_points, dtype=float).reshape(-1, 2),
                wx=np.array(cfg.reference_waypoints_x, dtype=float), wy=np.array(cfg.reference_waypoints_y, dtype=float))


def record_trajectory_file(n_steps=80, n_pred=5, sgan=None, out_name="rollout_s01_trajectory.npz"):
    """The reference's own result file: scenario_01_cv for `n_steps` steps, IntegratedSimulator.save_results(), and
    the first `n_steps` entries of every array of its trajectory.npz (predictions: first `n_pred` steps only) ->
    rollout_s01_trajectory.npz.  Ragged per-step arrays are stored padded with NaN plus their lengths."""
    import tempfile
    import yaml
    from src.config import SimulationConfig, validate_config
    from src.simulation.integrated_simulator import IntegratedSimulator
    from src.simulation.replay_source import ReplayPedestrianSource
    with open(os.path.join(REF, "scenarios", "scenario_01_cv.yaml")) as f:
        d = yaml.safe_load(f)
    peds = np.array(d.pop("ped_initial_states"), dtype=float)
    d.pop("ped_groups", None)
    d["sgan_model_path"] = None
    d["visualization_enabled"] = False
    d["prediction_method"] = "cv"
    if sgan:
        d["num_samples"] = sgan["num_samples"]
        d["distribution_aware_planning"] = bool(sgan["distribution_aware"])
        d["chance_epsilon"] = sgan["chance_epsilon"]
    cfg = SimulationConfig(**d)
    validate_config(cfg)
    sim = IntegratedSimulator(cfg)
    if sgan:
        _install_stub(sim, sgan)
    n_frames = int(cfg.total_time / cfg.dt) + 200
    t = np.arange(n_frames)[:, None, None] * cfg.dt
    sim.pedestrian_sim = ReplayPedestrianSource(peds[None, :, 0:2] + peds[None, :, 2:4] * t, dt=cfg.dt)
    sim.warmup()
    for _ in range(n_steps):
        sim.step()
    sim.termination_reason = "timeout"
    out = tempfile.mkdtemp()
    sim.save_results(out)
    z = np.load(os.path.join(out, "trajectory.npz"), allow_pickle=True)
    store = {"keys": np.array(sorted(z.files))}
    # the reference's own metrics files (integrated_simulator.py:1019-1065), verbatim
    store["metrics_csv"] = np.array(open(os.path.join(out, "metrics_summary.csv")).read())
    store["metrics_txt"] = np.array(open(os.path.join(out, "metrics_report.txt")).read())
    if sgan is not None:
        for key, val in sgan.items():
            store["sgan/" + key] = np.array(val, dtype=float)
    for key in z.files:
        a = z[key]
        if a.dtype == object:
            rows = [np.asarray(r, dtype=float) for r in a]
            if key == "predicted_trajectories":
                rows = rows[:n_pred]
            store[key + "/shape"] = np.array([r.shape + (0,) * (3 - r.ndim) for r in rows])
            flat = [r.reshape(-1) for r in rows]
            pad = np.full((len(flat), max(len(r) for r in flat)), np.nan)
            for i, r in enumerate(flat):
                pad[i, :len(r)] = r
            store[key] = pad
        elif a.dtype.kind in "US":
            store[key] = a.astype("U16")
        else:
            store[key] = a
    np.savez_compressed(os.path.join(HERE, out_name), **store)
    print("wrote", out_name, os.path.getsize(os.path.join(HERE, out_name)) // 1024, "KiB",
          {k: (z[k].shape, str(z[k].dtype)) for k in z.files})


def record(scenario, out_name, seeds, footprint=False, sgan=None):
    store = {}
    for k, seed in enumerate(seeds):
        r = run_variant(seed, scenario, footprint, None if sgan is None else dict(sgan, gen_seed=sgan["gen_seed"] + 1000 * k))
        for name in ("traj", "ego0", "ego", "fsm", "found", "calls", "wx", "wy"):
            store[f"v{k}/{name}"] = r[name]
        store[f"v{k}/reason"] = np.array(r["reason"])
        if k == 0:
            for key, val in r["knobs"].items():
                store["knob/" + key] = np.array(np.nan if val is None else val, dtype=float)
            store["static_points"] = r["static_points"]
        states = np.bincount(r["fsm"], minlength=3)
        print(f"{scenario} variant {k}: {len(r['ego'])} steps, {r['reason']}, plan calls {int(r['calls'].sum())}, "
              f"states N/C/E {states.tolist()}, failed steps {int((~r['found']).sum())}, static points {len(r['static_points'])}")
    store["n_variants"] = np.array(len(seeds))
    if sgan is not None:
        for key, val in sgan.items():
            store["sgan/" + key] = np.array(val, dtype=float)
    np.savez_compressed(os.path.join(HERE, out_name), **store)
    print("wrote", out_name, os.path.getsize(os.path.join(HERE, out_name)) // 1024, "KiB")


SGAN_DIST = dict(num_samples=6, distribution_aware=1, chance_epsilon=0.2, gen_seed=4242, sigma=0.05)
SGAN_BEST = dict(num_samples=4, distribution_aware=0, chance_epsilon=0.0, gen_seed=777, sigma=0.05)


def main():
    """No argument: scenario_01_cv and three perturbed variants -> rollout_s01.npz (as first recorded).
    `s02` / `s03`: scenario_02_cv (corridor between two static walls) / scenario_03_cv (right turn) and one
    perturbed variant each -> rollout_s02.npz / rollout_s03.npz."""
    which = sys.argv[1:] or ["s01"]
    if "s02" in which:
        record("scenario_02_cv", "rollout_s02.npz", (0, 1))
    if "trajfile" in which:
        record_trajectory_file()
    if "sgan" in which:         # distribution-aware planning on sample sets of a seeded stub generator (6 samples, epsilon 0.2)
        record("scenario_01_cv", "rollout_s01_dist.npz", (0, 2), sgan=SGAN_DIST)
    if "sgan_single" in which:  # 4 samples, the representative sample only (distribution_aware_planning off)
        record("scenario_01_cv", "rollout_s01_best.npz", (0,), sgan=SGAN_BEST)
    if "sgan_trajfile" in which:
        record_trajectory_file(sgan=dict(SGAN_DIST), out_name="rollout_s01_dist_trajectory.npz")
    if "s02fp" in which:          # the same corridor with the three-circle footprint of the 4.5 m x 2.0 m vehicle
        record("scenario_02_cv", "rollout_s02fp.npz", (0, 2), footprint=True)
    if "s03" in which:
        record("scenario_03_cv", "rollout_s03.npz", (0, 1))
    if "s01" in which:
        main_s01()


def main_s01():
    store = {}
    for k, seed in enumerate((0, 1, 2, 3)):
        r = run_variant(seed)
        for name in ("traj", "ego0", "ego", "fsm", "found", "calls", "wx", "wy"):
            store[f"v{k}/{name}"] = r[name]
        store[f"v{k}/reason"] = np.array(r["reason"])
        if k == 0:
            for key, val in r["knobs"].items():
                store["knob/" + key] = np.array(np.nan if val is None else val, dtype=float)
        states = np.bincount(r["fsm"], minlength=3)
        print(f"variant {k}: {len(r['ego'])} steps, {r['reason']}, plan calls {int(r['calls'].sum())}, "
              f"states N/C/E {states.tolist()}, failed steps {int((~r['found']).sum())}")
    store["n_variants"] = np.array(4)
    np.savez_compressed(os.path.join(HERE, "rollout_s01.npz"), **store)
    print("wrote rollout_s01.npz", os.path.getsize(os.path.join(HERE, "rollout_s01.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
