#!/usr/bin/env python
"""bench.py -- Frenet candidate evals/s of the batched sweep (BASELINE.json metric).

One "step" = one pass of the hot path over one batch: Q independent planning queries per GPU
(scenario_01 grid, 1261 candidates x 50 pedestrians x 1 sample,
51 obstacle steps), queries sharded over ranks with no data-path collective (weak scaling: Q per
GPU is fixed) and one small all_gather of the winners' (index, cost, stats) per step when N > 1.

  value      dense evaluations/s with every input already resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-pointer C-ABI call `fot_plan_batch_host` (what
             FrenetPlanner.plan() binds): pinned host inputs -> H2D -> kernels -> D2H of the winners
  roofline   the sweep kernel: algorithmic 5 FLOP x evaluations / its own CUDA-event time, against the
             FP64 FMA peak measured on this GPU by fot_probe_fma_tflops (MEASURED_PEAKS.json holds no
             FP64 figure)
  cpu_baseline  the NumPy oracle (a port of the reference planner) on a bounded sample of the same
             queries on this box's host cores (rank 0, N = 1 only)

`--impl reference` times only that CPU leg (all host cores) and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import scenarios  # noqa: E402  (seeded synthetic fields; no reference access)

METRIC = "frenet_candidate_evals_per_s"
UNIT = "evals/s"
FLOP_PER_EVAL = 5.0      # 2 SUB + 1 MUL + 1 FMA
N_PEDS, T_OBS = 50, 51
TARGET_SPEED = 6.0
NCU_DRAM_READ, NCU_DRAM_WRITE = 187.824896e6, 4.817152e6   # bytes per 4096-query launch (ncu --set full)


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def make_queries(first: int, count: int):
    """Queries first..first+count-1 of config 4 (seed = global query id): ego state -> Frenet
    state on the host exactly as plan() does it, plus that query's pedestrian field."""
    from integrated_path_planning_b200 import CubicSpline2D
    from integrated_path_planning_b200.frenet_host import CoordinateConverter, ego_to_frenet
    from integrated_path_planning_b200.types import EgoVehicleState
    spline = CubicSpline2D(*scenarios.STRAIGHT_60)
    frenet = np.empty((count, 6))
    dyn = np.empty((count, 1, N_PEDS, T_OBS, 2))
    for i in range(count):
        rng = np.random.default_rng(first + i)
        dyn[i, 0] = scenarios.pedestrian_field(rng, N_PEDS, T_OBS, scenarios.S1_KNOBS["dt"])
        ego = EgoVehicleState(x=rng.uniform(2, 20), y=rng.uniform(-1, 1), yaw=rng.normal(0, 0.05),
                              v=rng.uniform(0, 8), a=rng.uniform(-1, 1))
        fs = ego_to_frenet(CoordinateConverter(spline), ego, 0.0)
        frenet[i] = fs
    return spline, frenet, dyn


def _oracle_worker(args):
    frenet, dyn = args
    from oracle import frenet_oracle as O
    pl = O.OraclePlanner(O.Spline2D(*scenarios.STRAIGHT_60), O.Knobs(**scenarios.S1_KNOBS))
    res = pl.plan_frenet(tuple(frenet), np.empty((0, 2)), dyn, TARGET_SPEED)
    return int(res.n_points.sum()) * N_PEDS, res.best_index


def cpu_leg(frenet, dyn, n_sample, cores, steps, warmup):
    """Oracle (reference port) on `n_sample` queries per step over a process pool."""
    import multiprocessing as mp
    jobs = [(frenet[i], dyn[i, 0]) for i in range(n_sample)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_oracle_worker, jobs[:cores])
        times, evals = [], 0
        for _ in range(steps):
            t0 = time.perf_counter()
            out = pool.map(_oracle_worker, jobs)
            times.append(time.perf_counter() - t0)
            evals = sum(o[0] for o in out)
    ms = 1e3 * float(np.mean(times))
    return evals / (ms * 1e-3), ms, evals


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
def oracle_plan_latency(calls=3):
    """config 2 plan() on one host core with the NumPy oracle (the reference port): p50 in ms."""
    from oracle import frenet_oracle as O
    opl = O.OraclePlanner(O.Spline2D(*scenarios.STRAIGHT_60), O.Knobs(**scenarios.S1_KNOBS))
    dyn2 = scenarios.pedestrian_field(np.random.default_rng(1), N_PEDS, T_OBS, scenarios.S1_KNOBS["dt"])
    lat = []
    for _ in range(calls + 1):
        opl.reset_ego_curvature()
        t0 = time.perf_counter()
        opl.plan((5.0, 0.0, 0.0, 5.0, 0.0), np.empty((0, 2)), dyn2, TARGET_SPEED)
        lat.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(lat[1:]))


class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=4096, help="planning queries per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU sample (0 = 2 x cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0))
    config = {"workload": "config4: 4096 independent plan() queries per GPU, scenario_01 grid "
                          "(19 d x 11 T x 6 v + 7 brake = 1261 candidates, 58041 points), "
                          "50 pedestrians x 1 sample x 51 steps, seed = query id",
              "queries_per_gpu": args.queries, "parallelism": f"query-sharded x{world}",
              "l2": "inputs (167 MB of obstacle tracks per step) exceed the 126 MB L2; no explicit flush"}

    # ---------------- reference arm: CPU only --------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        n_sample = args.cpu_sample or 2 * cores
        _, frenet, dyn = make_queries(0, n_sample)
        val, ms, evals = cpu_leg(frenet, dyn, n_sample, cores, max(1, min(args.steps, 5)), 1)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n_sample} of the {args.queries} queries per step "
                                           f"({evals:.3g} dense evals), NumPy oracle over a {cores}-process pool; "
                                           f"timed steps capped at 5",
                                 "plan_p50_ms": oracle_plan_latency(),
                                 "plan_sample": "config2 plan() on one core, 3 calls after 1 warm-up"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- B200 arm -----------------------------------------------------------
    import torch
    import torch.distributed as dist
    from integrated_path_planning_b200 import BatchFrenetPlanner, DeviceBatch, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sweep has no CPU path")
    # stdout carries the one JSON line only: anything a library prints meanwhile (NCCL's version banner)
    # goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    Q = args.queries
    spline, frenet, dyn = make_queries(rank * Q, Q)
    planner = BatchFrenetPlanner(spline, device=local_rank, **scenarios.S1_KNOBS)
    eng = planner.engine

    # resident leg: inputs in HBM, torch tensors as buffers
    batch = DeviceBatch(planner, frenet, TARGET_SPEED, dyn, _lib.FOT_DYN_SINGLE)
    evals_step = batch.dense_evals()
    stream = torch.cuda.Stream(device=local_rank)
    # The one collective of the path: every step's winners (cost per query) are all-gathered.  The
    # gather runs on its own stream behind the sweep that produced them, four staging buffers deep, so the next
    # step's sweep does not wait for the slowest rank of this one; the timed region ends after the
    # last gather has completed on every rank.  The gather stream has the higher priority: the collective's few
    # CTAs then take the first SM slots the running sweep frees instead of queueing behind its whole grid.
    N_GBUF = 4
    gstream = torch.cuda.Stream(device=local_rank, priority=-1) if world > 1 else None
    stage_buf = [torch.empty_like(batch.out["best_cost"]) for _ in range(N_GBUF)] if world > 1 else None
    gathered = [torch.empty((world,) + tuple(batch.out["best_cost"].shape), dtype=torch.float64, device="cuda")
                for _ in range(N_GBUF)] if world > 1 else None
    g_done = [None] * N_GBUF
    step_no = [0]

    def resident_step():
        batch.launch(stream.cuda_stream)
        if world > 1:
            k = step_no[0] % N_GBUF
            step_no[0] += 1
            with torch.cuda.stream(stream):
                if g_done[k] is not None:
                    stream.wait_event(g_done[k])            # the gather that last read this staging buffer
                stage_buf[k].copy_(batch.out["best_cost"], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(stream)
            with torch.cuda.stream(gstream):
                gstream.wait_event(ready)
                dist.all_gather_into_tensor(gathered[k], stage_buf[k])
                g_done[k] = torch.cuda.Event()
                g_done[k].record(gstream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        resident_step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        resident_step()
    with torch.cuda.stream(stream):
        if gstream is not None:
            stream.wait_stream(gstream)                     # the timed region ends after the last gather
        e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    n_timed = min(args.steps, 256)
    stage = np.array([eng.launch_stage_ms(b) for b in range(n_timed)])   # [prepass, sweep, winner]
    sweep_ms = float(stage[:, 1].mean())
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = evals_step * world / (ms_step * 1e-3)
    # per-rank kernel time of a step (prepass + sweep + winner, CUDA events of each rank's own launches): with no
    # data-path collective the slowest GPU sets the pace, and this is where a scaling loss shows
    per_rank_ms = [float(stage.sum(axis=1).mean())]
    if world > 1:
        tk = torch.tensor(per_rank_ms, dtype=torch.float64, device="cuda")
        allk = torch.empty(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allk, tk)
        per_rank_ms = [float(v) for v in allk.cpu()]
    per_rank_mhz = [float(clocks.get("sm_mhz") or 0.0)]
    if world > 1:
        tk = torch.tensor(per_rank_mhz, dtype=torch.float64, device="cuda")
        allk = torch.empty(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allk, tk)
        per_rank_mhz = [float(v) for v in allk.cpu()]

    # e2e leg: host-pointer C-ABI call, pinned inputs, H2D + kernels + D2H inside the timed region
    dyn_pinned = torch.from_numpy(dyn).pin_memory()
    dyn_host = dyn_pinned.numpy()
    for _ in range(args.warmup):
        res = planner.plan_batch(frenet, TARGET_SPEED, dynamic_obstacles=dyn_host[:, 0])
    e2e_steps = max(3, args.steps // 3)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = planner.plan_batch(frenet, TARGET_SPEED, dynamic_obstacles=dyn_host[:, 0])
    if world > 1:
        dist.barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = dyn.nbytes + frenet.nbytes + Q * (8 + 32 + 8 + 4) + Q * 6 * 8
    d2h = res.best_idx.nbytes + res.best_cost.nbytes + res.stats.nbytes + res.winner_len.nbytes + res.winner.nbytes
    # resident and e2e legs must pick the same winners
    same = bool(np.array_equal(batch.out["best_idx"].cpu().numpy(), res.best_idx))

    # e2e with the predictor's post-processing on the device: the caller hands
    # over the last two pedestrian observations and the current positions (host, pinned) instead of the
    # obstacle tensor; constant-velocity extrapolation + t = 0 column are built on the GPU and feed the sweep.
    from integrated_path_planning_b200.prediction import DevicePredictionPostprocessor
    post = DevicePredictionPostprocessor(pred_len=12, sgan_dt=0.4, sim_dt=scenarios.S1_KNOBS["dt"], plan_horizon=5.0,
                                         device=local_rank)
    p0 = dyn[:, 0, :, 0, :]
    vel = (dyn[:, 0, :, 1, :] - dyn[:, 0, :, 0, :]) / scenarios.S1_KNOBS["dt"]
    host_in = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (p0, p0 - vel * 0.4, p0)]
    dev_in = [torch.empty_like(a, device="cuda") for a in host_in]
    dyn_buf = torch.empty((Q, 1, N_PEDS, post.n_steps + 1, 2), dtype=torch.float64, device="cuda")
    batch_cv = DeviceBatch(planner, frenet, TARGET_SPEED, dyn_buf, _lib.FOT_DYN_SINGLE)
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in batch_cv.out.items()}

    def cv_step():
        with torch.cuda.stream(stream):
            for d, h_ in zip(dev_in, host_in):
                d.copy_(h_, non_blocking=True)
            post.predict_cv(dev_in[0], dev_in[1], None, dev_in[2], out=dyn_buf)
        batch_cv.launch_to_host(host_out, stream.cuda_stream)    # ranges of queries; winners go back while the rest is swept

    for _ in range(args.warmup):
        cv_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cv_step()
    if world > 1:
        dist.barrier()
    cv_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([cv_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cv_ms = float(t.item())
    cv_match = float(np.mean(host_out["best_idx"].numpy() == res.best_idx))
    e2e_cv = {"value": evals_step * world / (cv_ms * 1e-3), "unit": UNIT, "ms_per_step": cv_ms,
              "h2d_bytes_per_step": int(sum(a.numel() * 8 for a in host_in)),
              "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host_out.values())),
              "winners_equal_to_tensor_path": cv_match,
              "note": "inputs = last two observations + current positions per query (host, pinned); constant-velocity "
                      "obstacle tensor built on the device (fot_predict_cv_device), then the same sweep"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # FP64 pipe peak measured here (2 FLOP per FMA)
    import ctypes as C
    peak = C.c_double()
    _lib.check(eng.lib.fot_probe_fma_tflops(local_rank, 0, C.byref(peak)), "probe")
    achieved_tf = FLOP_PER_EVAL * evals_step / (sweep_ms * 1e-3) / 1e12
    # DRAM traffic of one fot_sweep_items launch from the committed ncu --set full capture
    # (profiles/r1/sweep_items_ncu_full_summary.txt: read + written bytes for 4096 queries), per query
    traffic = (NCU_DRAM_READ + NCU_DRAM_WRITE) / 4096 * Q
    roofline = {"bound": "fp64_pipe", "achieved": achieved_tf, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved_tf / peak.value, "traffic": traffic, "traffic_unit": "bytes per launch (ncu, profiles/r1)",
                "kernel": "fot_sweep_items", "kernel_ms": sweep_ms,
                "stage_ms": {"prepass": float(stage[:, 0].mean()), "sweep": sweep_ms, "winner": float(stage[:, 2].mean())},
                "peak_source": "fot_probe_fma_tflops on this GPU (dependent-chain DFMA, 2 FLOP/FMA); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "note": "algorithmic 5 FLOP per dense evaluation; point generation (about 58041 points per query) "
                        "is extra work not credited here"}

    # plan() latency (BASELINE config 2: one call, default grid, 50 pedestrians x 1 sample), measured the
    # way the reference's simulator measures it: time.perf_counter around FrenetPlanner.plan()
    # (integrated_simulator.py:575-585) -- ego->Frenet on the host, H2D, kernels, D2H, FrenetPath rebuilt
    from integrated_path_planning_b200 import FrenetPlanner
    from integrated_path_planning_b200.types import EgoVehicleState
    single = FrenetPlanner(spline, device=local_rank, **scenarios.S1_KNOBS)
    ego2 = EgoVehicleState(x=5.0, y=0.0, yaw=0.0, v=5.0, a=0.0)
    dyn2 = scenarios.pedestrian_field(np.random.default_rng(1), N_PEDS, T_OBS, scenarios.S1_KNOBS["dt"])
    lat, kms = [], []
    for i in range(103):
        single.reset_ego_curvature()
        t0 = time.perf_counter()
        single.plan(ego2, np.empty((0, 2)), dyn2, TARGET_SPEED)
        lat.append(1e3 * (time.perf_counter() - t0))
        kms.append(single.last_result.kernel_ms)
    plan_latency = {"p50_ms": float(np.median(lat[3:])), "p95_ms": float(np.percentile(lat[3:], 95)),
                    "kernels_p50_ms": float(np.median(kms[3:])), "calls": 100,
                    "config": "config2: one plan() call, 1261 candidates x 50 pedestrians x 1 sample"}

    # closed-loop campaign: 256 simulations of the recorded scenario_01 variants (with
    # jitter) advanced in lock-step by the batched driver; informational, never allowed to break the bench line
    closed_loop = None
    try:
        golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "rollout_s01.npz")
        if world == 1 and os.path.exists(golden):
            from integrated_path_planning_b200.rollout import BatchedClosedLoop
            z = np.load(golden)
            knobs = {k[5:]: float(z[k]) for k in z.files if k.startswith("knob/")}
            rng_cl = np.random.default_rng(7)
            n_sims, n_var = 256, int(z["n_variants"])
            tracks = np.stack([z[f"v{i % n_var}/traj"] + rng_cl.normal(0.0, 0.3, (1, z["v0/traj"].shape[1], 2)) for i in range(n_sims)])
            ego0 = np.stack([z[f"v{i % n_var}/ego0"] for i in range(n_sims)])
            sim = BatchedClosedLoop(z["v0/wx"], z["v0/wy"], knobs, tracks, ego0, device=local_rank)
            sim.warmup()
            for _ in range(3):
                sim.step()
            calls0, t0, sim_steps = sim.n_plan_calls, time.perf_counter(), 0
            for _ in range(30):
                sim_steps += int(sim.active.sum())
                sim.step()
            wall = time.perf_counter() - t0
            closed_loop = {"sim_steps_per_s": sim_steps / wall, "plan_calls_per_s": (sim.n_plan_calls - calls0) / wall,
                           "sims": n_sims, "lockstep_ms": 1e3 * wall / 30,
                           "note": "BatchedClosedLoop: observer, CV prediction, safety metrics, fail-safe state machine, sweep "
                                   "(+ escalation retries), ego update per step; the reference runs ~4.8 such steps/s"}
    except Exception as exc:                     # pragma: no cover
        closed_loop = {"error": repr(exc)}

    cpu = None
    if world == 1 and not args.no_cpu:
        n_sample = args.cpu_sample or 2 * cores
        val, ms, ev = cpu_leg(frenet, dyn, n_sample, cores, 2, 1)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_sample} of the {Q} queries ({ev:.3g} dense evals), NumPy oracle over a "
                         f"{cores}-process pool, mean of 2 passes, {ms:.0f} ms per pass",
               "plan_p50_ms": oracle_plan_latency(),
               "plan_sample": "config2 plan() on one core, 3 calls after 1 warm-up"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": evals_step * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "winners_match_resident": same},
            "gpu_launches": 4 * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "per_rank_kernel_ms_per_step": per_rank_ms, "per_rank_sm_mhz": per_rank_mhz,
            "plan_latency": plan_latency, "e2e_device_prediction": e2e_cv, "closed_loop": closed_loop,
            "candidates_per_s": float(res.n_cand.sum()) * world / (ms_step * 1e-3),
            "evals_per_step_per_gpu": evals_step}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
