#!/usr/bin/env python
"""Record the differential campaign by running the UNMODIFIED reference planner (build container only).

    python tests/golden/make_golden_campaign.py [a] [b] [c] [d0] [d1]      (no argument: everything)

Inputs come from tests/campaign.py (seeded; the GPU tests rebuild them).  Stored per plan() call: the reference's
Frenet state, chosen index, cost, last_check_stats, number of candidates and a CRC-32 of the per-candidate category
vector (the vector itself for config 3) -> tests/golden/campaign_{a,b,c}.npz, config3_v{0,1}.npz.
The planner is driven through the reference's own plan() steps (frenet_planner.py:261-304) exactly as
tests/golden/make_golden.py does, so that every candidate's category is known.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as G  # noqa: E402  (imports the reference read-only, defines reference_plan)
from src.core.data_structures import FrenetState  # noqa: E402

from tests import campaign, scenarios  # noqa: E402

STATS = G.CAT_NAMES


def _pack(res):
    return dict(fs=res["fs"], best=np.int32(res["best"]), cost=np.float64(res.get("cost", np.inf)),
                stats=res["stats"].astype(np.int32), n_cand=np.int32(len(res["cats"])), crc=np.uint32(campaign.crc(res["cats"])))


def reference_plan_frenet(planner, fs, static, dyn, target, overrides, dist, msd):
    """G.reference_plan with the Frenet state given (skips _cartesian_to_frenet_state only)."""
    stub = planner._cartesian_to_frenet_state
    planner._cartesian_to_frenet_state = lambda ego: FrenetState(*[float(v) for v in fs])
    try:
        return G.reference_plan(planner, None, static, dyn, target, overrides, dist, msd)
    finally:
        planner._cartesian_to_frenet_state = stub


_PL = {}


def _planner(key, waypoints, knobs):
    if key not in _PL:
        _PL[key] = G.FrenetPlanner(G.CubicSpline2D(*waypoints), **knobs)
    return _PL[key]


def _work_a(i):
    ego, dyn = campaign.query_a(i)
    pl = _planner("a", scenarios.STRAIGHT_60, scenarios.S1_KNOBS)
    pl._last_kappa = 0.0
    if hasattr(pl.converter, "_prev_s"):
        del pl.converter._prev_s
    return _pack(G.reference_plan(pl, G.EgoVehicleState(*ego), np.empty((0, 2)), dyn, campaign.TARGET_SPEED, None, None, None))


def _work_b(i):
    q = campaign.query_b(i)
    pl = _planner(("b", q["path"], q["road"]), campaign.PATHS[q["path"]], campaign.knobs_b(q["road"]))
    return _pack(reference_plan_frenet(pl, q["fs"], np.empty((0, 2)), q["dyn"], q["target"], q["overrides"], None, q["msd"]))


def _stack(rows):
    return {k: np.stack([r[k] for r in rows]) for k in rows[0]}


def make_pool(name, work, n, procs):
    t0 = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        rows = pool.map(work, range(n), chunksize=8)
    store = _stack(rows)
    np.savez_compressed(os.path.join(HERE, f"campaign_{name}.npz"), **store)
    found = int((store["best"] >= 0).sum())
    print(f"campaign {name}: {n} queries in {time.time() - t0:.0f} s, {found} with a path, category totals "
          f"{dict(zip(STATS, store['stats'].clip(0).sum(axis=0).tolist()))}, "
          f"{os.path.getsize(os.path.join(HERE, f'campaign_{name}.npz')) // 1024} KiB")


def make_c():
    t0 = time.time()
    pl = G.FrenetPlanner(G.CubicSpline2D(*campaign.C_PATH), **scenarios.S1_KNOBS)
    rows = []
    for step, ego, dyn, static in campaign.rollout_c():
        for target, ovr, msd in campaign.plans_c():
            st = np.empty((0, 2)) if static is None else static
            kappa_in = pl._last_kappa
            res = G.reference_plan(pl, G.EgoVehicleState(*ego), st, dyn, target, ovr, None, msd)
            row = _pack(res)
            row["kappa_in"] = np.float64(kappa_in)
            rows.append(row)
    store = _stack(rows)
    np.savez_compressed(os.path.join(HERE, "campaign_c.npz"), **store)
    print(f"campaign c: {len(rows)} plan() calls in {time.time() - t0:.0f} s, {int((store['best'] >= 0).sum())} with a path, "
          f"candidate counts {sorted(set(store['n_cand'].tolist()))}, "
          f"{os.path.getsize(os.path.join(HERE, 'campaign_c.npz')) // 1024} KiB")


def make_d(variant):
    t0 = time.time()
    c = campaign.config3(variant)
    pl = G.FrenetPlanner(G.CubicSpline2D(*c["waypoints"]), **c["knobs"])
    res = reference_plan_frenet(pl, c["fs"], np.empty((0, 2)), None, c["target"], None, c["dist"], None)
    store = dict(cats=res["cats"], costs=res["costs"], best=res["best"], stats=res["stats"], fs=res["fs"])
    if int(res["best"]) >= 0:
        store["cost"] = res["cost"]
        for name in G.SERIES:
            store["w_" + name] = res["w_" + name]
    out = os.path.join(HERE, f"config3_v{variant}.npz")
    np.savez_compressed(out, **store)
    print(f"config 3 variant {variant}: {len(res['cats'])} candidates, one reference plan() = {time.time() - t0:.0f} s, best "
          f"{int(res['best'])}, stats {dict(zip(STATS, res['stats'].tolist()))}, {os.path.getsize(out) // 1024} KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["a", "b", "c", "d0", "d1"]
    procs = int(os.environ.get("PROCS", str(len(os.sched_getaffinity(0)))))
    if "d0" in which:
        make_d(0)
    if "d1" in which:
        make_d(1)
    if "c" in which:
        make_c()
    if "a" in which:
        make_pool("a", _work_a, campaign.N_A, procs)
    if "b" in which:
        make_pool("b", _work_b, campaign.N_B, procs)
