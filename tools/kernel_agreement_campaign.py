"""Differential campaign between the sweep kernels, far beyond what was recorded from the reference: N_QUERIES fresh
queries of the "limits through the candidates' range" kind (tests/campaign.py, kind B: straight / S-curve / arc, two road
widths, random limits, targets and stop distances) with a denser pedestrian field, every candidate's category and cost and
every winner compared bit for bit between fot_sweep_pairs (the default) and fot_sweep_items.  fot_sweep_items is the
kernel whose answers were pinned against the unmodified reference at the start of the round; the pair kernel is also held
to the recorded reference answers directly (tests/test_gpu_campaign.py).  Writes one JSON object to stdout.  GPU only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import campaign, runners, scenarios
from integrated_path_planning_b200 import BatchFrenetPlanner, CubicSpline2D

N_QUERIES = int(os.environ.get("N_QUERIES", 98304))
BATCH = 1024
SEED0 = 50_000_000
t0 = time.time()
planners = {}
out = dict(queries=0, candidates=0, category_mismatches=0, cost_bit_mismatches=0, winner_mismatches=0, stats_mismatches=0,
           with_a_path=0, categories={}, kinds_seen=set())
cat_hist = np.zeros(16, dtype=np.int64)
for b0 in range(0, N_QUERIES, BATCH):
    combo = (b0 // BATCH) % 6
    path, road = campaign.PATH_NAMES[combo % 3], (2.7, 1.1)[combo // 3]
    key = (path, road)
    if key not in planners:
        planners[key] = BatchFrenetPlanner(CubicSpline2D(*campaign.PATHS[path]), **campaign.knobs_b(road))
    pl = planners[key]
    rng = np.random.default_rng(SEED0 + b0)
    n = min(BATCH, N_QUERIES - b0)
    fs = np.stack([rng.uniform(3, 25, n), rng.uniform(0.0, 9.0, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-2.9, 2.9, n),
                   rng.uniform(-1.0, 1.0, n), rng.uniform(-0.5, 0.5, n)], axis=1)
    lim = np.stack([rng.uniform(2.0, 9.0, n), rng.uniform(0.3, 3.0, n), rng.uniform(0.01, 0.3, n), rng.uniform(0.05, 2.0, n)], axis=1)
    target = rng.uniform(0.5, 9.0, n)
    stop = np.full(n, np.nan)
    emer = np.arange(n) % 7 == 0
    target[emer] = 0.0
    stop[emer] = rng.uniform(2.0, 12.0, int(emer.sum()))
    dyn = np.stack([scenarios.pedestrian_field(rng, 24, x_range=(0.0, 60.0), y_range=(-7.0, 7.0)) for _ in range(n)])
    run = lambda: pl.plan_batch(fs, target, dynamic_obstacles=dyn, limits=lim, max_stop_distance=stop, want_candidates=True)
    pl.engine
    res = {}
    for kern in ("pairs", "items"):
        with runners.fot_env(FOT_SWEEP=kern):
            res[kern] = run()
            out["kinds_seen"].add(int(pl.engine.lib.fot_last_sweep_kind(pl.engine._h)))
    a, b = res["pairs"], res["items"]
    assert np.array_equal(a.n_cand, b.n_cand)
    live = np.arange(a.cand_cat.shape[1])[None, :] < a.n_cand[:, None]
    out["queries"] += n
    out["candidates"] += int(live.sum())
    out["category_mismatches"] += int(((a.cand_cat != b.cand_cat) & live).sum())
    ca, cb = a.cand_cost[:, :live.shape[1]].view(np.uint64), b.cand_cost[:, :live.shape[1]].view(np.uint64)
    out["cost_bit_mismatches"] += int(((ca != cb) & live).sum())
    out["winner_mismatches"] += int((a.best_idx != b.best_idx).sum() + (a.best_cost.view(np.uint64) != b.best_cost.view(np.uint64)).sum())
    out["stats_mismatches"] += int((a.stats != b.stats).any(axis=1).sum())
    out["with_a_path"] += int((a.best_idx >= 0).sum())
    cat_hist += np.bincount(a.cand_cat[live].astype(np.int64), minlength=16)[:16]
# ---- part 2: the campaign-shape instantiation (no per-candidate outputs): winners, costs, histograms and returned series
N2 = int(os.environ.get("N_QUERIES_2", 65536))
pl = BatchFrenetPlanner(CubicSpline2D(*scenarios.STRAIGHT_60), **scenarios.S1_KNOBS)
part2 = dict(queries=0, winner_mismatches=0, stats_mismatches=0, series_bit_mismatches=0, with_a_path=0)
for b0 in range(0, N2, 4096):
    n = min(4096, N2 - b0)
    rng = np.random.default_rng(SEED0 + 10_000_000 + b0)
    fs = np.stack([rng.uniform(2, 20, n), rng.uniform(0.0, 8.0, n), rng.uniform(-1, 1, n), rng.uniform(-1.0, 1.0, n),
                   rng.normal(0, 0.3, n), rng.normal(0, 0.05, n)], axis=1)
    dyn = np.stack([scenarios.pedestrian_field(rng, 50) for _ in range(n)])
    pl.engine
    res = {}
    for kern in ("pairs", "items"):
        with runners.fot_env(FOT_SWEEP=kern):
            res[kern] = pl.plan_batch(fs, 6.0, dynamic_obstacles=dyn)
    a, b = res["pairs"], res["items"]
    part2["queries"] += n
    part2["winner_mismatches"] += int((a.best_idx != b.best_idx).sum() + (a.best_cost.view(np.uint64) != b.best_cost.view(np.uint64)).sum())
    part2["stats_mismatches"] += int((a.stats != b.stats).any(axis=1).sum() + (a.winner_len != b.winner_len).sum())
    n_t = a.winner.shape[-1]
    livew = (np.arange(n_t)[None, None, :] < np.asarray(a.winner_len).reshape(-1, 1, 1)) & (np.asarray(a.best_idx).reshape(-1, 1, 1) >= 0)
    wa, wb = np.asarray(a.winner).reshape(n, -1, n_t), np.asarray(b.winner).reshape(n, -1, n_t)
    part2["series_bit_mismatches"] += int(((wa.view(np.uint64) != wb.view(np.uint64)) & livew).sum())
    part2["with_a_path"] += int((a.best_idx >= 0).sum())
out["campaign_shape_instantiation"] = part2
out["categories"] = {str(k): int(v) for k, v in enumerate(cat_hist) if v}
out["kinds_seen"] = sorted(out["kinds_seen"])
out["seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
