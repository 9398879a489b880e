"""Access to the UNMODIFIED reference implementation -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this
module; the product package `integrated_path_planning_b200` never does.

The reference is pure Python.  It is found at `/root/reference` in the build container and at `oracle/_ref/`
(the byte-for-byte copy made by `oracle/make_ref.py`, git-ignored, shipped with the gpurun snapshot) on the GPU
box.  `pysocialforce` (the reference simulator's ground-truth pedestrian model, not installed anywhere here) is
stubbed with an empty module: the harness replays recorded / constant-velocity pedestrians through the
reference's own `ReplayPedestrianSource` instead.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.environ.get("FOT_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref"))
_root: Optional[str] = None


def reference_root() -> Optional[str]:
    """Directory holding the reference's `src/` package, or None when there is none on this machine."""
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "src", "planning")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def load():
    """Import the reference; returns a namespace with its public classes.  Raises RuntimeError when absent."""
    global _root
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not available: neither /root/reference nor oracle/_ref exists "
                           "(run `python oracle/make_ref.py` in the build container)")
    if _root is None:
        # appended, not prepended: the reference's top-level package is `src`; its own `tests` package (present in
        # /root/reference) must not shadow this repository's
        sys.path.append(root)
        sys.modules.setdefault("pysocialforce", types.ModuleType("pysocialforce"))
        from loguru import logger
        logger.remove()
        _root = root
    from src.core.data_structures import EgoVehicleState
    from src.core.footprint import EgoFootprint
    from src.planning.cubic_spline import CubicSpline2D
    from src.planning.frenet_planner import FrenetPlanner
    return types.SimpleNamespace(root=root, EgoVehicleState=EgoVehicleState, EgoFootprint=EgoFootprint,
                                 CubicSpline2D=CubicSpline2D, FrenetPlanner=FrenetPlanner)


def kind() -> str:
    """`reference` when the unmodified reference can be imported here, else `port` (the NumPy oracle)."""
    return "reference" if available() else "port"


class IndexRecorder:
    """Learns WHICH candidate a stock `plan()` call returned, without changing what it computes.

    `FrenetPlanner.plan()` returns the winning FrenetPath, not its generation-order index.  This wraps two bound
    methods of one planner INSTANCE (`_generate_frenet_paths`, frenet_planner.py:376, and `_select_best_path`,
    :1235) with pass-through recorders: the candidate list as generated and the object selected; the index is the
    selected object's position in that list.  No reference source is touched and the arithmetic is the stock
    call's own."""

    def __init__(self, planner):
        self.planner = planner
        self.last_index = -1
        self.last_n_candidates = 0
        self._fps = None
        gen, sel = planner._generate_frenet_paths, planner._select_best_path

        def gen_rec(*a, **k):
            self._fps = gen(*a, **k)
            return self._fps

        def sel_rec(*a, **k):
            best = sel(*a, **k)
            fps = self._fps or []
            self.last_n_candidates = len(fps)
            self.last_index = -1
            if best is not None:
                for i, fp in enumerate(fps):
                    if fp is best:
                        self.last_index = i
                        break
            return best

        planner._generate_frenet_paths = gen_rec
        planner._select_best_path = sel_rec

    def plan(self, *a, **k):
        self.last_index, self.last_n_candidates, self._fps = -1, 0, None
        return self.planner.plan(*a, **k)


def scenario_simulator(scenario: str = "scenario_01_cv", footprint: bool = False, overrides: Optional[dict] = None,
                       planner_cls=None):
    """The reference's IntegratedSimulator on one of its own scenario files, made runnable offline (SURVEY.md
    section 8c): pysocialforce stubbed, SGAN weights not required (`prediction_method = cv`), the YAML's pedestrians
    replayed at their constant initial velocity through the reference's ReplayPedestrianSource.
    `planner_cls`: the drop-in of INTEGRATION.md section 1 -- `sim_mod.FrenetPlanner = planner_cls` while the
    simulator is constructed (the reference constructs its planner in exactly one place,
    integrated_simulator.py:342-366), restored afterwards.
    Returns (sim, cfg, module) where `module` is src.simulation.integrated_simulator."""
    ref = load()
    import yaml
    from src.config import SimulationConfig, validate_config
    import src.simulation.integrated_simulator as sim_mod
    from src.simulation.replay_source import ReplayPedestrianSource
    with open(os.path.join(ref.root, "scenarios", scenario + ".yaml")) as f:
        d = yaml.safe_load(f)
    peds = np.array(d.pop("ped_initial_states"), dtype=float)
    d.pop("ped_groups", None)
    d["sgan_model_path"] = None
    d["visualization_enabled"] = False
    d["prediction_method"] = "cv"
    if footprint:
        d["ego_footprint"] = "multi_circle"
    if overrides:
        d.update(overrides)
    cfg = SimulationConfig(**d)
    validate_config(cfg)
    stock = sim_mod.FrenetPlanner
    try:
        if planner_cls is not None:
            sim_mod.FrenetPlanner = planner_cls
        sim = sim_mod.IntegratedSimulator(cfg)
    finally:
        sim_mod.FrenetPlanner = stock
    n_frames = int(cfg.total_time / cfg.dt) + 200
    t = np.arange(n_frames)[:, None, None] * cfg.dt
    traj = peds[None, :, 0:2] + peds[None, :, 2:4] * t
    sim.pedestrian_sim = ReplayPedestrianSource(traj, dt=cfg.dt)
    return sim, cfg, sim_mod
