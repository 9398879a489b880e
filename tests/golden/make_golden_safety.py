"""Record golden vectors for the safety metrics from the UNMODIFIED reference
(src/core/data_structures.py:301-388).  Build container only:  python tests/golden/make_golden_safety.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from loguru import logger  # noqa: E402

logger.remove()
from src.core.data_structures import EgoVehicleState, PedestrianState, compute_safety_metrics_static  # noqa: E402
from src.core.footprint import EgoFootprint  # noqa: E402


def main():
    rng = np.random.default_rng(11)
    n_q, P = 48, 17
    ego = np.stack([rng.uniform(0, 40, n_q), rng.uniform(-2, 2, n_q), rng.normal(0, 0.6, n_q), rng.uniform(0, 8, n_q),
                    rng.uniform(-1, 1, n_q)], axis=1)
    pos = np.stack([rng.uniform(-5, 50, (n_q, P)), rng.uniform(-8, 8, (n_q, P))], axis=-1)
    vel = rng.normal(0, 1.2, (n_q, P, 2))
    pos[3, :4] = ego[3, :2] + rng.normal(0, 0.3, (4, 2))         # a collision
    n_peds = rng.integers(0, P + 1, n_q).astype(np.int32)
    n_peds[:4] = [0, 1, P, P]
    fp = EgoFootprint.multi_circle(4.5, 2.0, 3)
    keys = ("min_distance", "collision", "ttc", "clearance", "clearance_ahead")
    out = {False: np.zeros((n_q, 5)), True: np.zeros((n_q, 5))}
    for use_fp in (False, True):
        for q in range(n_q):
            e = EgoVehicleState(x=ego[q, 0], y=ego[q, 1], yaw=ego[q, 2], v=ego[q, 3], a=ego[q, 4])
            k = int(n_peds[q])
            ps = PedestrianState(positions=pos[q, :k], velocities=vel[q, :k], goals=np.zeros((k, 2)))
            m = compute_safety_metrics_static(e, ps, 1.0, 0.2, fp if use_fp else None)
            out[use_fp][q] = [float(m[key]) for key in keys]
    np.savez_compressed(os.path.join(HERE, "safety.npz"), ego=ego, pos=pos, vel=vel, n_peds=n_peds, single=out[False],
                        footprint=out[True], fp_offsets=fp.offsets, fp_radius=np.array([fp.radius]))
    print("wrote safety.npz")


if __name__ == "__main__":
    main()
