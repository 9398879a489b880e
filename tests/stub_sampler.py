"""A seeded stand-in for the reference's SGAN generator (the generator itself is outside the hot path).

`StubGenerator` has the call signature of `TrajectoryPredictor.generator` (trajectory_predictor.py:175): it returns
relative displacements `[pred_len, n_peds, 2]` float32 -- the last observed displacement plus seeded Gaussian noise,
a different draw on every call.  tests/golden/make_golden_rollout.py installs it in the unmodified reference
simulator (`predictor.method = 'sgan'`, `num_samples = S`) to record distribution-aware roll-outs;
`batched_sampler` restates, for N simulations at once, what the reference then does with the generator's output
(`relative_to_abs`, sgan_vendor/utils.py:9-23, float32) so that `BatchedClosedLoop(sampler=...)` sees the same raw
sample sets.  Nothing here reads the reference.
"""
from __future__ import annotations

import numpy as np
import torch


class StubGenerator:
    def __init__(self, seed: int, pred_len: int = 12, sigma: float = 0.05):
        self.seed, self.pred_len, self.sigma, self.calls = int(seed), int(pred_len), float(sigma), 0

    def noise(self, n_peds: int) -> np.ndarray:
        rng = np.random.default_rng(self.seed + self.calls)
        self.calls += 1
        return rng.normal(0.0, self.sigma, (self.pred_len, n_peds, 2)).astype(np.float32)

    def __call__(self, obs_traj, obs_traj_rel, seq_start_end):
        return obs_traj_rel[-1][None, :, :] + torch.from_numpy(self.noise(obs_traj.shape[1]))

    # the reference calls .to(device) / .eval() on real generators only in load_model; nothing else is needed


def batched_sampler(seeds, num_samples: int, pred_len: int = 12, sigma: float = 0.05):
    """-> sampler(obs [n, obs_len, P, 2] float64 observer history of the simulations `idx`, idx) ->
    raw absolute predictions [n, S, pred_len, P, 2] float64 (exactly the float32 values the reference computes)."""
    gens = {int(i): StubGenerator(s, pred_len, sigma) for i, s in enumerate(seeds)}

    def sampler(obs: np.ndarray, idx) -> np.ndarray:
        n, obs_len, P, _ = obs.shape
        out = np.empty((n, num_samples, pred_len, P, 2), dtype=np.float64)
        for j, i in enumerate(idx):
            obs64 = obs[j]
            rel64 = np.zeros_like(obs64)
            rel64[1:] = obs64[1:] - obs64[:-1]                       # observer.py:126-128 (float64, then .float())
            obs32, rel32 = torch.from_numpy(obs64).float(), torch.from_numpy(rel64).float()
            for s in range(num_samples):
                pred_rel = rel32[-1][None, :, :] + torch.from_numpy(gens[int(i)].noise(P))
                disp = torch.cumsum(pred_rel.permute(1, 0, 2), dim=1)    # relative_to_abs
                out[j, s] = (disp + obs32[-1].unsqueeze(1)).permute(1, 0, 2).numpy().astype(np.float64)
        return out

    return sampler
