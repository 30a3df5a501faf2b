"""``run_sim`` (utils.py:43-61) as one fused K-step rollout launch."""
from __future__ import annotations

import numpy as np
import torch

from .vec_env import VecMREnv


def run_sim(actions, init_pos=None, noise_var=1, a0=1, is_mismatched=False, device="cuda", noise="philox", seed=0,
            noise_table=None, return_state_prime=False):
    """Open-loop rollout of ONE env through ``actions[T, >=2]`` (columns f, alpha[, t]); steps straight
    through ``done`` like the reference.  Returns X, Y, alpha, time, freq (numpy, length T)."""
    actions = np.asarray(actions, dtype=np.float64)
    env = VecMREnv(1, device=device, dtype=torch.float64, noise=noise, seed=seed, noise_table=noise_table,
                   time_table_len=max(4096, len(actions) + 8))
    if init_pos is None:
        init_pos = env.init_space.sample()
    env.reset(init=np.asarray(init_pos, dtype=np.float64), noise_var=noise_var, a0=a0, is_mismatched=is_mismatched)
    res = env.rollout(actions=torch.from_numpy(np.ascontiguousarray(actions[:, :2])), record=True,
                      record_state_prime=return_state_prime)
    xy = res["xy"][:, :, 0].cpu().numpy()
    env.check_status()
    X, Y = xy[:, 0].copy(), xy[:, 1].copy()
    alpha = actions[:, 1]
    freq = actions[:, 0]
    time = np.linspace(0, (len(X) - 1) / 30.0, len(X))          # utils.py:59
    if return_state_prime:
        return X, Y, alpha, time, freq, res["state_prime"][:, :, 0].cpu().numpy()
    return X, Y, alpha, time, freq
