"""Dev tool: the fused K = 64 rollout with in-kernel random actions and with actions from a tensor [K][N][2]."""
import os, sys
sys.path.insert(0, os.getcwd())
sys.argv = ["stepbench"]
import numpy as np
import torch
import tools.stepbench as sb
from mr_rl_b200 import VecMREnv


def time_tensor_actions(n, sigma, K, reps=3):
    env = VecMREnv(n, device="cuda:0", noise="philox" if sigma else "none", seed=1, auto_reset=True)
    env.reset(init=None, noise_var=sigma, a0=1.0)
    acts = torch.rand(K, n, 2, device="cuda:0", dtype=torch.float64)
    acts[..., 0] *= 20; acts[..., 1] *= 2 * np.pi
    env.rollout(actions=acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(actions=acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, n * K / (ms * 1e-3)


for sigma in (0.0, 1.0):
    ms, eps = sb.time_rollout(1 << 20, torch.float64, sigma, 64)
    print(f"rollout K=64 sigma={sigma} random actions        {ms:8.3f} ms {eps/1e9:8.2f} G/s", flush=True)
for n in (1 << 20, 4096):
    for sigma in (0.0, 1.0):
        ms, eps = time_tensor_actions(n, sigma, 64, reps=3 if n > 4096 else 50)
        print(f"rollout K=64 sigma={sigma} tensor actions n={n:8d} {ms:8.3f} ms {eps/1e9:8.2f} G/s", flush=True)
