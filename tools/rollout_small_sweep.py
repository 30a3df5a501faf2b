"""Dev tool: fused K = 64 rollout (tensor actions, sigma = 1) over batch sizes, for the register-cap threshold
(MR_ROLLOUT_SMALL=0|1 forces the capped / uncapped instantiation)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from mr_rl_b200 import VecMREnv

for n in (4096, 16384, 32768, 49152, 56832, 65536, 98304, 131072, 262144):
    env = VecMREnv(n, device="cuda:0", noise="philox", seed=1, auto_reset=True)
    env.reset(init=None, noise_var=1.0, a0=1.0)
    acts = torch.rand(64, n, 2, device="cuda:0", dtype=torch.float64)
    acts[..., 0] *= 20; acts[..., 1] *= 2 * np.pi
    for _ in range(3):
        env.rollout(actions=acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for _ in range(reps):
        env.rollout(actions=acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"small={os.environ.get('MR_ROLLOUT_SMALL', 'auto')} n={n:7d}  {ms*1e3:8.1f} us  {n*64/ms/1e6:7.2f} G/s", flush=True)
