"""Latency of the single-env facade (MR_Env.reset / step returning numpy like the reference).  GPU box only."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from mr_rl_b200 import MR_Env

env = MR_Env(device="cuda:0")
env.reset(init=np.array([110.0, 105.0]), noise_var=1, a0=1)
rng = np.random.default_rng(0)
acts = np.stack([rng.uniform(0, 20, 2000), rng.uniform(0, 6.28, 2000)], 1)
for k in range(50):
    env.step(acts[k])
t0 = time.perf_counter()
n = 0
for k in range(50, 2000):
    obs, r, d, _ = env.step(acts[k]); n += 1
    if d:
        env.reset(init=np.array([110.0, 105.0]), noise_var=1, a0=1)
dt = time.perf_counter() - t0
print(f"MR_Env.step (single env, numpy in/out): {dt / n * 1e6:.1f} us per step  ({n / dt:.0f} steps/s; the reference on one host core: ~7e3-1.2e4 steps/s)")
