"""Small, ragged workload over the env kernels (no tensor-core / TMEM kernels) for compute-sanitizer memcheck:
every step-kernel variant incl. table rows staged by TMA with a scalar tail, reset, fused rollout with recording, per-env rows,
float32 wire rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import VecMREnv, _lib as L

n, T = 128 * 5 + 37, 6
rng = np.random.default_rng(0)
acts = torch.as_tensor(np.stack([rng.uniform(0, 20, (T, n)), rng.uniform(0, 6.28, (T, n))], -1), device="cuda:0")
z = rng.standard_normal((40 * T + 8, n))
for path in ("scalar", "vec", "tma"):
    L.set_step_path(path)
    for kind, kw in (("none", {}), ("philox", {}), ("table", {"noise_table": z})):
        for dt in (torch.float64, torch.float32):
            env = VecMREnv(n, device="cuda:0", dtype=dt, noise=kind, auto_reset=True, **kw)
            env.max_timesteps = 3
            env.reset(init=None, noise_var=0.0 if kind == "none" else 1.0, a0=1.0)
            for k in range(T):
                env.step(acts[k].to(dt))
            env.rollout(actions=acts.to(dt), record_episodes=True)
            env.rollout(policy="random", k_steps=4)
L.set_step_path("default")
env = VecMREnv(n, device="cuda:0", noise="philox", per_env_params=True)
env.reset(init=None, noise_var=rng.uniform(0, 1, n), a0=rng.uniform(0.5, 2, n), is_mismatched=(rng.random(n) < 0.5).astype(np.uint8))
for k in range(T):
    env.step(acts[k])
env.rollout(actions=acts)
env2 = VecMREnv(n, device="cuda:0", noise="philox")
env2.reset(init=None, noise_var=1.0, a0=1.0)
env2.host_io_dtype = torch.float32
env2.step_host(acts[0].float().cpu().numpy())
env2.host_io_dtype = None
env2.step_host(acts[0].cpu().numpy())
torch.cuda.synchronize()
print("sanitize workload ok")
