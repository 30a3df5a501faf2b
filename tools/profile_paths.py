"""Small, fixed workload touching every kernel once or twice (for ncu / compute-sanitizer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import DeviceGP, VecMREnv, init_actor, pack_actor

n = int(os.environ.get("PROF_ENVS", 1 << 18))
K = int(os.environ.get("PROF_K", 16))
nq = int(os.environ.get("PROF_NQ", 16384))
ntr = int(os.environ.get("PROF_NTRAIN", 2000))

env = VecMREnv(n, device="cuda:0", noise="philox", seed=1, auto_reset=True)
env.reset(init=None, noise_var=1.0, a0=1.0)
acts = torch.rand(n, 2, device="cuda:0", dtype=torch.float64) * torch.tensor([20.0, 6.28], device="cuda:0", dtype=torch.float64)
for _ in range(3):
    env.step(acts)
for _ in range(2):
    env.rollout(policy="random", k_steps=K)
packed = pack_actor(init_actor(0), "cuda:0")
for _ in range(2):
    env.rollout(policy=packed, k_steps=K)
env.check_status()

rng = np.random.default_rng(0)
X = np.sort(rng.uniform(-np.pi, np.pi, ntr))
y = 0.2 + 0.5 * np.cos(X + 0.3) + 0.09 * rng.standard_normal(ntr)
gp = DeviceGP.fit(X, y, 0.2, 0.008, device="cuda:0")
q = torch.rand(nq, device="cuda:0", dtype=torch.float64) * 6.28 - 3.14
for _ in range(2):
    mean, std = gp.predict(q, True)
rows = gp.enable_spectral_variance()          # verified low-rank form: the fused posterior kernel
for _ in range(2):
    mean, std = gp.predict(q, True)
torch.cuda.synchronize()
print("ok", float(mean.mean()), float(std.mean()), env.stats_dict()["env_steps"])
