"""Dev tool: BASELINE configs[4] (actor in the loop) for each in-kernel actor implementation.

    python tools/actorbench.py [--n 1048576] [--k 50] [--launches 4] [--paths default tf32]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mr_rl_b200 import VecMREnv, _lib as L, init_actor, pack_actor

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--sigma", type=float, default=1.0)
ap.add_argument("--paths", nargs="+", default=["default", "tf32"])
a = ap.parse_args()
packed = pack_actor(init_actor(0), "cuda:0")
for path in a.paths:
    L.set_actor_path(path)
    env = VecMREnv(a.n, device="cuda:0", noise="philox" if a.sigma else "none", seed=11, auto_reset=True)
    env.reset(init=None, noise_var=a.sigma, a0=1.0)
    env.rollout(policy=packed, k_steps=a.k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.launches):
        env.rollout(policy=packed, k_steps=a.k)
    e1.record()
    torch.cuda.synchronize()
    env.check_status()
    ms = e0.elapsed_time(e1)
    print(f"actor path {path:8s}: {a.n} envs x {a.k * a.launches} steps in {ms:8.2f} ms = {a.n * a.k * a.launches / ms / 1e6:8.2f} Genv-steps/s",
          flush=True)
L.set_actor_path("default")
