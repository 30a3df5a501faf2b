"""Debug: dump the step's and the auto reset's draws for the same (env, step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mr_rl_b200 import VecMREnv
n = 4096
out = {}
for tag, ar in (("plain", 0), ("reset", 1)):
    env = VecMREnv(n, device="cuda:0", noise="philox", seed=4, auto_reset=True)
    env.max_timesteps = 0
    env.reset(init=None, noise_var=1.0, a0=0.0)
    out[tag + "_f_after_reset"] = env._state[2:4, :n].cpu().numpy().copy()
    out[tag + "_sp_after_reset"] = env.state_prime.cpu().numpy().copy()
    env.params.auto_reset = ar
    a = torch.zeros(n, 2, dtype=torch.float64, device="cuda:0")
    env.step(a)
    out[tag + "_f"] = env._state[2:4, :n].cpu().numpy().copy()
    out[tag + "_sp"] = env.state_prime.cpu().numpy().copy()
    out[tag + "_xy"] = env.last_pos.cpu().numpy().copy()
    out[tag + "_counter"] = env.counter.cpu().numpy().copy()
np.savez("gpurun_out/dbg_reset.npz", **out)
for k, v in out.items():
    print(k, v.shape, v.flatten()[:4])
