"""Dev tool: the parity (table-noise) mode of the single-step path at full size — the tiled TMA kernel with the noise
rows bulk-copied per tile against the scalar kernel (281 B per env-step in fp64).

    python tools/tablebench.py [--n 1048576] [--steps 24]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import VecMREnv, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--paths", nargs="+", default=["tmap", "tma", "scalar"])
a = ap.parse_args()
n, K = a.n, a.steps
rows = 4 + 16 * K
gen = torch.Generator(device="cuda:0").manual_seed(7)
table = torch.randn(rows, n, generator=gen, device="cuda:0", dtype=torch.float64)
acts = torch.rand(8, n, 2, generator=gen, device="cuda:0", dtype=torch.float64)
acts[..., 0] *= 20.0
acts[..., 1] *= 2 * np.pi
for path in a.paths:
    L.set_step_path(path)
    env = VecMREnv(n, device="cuda:0", noise="table", noise_table=table, auto_reset=False)
    env.want_state_prime = False
    for rep in range(3):                                       # two warm cycles, the third is timed
        env.reset(init=None, noise_var=1.0, a0=1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            env.step(acts[k % 8])
        e1.record()
        torch.cuda.synchronize()
    env.check_status()
    assert int(env._cursor[:n].min()) == int(env._cursor[:n].max()) == rows
    ms = e0.elapsed_time(e1) / K
    print(f"table noise, path {path:7s}: {ms * 1e3:7.1f} us per step = {n / ms / 1e6:7.2f} Genv-steps/s, "
          f"{281 * n / (ms * 1e-3) / 1e9:6.0f} GB/s algorithmic (281 B per env-step)", flush=True)
L.set_step_path("default")
