"""Dev tool: the single-step kernel at the per-rank sizes of the strong-scaling run (2^20 envs over 8 / 4 / 2 ranks),
K launches as one CUDA-graph replay — what a rank of `bench.py --gpus N` times, on one GPU.

    [MR_STEP_GRID=b|o] [MR_STEP_PATH=...] python tools/smallbench.py [--k 50] [--reps 40]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import VecMREnv


def run(n, sigma, k, reps, dtype=torch.float64):
    env = VecMREnv(n, device="cuda:0", dtype=dtype, noise="philox" if sigma else "none", seed=1, auto_reset=True)
    env.want_state_prime = False
    env.reset(init=None, noise_var=sigma, a0=1.0)
    acts = torch.rand(8, n, 2, device="cuda:0", dtype=torch.float64)
    acts[..., 0] *= 20; acts[..., 1] *= 2 * np.pi
    acts = acts.to(dtype)
    g = env.capture_steps([acts[j] for j in range(8)], k)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * k) * 1e3          # us per step


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--sizes", default="32768,65536,131072,262144,524288")
    a = ap.parse_args()
    tag = "grid=%s path=%s" % (os.environ.get("MR_STEP_GRID", "default"), os.environ.get("MR_STEP_PATH", "default"))
    for n in [int(v) for v in a.sizes.split(",")]:
        row = []
        for sigma in (0.0, 1.0):
            us = run(n, sigma, a.k, a.reps)
            row.append("sigma=%g %6.2f us (%6.2f Genv-steps/s)" % (sigma, us, n / us / 1e3))
        print("%-28s n=%7d  %s" % (tag, n, "   ".join(row)), flush=True)
