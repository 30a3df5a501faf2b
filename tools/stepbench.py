"""Dev tool: time the single-step / rollout kernels for a few configurations (CUDA events).

    python tools/stepbench.py [--n 1048576] [--steps 200]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import VecMREnv

BYTES = {torch.float64: 153, torch.float32: 81}


def time_step(n, dtype, sigma, steps, auto_reset=True, sp=False):
    env = VecMREnv(n, device="cuda:0", dtype=dtype, noise="philox" if sigma else "none", seed=1, auto_reset=auto_reset)
    env.want_state_prime = sp
    env.reset(init=None, noise_var=sigma, a0=1.0)
    acts = torch.rand(8, n, 2, device="cuda:0", dtype=torch.float64)
    acts[..., 0] *= 20; acts[..., 1] *= 2 * np.pi
    acts = acts.to(dtype)
    for k in range(10):
        env.step(acts[k % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, BYTES[dtype] * n / (ms * 1e-3) / 1e9


def time_rollout(n, dtype, sigma, K, reps=3):
    env = VecMREnv(n, device="cuda:0", dtype=dtype, noise="philox" if sigma else "none", seed=1, auto_reset=True)
    env.reset(init=None, noise_var=sigma, a0=1.0)
    env.rollout(policy="random", k_steps=K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(policy="random", k_steps=K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, n * K / (ms * 1e-3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--rollout", action="store_true")
    a = ap.parse_args()
    tag = os.environ.get("MR_STEP_PATH", "default")
    # box calibration: plain device copy of the same footprint (63 MB read + 97 MB written ~ 80 MB copy)
    src = torch.empty(80 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
    dst = torch.empty_like(src)
    for _ in range(5):
        dst.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dst.copy_(src)
    e1.record()
    torch.cuda.synchronize()
    cms = e0.elapsed_time(e1) / 50
    print(f"calibration: 160 MB copy traffic in {cms*1e3:.1f} us = {160*1.048576/cms:.0f} GB/s", flush=True)
    for dtype in (torch.float64, torch.float32):
        for sigma in (0.0, 1.0):
            ms, gbs = time_step(a.n, dtype, sigma, a.steps)
            print(f"step vec={tag} {str(dtype)[6:]:8s} sigma={sigma} {ms*1e3:8.1f} us  {a.n/ms/1e6:8.2f} Genv-steps/s  {gbs:7.0f} GB/s algorithmic", flush=True)
    if a.rollout:
        for dtype in (torch.float64,):
            for sigma in (0.0, 1.0):
                ms, eps = time_rollout(a.n, dtype, sigma, 64)
                print(f"rollout K=64 {str(dtype)[6:]:8s} sigma={sigma} {ms:8.3f} ms  {eps/1e9:8.2f} Genv-steps/s", flush=True)
