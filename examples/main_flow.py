"""The reference's simulation study (main.py) on the device, without the plots.

    python examples/main_flow.py [--steps 600] [--envs 1]

1. idle run                 -> LearningModule.estimateDisturbance          (main.py:56-63)
2. circle run, mismatched   -> LearningModule.learn (sklearn fit on host)  (main.py:66-75)
3. desired / baseline runs  -> utils.run_sim equivalents                    (main.py:119-131)
4. corrected headings       -> LearningModule.predict for every control step, batched on the GPU (main.py:145-155)
5. corrected run            -> compare the tracking error with and without the GP correction
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_rl_b200 import LearningModule, run_sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600, help="training steps (the reference uses 1800)")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    freq, a0_def, dt, noise_var = 4.0, 1.5, 0.030, 0.5                 # main.py:9-11,53
    idle = np.zeros((100, 3)); idle[:, 2] = np.arange(100) * dt
    cyc = a.steps // 3
    circle = np.zeros((cyc, 3)); circle[:, 0] = freq; circle[:, 1] = np.linspace(-np.pi, np.pi, cyc)
    learn = np.vstack([circle] * 3); learn[:, 2] = np.arange(len(learn)) * dt
    T = 500
    test = np.zeros((T, 3)); test[:, 0] = freq; test[:, 2] = np.arange(T) * dt
    test[:, 1] = np.concatenate([np.linspace(0, np.pi / 2, 100), np.linspace(np.pi / 2, -np.pi / 2, 100),
                                 np.linspace(-np.pi / 2, 0, 100), np.linspace(0, np.pi / 8, 100), np.linspace(np.pi / 8, -np.pi, 100)])
    kw = dict(device="cuda:0", noise="philox")
    t0 = time.perf_counter()
    gp = LearningModule(device="cuda:0")
    gp.gprX.n_restarts_optimizer = gp.gprY.n_restarts_optimizer = 1
    px, py, _, tm, _ = run_sim(idle, init_pos=np.array([0, 0]), noise_var=noise_var, a0=a0_def, is_mismatched=True, seed=a.seed, **kw)
    gp.estimateDisturbance(px, py, tm)
    px, py, al, tm, _ = run_sim(learn, init_pos=np.array([0, 0]), noise_var=noise_var, a0=a0_def, is_mismatched=True, seed=a.seed + 1, **kw)
    a0_sim = gp.learn(px, py, al, tm, learn)
    t_learn = time.perf_counter() - t0
    print(f"drift D = ({gp.Dx:.3f}, {gp.Dy:.3f}), a0 estimate {a0_sim:.3f}, GP kernels {gp.gprX.kernel_} | {gp.gprY.kernel_}  [{t_learn:.1f} s]")

    xd, yd, *_ = run_sim(test, init_pos=np.array([0, 0]), noise_var=0.0, a0=a0_sim, **kw)                 # desired
    xb, yb, *_ = run_sim(test, init_pos=np.array([0, 0]), noise_var=noise_var, a0=a0_def, is_mismatched=True, seed=a.seed + 2, **kw)
    vd = a0_sim * freq * np.stack([np.cos(test[:, 1]), np.sin(test[:, 1])], 1)                            # main.py:147
    t0 = time.perf_counter()
    alpha, mux, muy, sgx, sgy = gp.predict_batch(vd)                                                      # all T control steps at once
    torch.cuda.synchronize()
    t_pred = time.perf_counter() - t0
    corrected = test.copy(); corrected[:, 1] = alpha.cpu().numpy()
    xl, yl, *_ = run_sim(corrected, init_pos=np.array([0, 0]), noise_var=noise_var, a0=a0_def, is_mismatched=True, seed=a.seed + 2, **kw)
    err_b = np.hypot(xb - xd, yb - yd)
    err_l = np.hypot(xl - xd, yl - yd)
    print(f"corrected headings for {T} control steps in {t_pred*1e3:.1f} ms (sigma of the GP: {float(sgx.mean()):.3f}, {float(sgy.mean()):.3f})")
    print(f"final tracking error  baseline {err_b[-1]:.2f}  corrected {err_l[-1]:.2f}   mean  baseline {err_b.mean():.2f}  corrected {err_l.mean():.2f}")


if __name__ == "__main__":
    main()
