"""Soak run of the randomised parity sweeps with seeds the test-suite does not use (GPU box; tests/ only samples 12).

    python tools/soak.py [--first 100] [--count 200]

Calls the sweep tests of tests/test_gpu_env.py (CUDA path against the plain-C oracle, every env: done flags and draw
counts exact, positions 1e-9) with fresh seeds and reports every failing (test, seed) instead of stopping at the first.
"""
import argparse
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_env as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=100)
ap.add_argument("--count", type=int, default=200)
a = ap.parse_args()



def failing_regime(seed):
    """Regimes in which some envs fail in the oracle (scipy would raise: the step size collapses near the origin under
    strong noise, or the draw table runs out): the SAME envs must carry a sticky status flag on the GPU, and every other
    env must match as usual (single-step kernel and fused rollout)."""
    import numpy as np
    import torch
    from oracle import c_oracle
    n, Tn = 256, 40
    rng, scale, sigma, a0, mism, init, acts = T._random_regime(seed, n, Tn)
    z = rng.standard_normal((n, 200 * Tn + 64))
    ref = c_oracle.rollout(init, acts, sigma, a0, mism=mism, mism_at_reset=False, z=z)
    if ref["bad"] == 0:
        return False
    moved = np.abs(ref["pos"]).sum(axis=2) > 0                     # a failed env is not stepped any further: rows stay 0
    steps_ok = moved.sum(axis=0)
    bad_env = (steps_ok < Tn) | (ref["cursor"] > z.shape[1])       # solver failure | draw table exhausted
    assert bad_env.sum() == ref["bad"], (bad_env.sum(), ref["bad"])
    for fused in (False, True):
        env = T.make_env(n, noise="table", noise_table=np.ascontiguousarray(z.T))
        env.reset(init=init, noise_var=sigma, a0=a0, is_mismatched=mism)
        a_dev = torch.as_tensor(acts, device="cuda:0")
        if fused:
            res = env.rollout(actions=a_dev, record=True, record_done=True)
            xy = res["xy"].cpu().numpy().transpose(0, 2, 1)
            dn = res["done_traj"].cpu().numpy()
        else:
            xy, dn = [], []
            for k in range(Tn):
                _, _, d, _ = env.step(a_dev[k])
                xy.append(env.last_pos.cpu().numpy().copy()); dn.append(d.cpu().numpy().copy())
            xy, dn = np.stack(xy), np.stack(dn)
        flagged = env._status[:n].cpu().numpy() != 0
        assert np.array_equal(flagged, bad_env), (np.nonzero(flagged)[0], np.nonzero(bad_env)[0], fused)
        good = ~bad_env
        assert np.array_equal(dn[:, good].astype(bool), ref["done"][:, good].astype(bool))
        assert np.array_equal(env._cursor[:n].cpu().numpy().astype(np.int64)[good], ref["cursor"][good])
        assert T.rel_err(xy[:, good], ref["pos"][:, good]) < T.FP64_TOL
        for i in np.nonzero(bad_env)[0]:                           # and the failed envs agree up to the failing step
            k = steps_ok[i]
            if k:
                assert T.rel_err(xy[:k, i], ref["pos"][:k, i]) < T.FP64_TOL
                assert np.array_equal(dn[:k, i].astype(bool), ref["done"][:k, i].astype(bool))
    return True


def sweep(seed):
    if not failing_regime(seed):
        T.test_random_regimes_against_c_oracle(seed)


def forced(seed, path):
    import numpy as np
    from oracle import c_oracle
    n = 3 * 128 + 34                                               # as in the test
    rng, scale, sigma, a0, mism, init, acts = T._random_regime(seed, n, 40)
    z = rng.standard_normal((n, 200 * 40 + 64))
    if c_oracle.rollout(init, acts, sigma, a0, mism=mism, mism_at_reset=False, z=z)["bad"]:
        return                                                     # covered by failing_regime()
    T.test_random_regimes_table_noise_both_step_kernels(seed, path)


cases = [("table noise, step + fused rollout", sweep)]
for path in ("scalar", "tma", "tmap"):
    cases.append((f"table noise, forced {path}", lambda s, p=path: forced(s, p)))
import torch  # noqa: E402
for dt in (torch.float64, torch.float32):
    cases.append((f"noise-free tiled kernel, {str(dt)[6:]}",
                  lambda s, d=dt: T.test_noise_free_tma_kernel_random_regimes_all_envs(s, d)))

bad, t0 = [], time.time()
for name, fn in cases:
    ok = 0
    for seed in range(a.first, a.first + a.count):
        try:
            fn(seed)
            ok += 1
        except Exception as e:                      # noqa: BLE001
            bad.append((name, seed, repr(e)[:300]))
            traceback.print_exc(limit=1)
    print(f"{name:40s} {ok}/{a.count} seeds ok  ({time.time() - t0:.0f} s)", flush=True)
print("FAILURES:" if bad else "no failures")
for b in bad:
    print("  ", b)
sys.exit(1 if bad else 0)
