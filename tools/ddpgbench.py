"""Time the device DDPG learner: one update launch at several batch sizes, and the vectorised acting/learning loop.
GPU box only.  Writes gpurun_out/ddpg.json."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from mr_rl_b200 import VecMREnv  # noqa: E402
from mr_rl_b200.ddpg import DDPGLearner, OUNoise, ReplayBuffer, train  # noqa: E402


def main():
    out = {}
    dev = "cuda:0"
    rb = ReplayBuffer(100_000, 0, device=dev)
    rb.s.normal_(0, 40); rb.s2.copy_(rb.s); rb.a.uniform_(0, 6); rb.r.fill_(10.0); rb.count = rb.buffer_size
    for batch in (64, 256, 1024, 4096, 16384, 65536):
        learner = DDPGLearner(device=dev)
        for _ in range(5):
            learner.update(rb, batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            learner.update(rb, batch)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out[f"update_batch_{batch}_us"] = us
        print(f"update batch {batch}: {us:.1f} us per launch ({batch / us:.2f} Msamples/s)", flush=True)
    for n_envs in (1, 4096, 65536):
        env = VecMREnv(n_envs, device=dev, noise="philox", seed=0, auto_reset=True)
        learner = DDPGLearner(device=dev)
        ou = OUNoise(n_envs, device=dev)
        rbt = ReplayBuffer(max(10000, 4 * n_envs), 0, device=dev)
        train(env, learner, ou, min_batch=64, steps=70 if n_envs == 1 else 10, replay=rbt)
        torch.cuda.synchronize()
        steps = 300
        t0 = time.perf_counter()
        train(env, learner, ou, min_batch=64, steps=steps, replay=rbt)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"train_{n_envs}_envs"] = {"iterations_per_s": steps / dt, "env_steps_per_s": steps * n_envs / dt,
                                       "updates_per_s": steps / dt, "ms_per_iteration": dt / steps * 1e3}
        print(f"train loop {n_envs} envs: {dt / steps * 1e3:.3f} ms per iteration (act + OU + step + ring add + 1 update), "
              f"{steps * n_envs / dt:.3g} env-steps/s", flush=True)
    # BASELINE configs[4] as full training: 2^20 envs collecting, one 65536-sample update per iteration
    n_envs = 1 << 20
    env = VecMREnv(n_envs, device=dev, noise="philox", seed=0, auto_reset=True)
    learner, ou = DDPGLearner(device=dev), OUNoise(n_envs, device=dev)
    rbt = ReplayBuffer(4 * n_envs, 0, device=dev)
    train(env, learner, ou, min_batch=65536, steps=5, replay=rbt)
    torch.cuda.synchronize()
    steps = 50
    t0 = time.perf_counter()
    train(env, learner, ou, min_batch=65536, steps=steps, replay=rbt)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["train_1M_envs_batch_65536"] = {"ms_per_iteration": dt / steps * 1e3, "env_steps_per_s": steps * n_envs / dt,
                                        "learner_samples_per_s": steps * 65536 / dt}
    print(f"train loop 2^20 envs, batch 65536: {dt / steps * 1e3:.3f} ms per iteration, {steps * n_envs / dt:.3g} env-steps/s, "
          f"{steps * 65536 / dt:.3g} learner samples/s", flush=True)
    json.dump(out, open("gpurun_out/ddpg.json", "w"), indent=1)


if __name__ == "__main__":
    main()
