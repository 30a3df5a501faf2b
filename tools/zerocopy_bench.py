"""Experiment: which direction limits the direct (zero-copy) host step?  Times the step kernel with
(a) host actions + host outputs, (b) device actions + host outputs, (c) host actions + device outputs."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mr_rl_b200 import VecMREnv, _lib as L

n = 1 << 20
env = VecMREnv(n, device="cuda:0", noise="philox", seed=1, auto_reset=True)
env.reset(init=None, noise_var=1.0, a0=1.0)
acts = [(torch.rand(n, 2, dtype=torch.float64) * torch.tensor([20.0, 6.28], dtype=torch.float64)).pin_memory() for _ in range(4)]
acts_dev = [a.cuda() for a in acts]
o_pin = torch.zeros(5, n, dtype=torch.float64).pin_memory()
r_pin = torch.zeros(n, dtype=torch.float64).pin_memory()
d_pin = torch.zeros(n, dtype=torch.uint8).pin_memory()
out_host = L.StepOut(o_pin.data_ptr(), r_pin.data_ptr(), d_pin.data_ptr(), None, n, 1, 0)
lib = env.lib


def run(a_list, out, label):
    def one(k):
        nz = env._noise_for(env.params.noise_var)
        rc = lib.mr_env_step(env._b_state, n, env._dt, env._b_params, C.byref(nz), env._b_tt, a_list[k % 4].data_ptr(), out,
                             torch.cuda.current_stream().cuda_stream)
        L.check(rc, "mr_env_step")
        env._step_index += 1
        torch.cuda.current_stream().synchronize()
    for k in range(3):
        one(k)
    t = time.perf_counter()
    for k in range(20):
        one(k)
    dt = (time.perf_counter() - t) / 20
    print(f"{label}: {dt*1e3:.3f} ms  {n/dt/1e9:.2f} Genv-steps/s", flush=True)


run(acts, C.byref(out_host), "host actions, host outputs ")
run(acts_dev, C.byref(out_host), "device actions, host outputs")
run(acts, env._b_out_lean, "host actions, device outputs")
run(acts_dev, env._b_out_lean, "device actions, device outputs")
for path in ("vec", "scalar"):
    os.environ["MR_STEP_PATH"] = path
