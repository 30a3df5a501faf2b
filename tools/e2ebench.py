import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mr_rl_b200 import VecMREnv
n = 1 << 20
for maxc in [int(v) for v in os.environ.get("E2E_MODES", "0,1,2,4,-2,-4").split(",")]:
    env = VecMREnv(n, device="cuda:0", noise="philox", seed=1, auto_reset=True)
    env.want_state_prime = False
    env.reset(init=None, noise_var=1.0, a0=1.0)
    env.host_mode = "direct" if maxc == 0 else "staged" if maxc > 0 else "staged_zc"   # < 0: kernels read the host actions
    env.host_chunks = abs(maxc)
    acts = [(torch.rand(n, 2, dtype=torch.float64) * torch.tensor([20.0, 6.28], dtype=torch.float64)).pin_memory() for _ in range(4)]
    for k in range(3): env.step_host(acts[k % 4])
    torch.cuda.synchronize(); t = time.perf_counter()
    for k in range(20): env.step_host(acts[k % 4])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
    print(f"chunks<={maxc}: {dt*1e3:.3f} ms per host step  {n/dt/1e9:.2f} Genv-steps/s  ({(16.8+26.2)/dt/1e3:.1f} GB/s PCIe both ways)")
