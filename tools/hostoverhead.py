import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from mr_rl_b200 import VecMREnv
env = VecMREnv(1024, device="cuda:0", noise="philox", seed=1, auto_reset=True)
env.want_state_prime = False
env.reset(init=None, noise_var=1.0, a0=1.0)
a = torch.rand(1024, 2, device="cuda:0", dtype=torch.float64)
for _ in range(100): env.step(a)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5000): env.step(a)
t1=time.perf_counter()-t
torch.cuda.synchronize()
print("host us per step call:", t1/5000*1e6)
