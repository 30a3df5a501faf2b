"""Three data-parallel DDPG updates at batch 65536 and two tensor-core actor forwards over 2^20 envs, for ncu captures."""
import sys

import torch

sys.path.insert(0, ".")
from mr_rl_b200 import DDPGLearner, ReplayBuffer, VecMREnv  # noqa: E402

dev = "cuda:0"
env = VecMREnv(1 << 20, device=dev, noise="philox", seed=0, auto_reset=True)
env.reset(noise_var=1, a0=1)
rb = ReplayBuffer(1 << 20, 0, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
rb.s.normal_(0, 40, generator=g); rb.s2.copy_(rb.s); rb.a.uniform_(0, 6, generator=g); rb.r.fill_(10.0); rb.count = rb.buffer_size
learner = DDPGLearner(device=dev)
for _ in range(3):
    info = learner.update(rb, 65536)
for _ in range(2):
    a = learner.predict(env._obs, env.num_envs)
torch.cuda.synchronize()
print("critic loss", float(info[0]), "mean action", a.mean(0).tolist())
