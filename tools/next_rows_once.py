"""One GP fit (+ gradient) at n_train = 2048 and three DDPG updates at batch 64, for ncu captures.  GPU box only."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mr_rl_b200.ddpg import DDPGLearner, ReplayBuffer  # noqa: E402
from mr_rl_b200.gp import DeviceGP  # noqa: E402

rng = np.random.default_rng(0)
X = rng.uniform(-np.pi, np.pi, size=(2048, 1))
y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(2048)
for _ in range(2):
    gp = DeviceGP.fit(X, y, 0.5, 0.01, eval_gradient=True)
rb = ReplayBuffer(10000, 0, device="cuda:0")
rb.s.normal_(0, 40); rb.s2.copy_(rb.s); rb.a.uniform_(0, 6); rb.r.fill_(10.0); rb.count = rb.buffer_size
learner = DDPGLearner(device="cuda:0")
for _ in range(3):
    info = learner.update(rb, 64)
torch.cuda.synchronize()
print("lml", gp.log_marginal_likelihood_value_, "critic loss", float(info[0]))
