"""Two device GP fits at n_train = 2048 (for an ncu launch list).  GPU box only."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mr_rl_b200.gp import DeviceGP  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(n)
X = rng.uniform(-np.pi, np.pi, size=(n, 1))
y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
for _ in range(2):
    gp = DeviceGP.fit(X, y, 0.5, 0.01, eval_gradient=True)
torch.cuda.synchronize()
print("lml", gp.log_marginal_likelihood_value_, gp.log_marginal_likelihood_gradient_)
