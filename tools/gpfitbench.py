"""Time DeviceGP.fit (fixed hyper-parameters) against sklearn on the host.  GPU box only."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mr_rl_b200.gp import DeviceGP  # noqa: E402


def main():
    out = []
    for n in (512, 2048, 4096, 8192):
        rng = np.random.default_rng(n)
        X = rng.uniform(-np.pi, np.pi, size=(n, 1))
        y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
        Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
        for _ in range(2):
            gp = DeviceGP.fit(Xd, yd, 0.5, 0.01)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            gp = DeviceGP.fit(Xd, yd, 0.5, 0.01)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row = {"n_train": n, "device_fit_ms": ms, "flops_chol_plus_inv": n ** 3 * (1 / 3 + 1 / 3) , "lml": gp.log_marginal_likelihood_value_}
        if n <= 4096:
            from sklearn.gaussian_process import GaussianProcessRegressor
            from sklearn.gaussian_process.kernels import RBF, WhiteKernel
            t0 = time.perf_counter()
            ref = GaussianProcessRegressor(kernel=RBF(0.5) + WhiteKernel(0.01), optimizer=None).fit(X, y)
            row["sklearn_fit_ms"] = (time.perf_counter() - t0) * 1e3
            row["lml_sklearn"] = float(ref.log_marginal_likelihood(ref.kernel_.theta))
        out.append(row)
        print(json.dumps(row), flush=True)
    json.dump(out, open("gpurun_out/gpfit.json", "w"), indent=1)


if __name__ == "__main__":
    main()
