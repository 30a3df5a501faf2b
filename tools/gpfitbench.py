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
    # the reference's fit (Learning_module.py:28-33,122): ~2k samples, 5 restarts
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200.gpr import DeviceGPR
    n = 1970
    rng = np.random.default_rng(1)
    X = np.sort(rng.uniform(-np.pi, np.pi, size=(n, 1)), axis=0)
    y = 0.8 * np.sin(2 * X[:, 0]) + 0.3 * np.cos(X[:, 0]) + 0.15 * rng.standard_normal(n)
    DeviceGPR(n_restarts_optimizer=0, random_state=3).fit(X[:256], y[:256])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = DeviceGPR(n_restarts_optimizer=5, random_state=3).fit(X, y)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = GaussianProcessRegressor(kernel=RBF(1.0, (1e-2, 10.0)) + WhiteKernel(), n_restarts_optimizer=5, random_state=3).fit(X, y)
    t_ref = time.perf_counter() - t0
    row = {"search_n_train": n, "restarts": 5, "device_s": t_dev, "device_objective_evals": dev.n_objective_evals,
           "sklearn_s": t_ref, "theta_device": dev.kernel_.theta.tolist(), "theta_sklearn": ref.kernel_.theta.tolist(),
           "lml_device": dev.log_marginal_likelihood_value_, "lml_sklearn": float(ref.log_marginal_likelihood_value_)}
    out.append(row)
    print(json.dumps(row), flush=True)
    json.dump(out, open("gpurun_out/gpfit.json", "w"), indent=1)


if __name__ == "__main__":
    main()
