"""Dev tool: time the GP (config 4) and actor-in-loop (config 5) paths with CUDA events."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mr_rl_b200 import DeviceGP, VecMREnv, init_actor, pack_actor


def ev():
    return torch.cuda.Event(enable_timing=True)


def gp_bench(n_train=2000, n_q=262144):
    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(-np.pi, np.pi, n_train))
    yx = 0.2 + 0.5 * np.cos(X + 0.3) + 0.09 * rng.standard_normal(n_train)
    t0 = time.perf_counter()
    d = DeviceGP.fit(X, yx, 0.2, 0.008, device="cuda:0")
    torch.cuda.synchronize()
    print(f"gp: device fit {time.perf_counter()-t0:.2f} s incl. first-call set-up (n_train={n_train}, n_pad={d.n_pad})")
    q = torch.rand(n_q, device="cuda:0", dtype=torch.float64) * 2 * np.pi - np.pi
    for want_std in (False, True):
        d.predict(q, want_std)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        reps = 5 if not want_std else 2
        e0.record()
        for _ in range(reps):
            d.predict(q, want_std)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pairs = n_q * n_train
        if want_std:
            flops = n_q * float(d.n_pad) ** 2      # triangular: 2 * n^2 / 2 per query
            print(f"gp mean+std: {ms:8.2f} ms per GP  ({n_q/ms/1e3:.1f} Mquery/s, variance contraction {flops/ms/1e9:.2f} TFLOP/s fp64)")
        else:
            print(f"gp mean    : {ms:8.3f} ms per GP  ({pairs/ms/1e6:.1f} Gpair/s)")
    rows = d.enable_spectral_variance()
    d.predict(q, True); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(3):
        d.predict(q, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"gp mean+std, spectral variance ({rows} projection rows): {ms:8.2f} ms per GP  ({n_q/ms/1e3:.1f} Mquery/s, "
          f"{2.0 * n_q * d.n_pad * max(rows, 1) / ms / 1e9:.2f} TFLOP/s fp64; fused kernel, K_q stays on chip; MR_GP_FUSED=0 for the two-kernel path)")
    # spot parity
    qs = q[:512].cpu().numpy().reshape(-1, 1)
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    sk = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None).fit(X.reshape(-1, 1), yx)
    mo_m, mo_s = sk.predict(qs, return_std=True)
    g_m, g_s = d.predict(q[:512], True)
    print("gp parity: mean rel", float(np.max(np.abs(g_m.cpu().numpy() - mo_m) / np.maximum(np.abs(mo_m), 1e-12))),
          "std rel", float(np.max(np.abs(g_s.cpu().numpy() - mo_s) / mo_s)))


def actor_bench(n=1 << 20, K=64):
    packed = pack_actor(init_actor(0), "cuda:0")
    for sigma in (0.0, 1.0):
        env = VecMREnv(n, device="cuda:0", noise="philox" if sigma else "none", seed=3, auto_reset=True)
        env.reset(init=None, noise_var=sigma, a0=1.0)
        env.rollout(policy=packed, k_steps=K)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        env.rollout(policy=packed, k_steps=K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"actor-in-loop rollout K={K} sigma={sigma}: {ms:8.2f} ms  {n*K/ms/1e6:8.2f} Genv-steps/s  "
              f"({n*K*2*4544/ms/1e9:.1f} TFLOP/s fp32 in the MLP)")


def heading_bench(n=262144, n_train=2000):
    """LearningModule.predict for a batch: device bounded minimiser + posterior at the minimiser."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200 import LearningModule
    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(-np.pi, np.pi, n_train)).reshape(-1, 1)
    yx = 0.2 + 0.5 * np.cos(X[:, 0] + 0.3) + 0.09 * rng.standard_normal(n_train)
    yy = -0.1 + 0.4 * np.sin(X[:, 0] - 0.2) + 0.09 * rng.standard_normal(n_train)
    gX = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None).fit(X, yx)
    gY = GaussianProcessRegressor(kernel=RBF(0.25) + WhiteKernel(0.008), optimizer=None).fit(X, yy)
    lm = LearningModule(device="cuda:0")
    lm.set_models(gX, gY, 1.5, 4.0, 0.2, -0.1)
    ang = torch.rand(n, device="cuda:0", dtype=torch.float64) * 6.28 - 3.14
    vd = 6.0 * torch.stack([torch.cos(ang), torch.sin(ang)], 1)
    import ctypes as C
    from mr_rl_b200 import _lib as L
    alpha = torch.empty(n, dtype=torch.float64, device="cuda:0"); nfev = torch.empty(n, dtype=torch.int32, device="cuda:0")
    def run():
        return L.load().mr_gp_correct_heading(C.byref(lm._dx._c), C.byref(lm._dy._c), vd.data_ptr(), n, 1.5, 4.0, 0.2, -0.1,
                                              alpha.data_ptr(), nfev.data_ptr(), None)
    run(); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    evals = float(nfev.sum())
    if lm._cheb is not None:
        a2 = torch.empty_like(alpha); nf2 = torch.empty_like(nfev)
        cheb = lm._cheb
        def run2():
            return L.load().mr_gp_correct_heading_cheb(cheb[0].data_ptr(), cheb[1].data_ptr(), cheb.shape[1], vd.data_ptr(), n,
                                                       1.5, 4.0, 0.2, -0.1, a2.data_ptr(), nf2.data_ptr(), None)
        run2(); torch.cuda.synchronize()
        e2, e3 = ev(), ev()
        e2.record(); run2(); e3.record(); torch.cuda.synchronize()
        d = (a2 - alpha).abs()
        print(f"heading correction on Chebyshev interpolants ({cheb.shape[1]} coefficients per GP): {e2.elapsed_time(e3):8.3f} ms for {n} "
              f"velocities; vs direct sums: max |d alpha| {float(d.max()):.2e}, same nfev in {float((nf2 == nfev).double().mean()) * 100:.2f} %")
    print(f"heading correction (bounded minimiser, 2 GPs in the loop): {ms:8.2f} ms for {n} velocities "
          f"({n/ms/1e3:.2f} M/s, mean nfev {evals/n:.1f}, {evals*n_train*2/ms/1e6:.1f} G kernel evals/s)")
    # CPU reference cost for scale: sklearn + scipy on one velocity
    import time
    from scipy.optimize import minimize_scalar
    t0 = time.perf_counter()
    for i in range(5):
        v = vd[i].cpu().numpy()
        def obj(a):                                   # Learning_module.py:10-24 with sklearn's predict, as the reference runs it
            mx, my = gX.predict(np.array([[a]]))[0], gY.predict(np.array([[a]]))[0]
            return (1.5 * 4.0) ** 2 + (mx + 0.2 - v[0]) ** 2 + 2 * 1.5 * 4.0 * np.cos(a) * (mx + 0.2 - v[0]) \
                + (my - 0.1 - v[1]) ** 2 + 2 * 1.5 * 4.0 * np.sin(a) * (my - 0.1 - v[1])
        minimize_scalar(obj, method="Bounded", bounds=[-np.pi, np.pi])
    print(f"  host (sklearn predict in the objective + scipy minimiser): {(time.perf_counter()-t0)/5*1e3:.2f} ms per velocity")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "gp"):
        gp_bench()
    if what in ("all", "actor"):
        actor_bench()
    if what in ("all", "heading"):
        heading_bench()
