"""Device GP fit at fixed hyper-parameters (csrc/mr_gpfit.cu) against sklearn's GaussianProcessRegressor with
optimizer=None — the factorisation half of Learning_module.py:122-123."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sk_fit(X, y, ls, noise):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    gpr = GaussianProcessRegressor(kernel=RBF(ls) + WhiteKernel(noise), optimizer=None)
    gpr.fit(X, y)
    return gpr


@pytest.mark.parametrize("n,dim,ls,noise", [(100, 1, 0.7, 0.05), (128, 1, 1.3, 0.01), (300, 1, 0.5, 0.1),
                                             (1000, 2, 0.8, 0.02), (2000, 1, 0.3, 1e-3)])
def test_device_fit_matches_sklearn_fixed_theta(n, dim, ls, noise):
    from scipy.linalg import solve_triangular
    from mr_rl_b200.gp import DeviceGP
    rng = np.random.default_rng(n + dim)
    X = rng.uniform(-np.pi, np.pi, size=(n, dim))
    y = np.sin(X).sum(axis=1) + 0.1 * rng.standard_normal(n)
    ref = sk_fit(X, y, ls, noise)
    gp = DeviceGP.fit(X, y, ls, noise, eval_gradient=True)
    n_pad = gp.n_pad
    alpha = gp._alpha.cpu().numpy()
    W = gp._linv.cpu().numpy()
    # the conditioning of K (1/noise) bounds what any factorisation can agree on
    tol = 1e-11 / noise
    assert np.allclose(alpha[:n], ref.alpha_, rtol=tol, atol=tol * np.abs(ref.alpha_).max())
    assert np.all(alpha[n:] == 0.0)
    Linv = solve_triangular(ref.L_, np.eye(n), lower=True)
    assert np.allclose(W[:n, :n], Linv, rtol=0, atol=tol * np.abs(Linv).max())
    assert np.all(W[n:, :] == 0.0) and np.all(W[:, n:] == 0.0)
    assert np.all(np.triu(W, 1) == 0.0)
    lml = ref.log_marginal_likelihood(ref.kernel_.theta)
    assert abs(gp.log_marginal_likelihood_value_ - lml) <= 1e-9 * max(1.0, abs(lml))
    _, grad = ref.log_marginal_likelihood(ref.kernel_.theta, eval_gradient=True)
    assert np.allclose(gp.log_marginal_likelihood_gradient_, grad, rtol=1e-7, atol=1e-7 * np.abs(grad).max())
    assert np.allclose(gp._xs.cpu().numpy()[:n], X / ls, rtol=1e-15, atol=0)
    # and the fitted model predicts like sklearn's
    q = rng.uniform(-np.pi, np.pi, size=(513, dim))
    m, s = gp.predict(q, return_std=True)
    mr, sr = ref.predict(q, return_std=True)
    assert np.allclose(m.cpu().numpy(), mr, rtol=1e-8, atol=1e-8)
    assert np.allclose(s.cpu().numpy(), sr, rtol=1e-6, atol=1e-8)
    assert n_pad % 128 == 0


def test_device_fit_reports_non_positive_definite():
    from mr_rl_b200.gp import DeviceGP
    X = np.zeros((64, 1))                       # identical inputs, no noise, negative jitter -> singular
    y = np.ones(64)
    with pytest.raises(np.linalg.LinAlgError):
        DeviceGP.fit(X, y, 1.0, 0.0, jitter=-1e-3)


def test_device_fit_argument_errors():
    from mr_rl_b200 import _lib as L
    lib = L.load()
    rc = lib.mr_gp_fit(None, None, 10, 128, 1, 1.0, 0.1, 1e-10, None, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"null" in lib.mr_last_error()
    d = torch.zeros(128 * 130, dtype=torch.float64, device="cuda")
    p = d.data_ptr()
    rc = lib.mr_gp_fit(p, p, 10, 100, 1, 1.0, 0.1, 1e-10, p, p, p, None, None, None, p, 1 << 30, None)
    assert rc != 0 and b"multiple" in lib.mr_last_error()
    rc = lib.mr_gp_fit(p, p, 10, 128, 1, 1.0, 0.1, 1e-10, p, p, p, None, None, None, p, 16, None)
    assert rc != 0 and b"workspace" in lib.mr_last_error()


@pytest.mark.parametrize("restarts", [0, 3])
def test_device_gpr_hyperparameter_search_matches_sklearn(restarts):
    """gpr.DeviceGPR.fit = sklearn's search (L-BFGS-B from the initial theta + restarts drawn from the estimator's
    RandomState), each objective evaluation on the device: same optimum, same predictions, same r^2
    (Learning_module.py:28-33,122-126)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200.gpr import DeviceGPR
    rng = np.random.default_rng(7)
    n = 400
    X = np.sort(rng.uniform(-np.pi, np.pi, size=(n, 1)), axis=0)
    y = 0.8 * np.sin(2 * X[:, 0]) + 0.3 * np.cos(X[:, 0]) + 0.15 * rng.standard_normal(n)
    ref = GaussianProcessRegressor(kernel=RBF(1.0, (1e-2, 10.0)) + WhiteKernel(), n_restarts_optimizer=restarts,
                                   random_state=11).fit(X, y)
    dev = DeviceGPR(n_restarts_optimizer=restarts, random_state=11).fit(X, y)
    assert np.allclose(dev.kernel_.theta, ref.kernel_.theta, rtol=1e-4, atol=1e-4)
    assert abs(dev.log_marginal_likelihood_value_ - ref.log_marginal_likelihood_value_) < 1e-7 * abs(ref.log_marginal_likelihood_value_)
    q = np.linspace(-3, 3, 257).reshape(-1, 1)
    m, s = dev.predict(q, return_std=True)
    mr, sr = ref.predict(q, return_std=True)
    assert np.allclose(m, mr, rtol=1e-4, atol=1e-5) and np.allclose(s, sr, rtol=1e-4, atol=1e-6)
    assert abs(dev.score(X, y) - ref.score(X, y)) < 1e-6
    lml, grad = dev.log_marginal_likelihood(ref.kernel_.theta, eval_gradient=True)
    lml_r, grad_r = ref.log_marginal_likelihood(ref.kernel_.theta, eval_gradient=True)
    assert abs(lml - lml_r) < 1e-9 * abs(lml_r) and np.allclose(grad, grad_r, rtol=1e-6, atol=1e-6)


def test_diagonal_block_kernels_are_bit_identical(tmp_path):
    """The default diagonal-block kernel (chol_diag2_kernel: split arrive / wait per step, look-ahead pivot reciprocal)
    does the arithmetic of chol_diag_kernel (MR_CHOL_DIAG=1) entry by entry in the same order: L^-1, alpha and the log
    marginal likelihood must agree bit for bit.  The switch is read once per process, hence the two subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from mr_rl_b200.gp import DeviceGP\n"
        "out = {}\n"
        "for n, ls, noise in ((300, 0.5, 0.1), (1000, 0.3, 1e-3)):\n"
        "    rng = np.random.default_rng(n)\n"
        "    X = rng.uniform(-np.pi, np.pi, size=(n, 1))\n"
        "    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)\n"
        "    gp = DeviceGP.fit(X, y, ls, noise, eval_gradient=True)\n"
        "    out[f'W{n}'] = gp._linv.cpu().numpy(); out[f'a{n}'] = gp._alpha.cpu().numpy()\n"
        "    out[f'l{n}'] = np.array([gp.log_marginal_likelihood_value_, *gp.log_marginal_likelihood_gradient_])\n"
        "np.savez(sys.argv[1], **out)\n")
    res = {}
    for tag, val in (("v1", "1"), ("v2", "")):
        env = dict(os.environ)
        env.pop("MR_CHOL_DIAG", None)
        if val:
            env["MR_CHOL_DIAG"] = val
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        res[tag] = np.load(path)
    for key in res["v1"].files:
        assert np.array_equal(res["v1"][key], res["v2"][key]), key
