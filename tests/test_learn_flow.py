"""The main.py flow (idle run -> estimateDisturbance, circle run -> learn) through this repo's
LearningModule: host preprocessing parity against the live reference (CPU), and the whole flow on the GPU."""
import numpy as np
import pytest

from conftest import Golden, rel_err


@pytest.fixture(scope="module")
def golden_learn():
    return Golden("learn.npz")


def test_learn_preprocessing_matches_reference_on_cpu(golden_learn):
    """estimateDisturbance / learn up to the GPR fit are pure host numpy: filtered velocities, drift, a0 and
    the GP training targets must equal the reference's (Learning_module.py:46-59,63-120)."""
    from mr_rl_b200.learning_module import LearningModule
    g = golden_learn
    lm = LearningModule.__new__(LearningModule)          # no device needed for the preprocessing
    lm.Dx = lm.Dy = 0
    lm.estimateDisturbance(g["px_idle"], g["py_idle"], g["t_idle"])
    assert rel_err(lm.Dx, g["Dx"]) < 1e-12 and rel_err(lm.Dy, g["Dy"]) < 1e-12
    N, px, py, vx, vy = lm._velocities(g["px"].astype(float), g["py"].astype(float), g["time"] - g["time"][0])
    freq = g["circ"][0, 0]
    speed = np.sqrt((vx - lm.Dx) ** 2 + (vy - lm.Dy) ** 2)[N:-N]
    a0 = np.median(speed / freq)
    assert rel_err(a0, g["a0"]) < 1e-12
    al = g["alpha"][N:-N]
    assert rel_err(vx[N:-N] - a0 * freq * np.cos(al), g["Yx"]) < 1e-10
    assert rel_err(al.reshape(-1, 1), g["X"]) < 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("fit", ["host", "device"])
def test_main_flow_on_the_device(golden_learn, fit):
    """run_sim (fused rollout, parity noise) -> learn (sklearn fit on the host, or the device fit) -> batched
    corrected headings."""
    import torch

    from mr_rl_b200 import LearningModule, run_sim
    g = golden_learn
    sigma, a0_def, freq = g["params"]
    px_i, py_i, _, t_i, _ = run_sim(g["idle"], init_pos=np.array([0, 0]), noise_var=sigma, a0=a0_def, is_mismatched=True,
                                    device="cuda:0", noise="table", noise_table=g["z1"][:, None])
    px, py, al, tm, _ = run_sim(g["circ"], init_pos=np.array([0, 0]), noise_var=sigma, a0=a0_def, is_mismatched=True,
                                device="cuda:0", noise="table", noise_table=g["z2"][:, None])
    assert rel_err(px_i, g["px_idle"]) < 1e-9 and rel_err(py, g["py"]) < 1e-9      # trajectories = the reference's
    lm = LearningModule(device="cuda:0", fit=fit)
    lm.gprX.n_restarts_optimizer = 0
    lm.gprY.n_restarts_optimizer = 0
    lm.estimateDisturbance(px_i, py_i, t_i)
    a0 = lm.learn(px, py, al, tm.copy(), g["circ"])
    assert rel_err(a0, g["a0"]) < 1e-9 and rel_err(lm.Yx, g["Yx"]) < 1e-7
    # device GP == the sklearn model it was uploaded from / == sklearn's own fit of the same data
    q = np.linspace(-3, 3, 101)
    ref_gpr = lm.gprX
    if fit == "device":
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import RBF, WhiteKernel
        ref_gpr = GaussianProcessRegressor(kernel=RBF(1.0, (1e-2, 10.0)) + WhiteKernel()).fit(lm.X, lm.Yx)
        assert np.allclose(lm.gprX.kernel_.theta, ref_gpr.kernel_.theta, rtol=1e-4, atol=1e-4)
        assert abs(lm.gprX.log_marginal_likelihood_value_ - ref_gpr.log_marginal_likelihood_value_) < 1e-6 * abs(ref_gpr.log_marginal_likelihood_value_)
    m_ref, s_ref = ref_gpr.predict(q.reshape(-1, 1), return_std=True)
    m, _, s, _ = lm.gp_batch(torch.as_tensor(q, device="cuda:0"), True)
    tol = 1.0 if fit == "host" else 1e3          # the device fit stops at its own (equally converged) optimiser iterate
    assert np.allclose(m.cpu().numpy(), m_ref, rtol=1e-7 * tol, atol=1e-9 * tol)
    assert np.allclose(s.cpu().numpy(), s_ref, rtol=1e-5 * tol, atol=1e-8 * tol)
    # corrected headings for a batch of desired velocities (main.py:145-155)
    ang = np.linspace(0, np.pi / 2, 64)
    vd = a0 * freq * np.stack([np.cos(ang), np.sin(ang)], 1)
    alpha, mx, my, sx, sy = lm.predict_batch(vd)
    assert alpha.shape == (64,) and torch.isfinite(alpha).all() and float(sx.min()) > 0
    vxp, vyp = lm.velocity_model(torch.full((64,), float(freq), device="cuda:0"), alpha)
    # the corrected heading brings the predicted velocity closer to the desired one than the naive heading
    vxn, vyn = lm.velocity_model(torch.full((64,), float(freq), device="cuda:0"), torch.as_tensor(ang, device="cuda:0"))
    vd_t = torch.as_tensor(vd, device="cuda:0")
    err_c = torch.hypot(vxp + lm.Dx - vd_t[:, 0], vyp + lm.Dy - vd_t[:, 1]).mean()
    err_n = torch.hypot(vxn + lm.Dx - vd_t[:, 0], vyn + lm.Dy - vd_t[:, 1]).mean()
    assert float(err_c) <= float(err_n) + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("cut", [None, 0.7])
def test_device_preprocessing_matches_the_host_restatement(golden_learn, cut):
    """mr_learn_preprocess (filters, np.gradient, drift means, median a0, GP targets) against the numpy/scipy
    restatement that the CPU test pins to the live reference (Learning_module.py:46-59, :63-120)."""
    from mr_rl_b200.learning_module import LearningModule
    g = golden_learn
    host = LearningModule.__new__(LearningModule)
    host.Dx = host.Dy = 0
    host.preprocess = "host"
    dev = LearningModule(device="cuda:0", fit="device")
    assert dev.preprocess == "device"
    host.estimateDisturbance(g["px_idle"], g["py_idle"], g["t_idle"])
    dev.estimateDisturbance(g["px_idle"], g["py_idle"], g["t_idle"])
    assert rel_err(dev.Dx, host.Dx) < 1e-10 and rel_err(dev.Dy, host.Dy) < 1e-10
    assert rel_err(dev.Dx, g["Dx"]) < 1e-10
    # the learn() half, optionally with controller-off frames (alpha >= 500) that truncate the record
    actions = np.array(g["circ"], dtype=float)
    time = np.asarray(g["time"], float)
    if cut is not None:
        cut = int(cut * len(time))
        actions[cut:, 1] = 1000.0
    N, _, _, vx, vy = host._velocities(g["px"].astype(float), g["py"].astype(float), time - time[0])
    n_valid = len(time) if cut is None else cut - 1
    al = np.asarray(g["alpha"], float)[:n_valid][N:-N]
    freq = actions[0, 0]
    speed = np.sqrt((vx - host.Dx) ** 2 + (vy - host.Dy) ** 2)[:n_valid][N:-N]
    a0_ref = np.median(speed / freq)
    _, dvx, dvy, X, Yx, Yy, scal = dev._device_preprocess(g["px"], g["py"], time, g["alpha"], freq, None if cut is None else n_valid,
                                                          subtract_t0=True)
    assert rel_err(dvx.cpu().numpy(), vx) < 1e-10 and rel_err(dvy.cpu().numpy(), vy) < 1e-10
    assert rel_err(float(scal[2]), a0_ref) < 1e-10 and int(scal[3]) == len(al)
    assert np.array_equal(X.cpu().numpy(), al)
    assert rel_err(Yx.cpu().numpy(), vx[:n_valid][N:-N] - a0_ref * freq * np.cos(al)) < 1e-9
    assert rel_err(Yy.cpu().numpy(), vy[:n_valid][N:-N] - a0_ref * freq * np.sin(al)) < 1e-9
    if cut is None:
        assert rel_err(float(scal[2]), g["a0"]) < 1e-9 and rel_err(Yx.cpu().numpy(), g["Yx"]) < 1e-7


@pytest.mark.gpu
def test_learn_preprocess_argument_errors():
    import torch
    from mr_rl_b200 import _lib as L
    lib = L.load()
    d = torch.zeros(4096, dtype=torch.float64, device="cuda")
    p = d.data_ptr()
    assert lib.mr_learn_preprocess(None, p, p, 100, 14, 0, 0.0, 0.0, None, 1.0, 100, p, p, None, None, None, p, p, 1 << 20, None) != 0
    assert lib.mr_learn_preprocess(p, p, p, 20, 14, 0, 0.0, 0.0, None, 1.0, 20, p, p, None, None, None, p, p, 1 << 20, None) != 0
    assert b"filter" in lib.mr_last_error()
    assert lib.mr_learn_preprocess(p, p, p, 100, 14, 0, 0.0, 0.0, None, 1.0, 100, p, p, None, None, None, p, p, 8, None) != 0
    assert b"workspace" in lib.mr_last_error()
    assert lib.mr_learn_preprocess(p, p, p, 100, 14, 0, 0.0, 0.0, p, 1.0, 20, p, p, p, p, p, p, p, 1 << 20, None) != 0
    assert b"n_valid" in lib.mr_last_error()


@pytest.mark.gpu
def test_device_learn_reproduces_the_live_reference_fit(golden_learn):
    """LearningModule(fit="device").learn against the live reference's learn() with its own settings (5 restarts of
    sklearn's optimiser, restart points from numpy's global RandomState seeded identically): fitted hyper-parameters,
    log marginal likelihoods, the posterior on a grid and predict() (Learning_module.py:28-33,122-123,198-224;
    golden made by oracle/gen_golden.py:gen_learn_fit)."""
    import torch
    from mr_rl_b200 import LearningModule
    g, f = golden_learn, Golden("learn_fit.npz")
    lm = LearningModule(device="cuda:0", fit="device")
    lm.estimateDisturbance(g["px_idle"], g["py_idle"], g["t_idle"])
    np.random.seed(int(f["seed"]))
    a0 = lm.learn(g["px"].copy(), g["py"].copy(), g["alpha"].copy(), g["time"].copy(), g["circ"])
    assert rel_err(a0, f["a0"]) < 1e-9
    # the optimum: same basin, converged to the optimiser's own tolerance (noise_level sits near its lower bound here,
    # K is conditioned ~1e5, so theta agrees to ~1e-3 and the likelihood to ~1e-7)
    assert np.allclose(lm.gprX.kernel_.theta, f["theta_x"], atol=5e-3) and np.allclose(lm.gprY.kernel_.theta, f["theta_y"], atol=5e-3)
    assert abs(lm.gprX.log_marginal_likelihood_value_ - f["lml_x"]) < 1e-6 * abs(f["lml_x"])
    assert abs(lm.gprY.log_marginal_likelihood_value_ - f["lml_y"]) < 1e-6 * abs(f["lml_y"])
    q = torch.as_tensor(f["grid"], device="cuda:0")
    mx, my, sx, sy = lm.gp_batch(q, True)
    scale = np.abs(f["grid_mx"]).max()
    assert np.abs(mx.cpu().numpy() - f["grid_mx"]).max() < 1e-3 * scale and np.abs(my.cpu().numpy() - f["grid_my"]).max() < 1e-3 * scale
    inside = np.abs(f["grid"]) < 2.8                     # away from the edge of the data the std is well determined
    assert np.allclose(sx.cpu().numpy()[inside], f["grid_sx"][inside], rtol=2e-2, atol=1e-5)
    alpha, pmx, pmy, psx, psy = lm.predict_batch(f["vd"])
    ref = f["predict"]
    d_alpha = np.angle(np.exp(1j * (alpha.cpu().numpy() - ref[:, 0])))
    assert np.abs(d_alpha).max() < 2e-3
    assert np.abs(pmx.cpu().numpy() - ref[:, 1]).max() < 2e-3 * scale and np.abs(pmy.cpu().numpy() - ref[:, 2]).max() < 2e-3 * scale
