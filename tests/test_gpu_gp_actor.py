"""GPU parity: GP inference (Learning_module -> sklearn GPR.predict), the DDPG actor forward, and
the single-env façades (MR_Env, run_sim, LearningModule) against golden vectors / the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import mr_oracle as mo

pytestmark = pytest.mark.gpu


def fitted_pair(g):
    gx = mo.fit_fixed_gp(g["X"], g["yx"], float(g["lsx"]), float(g["noise"]))
    gy = mo.fit_fixed_gp(g["X"], g["yy"], float(g["lsy"]), float(g["noise"]))
    return gx, gy


def to_device(m):
    from mr_rl_b200 import DeviceGP
    return DeviceGP(m.X_train, m.alpha, m.L, m.length_scale, m.noise_level, device="cuda:0")


def test_gp_predict_matches_reference_learning_module(golden_gp):
    """mean and std on the golden grid recorded from the reference's LearningModule / sklearn."""
    g = golden_gp
    gx, gy = fitted_pair(g)
    dx, dy = to_device(gx), to_device(gy)
    mx, sx = dx.predict(g["grid"], return_std=True)
    my, sy = dy.predict(g["grid"], return_std=True)
    assert np.allclose(mx.cpu().numpy(), g["grid_mx"], rtol=1e-8, atol=1e-10)
    assert np.allclose(my.cpu().numpy(), g["grid_my"], rtol=1e-8, atol=1e-10)
    assert np.allclose(sx.cpu().numpy(), g["grid_sx"], rtol=1e-7, atol=1e-10)
    assert np.allclose(sy.cpu().numpy(), g["grid_sy"], rtol=1e-7, atol=1e-10)
    # mean-only path (the objective's calls, Learning_module.py:17-18)
    assert np.array_equal(dx.predict(g["grid"]).cpu().numpy(), mx.cpu().numpy())
    # LearningModule.error() rows
    a = np.arctan2(g["vd"][:, 1], g["vd"][:, 0])
    emx, esx = dx.predict(a, True)
    emy, esy = dy.predict(a, True)
    got = np.stack([emx.cpu().numpy(), emy.cpu().numpy(), esx.cpu().numpy(), esy.cpu().numpy()], 1)
    assert np.allclose(got, g["error"], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("n_train,n_q,dim", [(1, 5, 1), (37, 1, 1), (300, 1000, 1), (513, 700, 2), (2000, 4096, 1)])
def test_gp_predict_matches_oracle_ragged_sizes(n_train, n_q, dim):
    """SURVEY C4 kernels (RBF(0.2)+White(0.008)) incl. ragged sizes, the 2-D input variant
    (Learning_module_2d.py) and the full N_train = 2000 of config 4."""
    rng = np.random.default_rng(n_train)
    X = np.sort(rng.uniform(-np.pi, np.pi, (n_train, dim)), axis=0)
    y = 0.2 + 0.5 * np.cos(X[:, 0] + 0.3) + 0.09 * rng.standard_normal(n_train)
    m = mo.fit_fixed_gp(X, y, 0.2 if dim == 1 else 0.6, 0.008)
    q = rng.uniform(-np.pi, np.pi, (n_q, dim))
    mean_o, std_o = mo.gp_predict(m, q)
    d = to_device(m)
    mean, std = d.predict(q, return_std=True)
    assert np.allclose(mean.cpu().numpy(), mean_o, rtol=1e-8, atol=1e-9)
    assert np.allclose(std.cpu().numpy(), std_o, rtol=1e-6, atol=1e-9)     # variance cancellation: 1+noise - |V|^2


def test_gp_matches_sklearn_directly():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200 import DeviceGP
    rng = np.random.default_rng(3)
    X = np.sort(rng.uniform(-np.pi, np.pi, 400)).reshape(-1, 1)
    y = -0.1 + 0.4 * np.sin(X[:, 0] - 0.2) + 0.09 * rng.standard_normal(400)
    gpr = GaussianProcessRegressor(kernel=RBF(0.25) + WhiteKernel(0.008), optimizer=None).fit(X, y)
    q = rng.uniform(-np.pi, np.pi, (999, 1))
    m_ref, s_ref = gpr.predict(q, return_std=True)
    d = DeviceGP.from_sklearn(gpr, device="cuda:0")
    m, s = d.predict(q, True)
    assert np.allclose(m.cpu().numpy(), m_ref, rtol=1e-8, atol=1e-10)
    assert np.allclose(s.cpu().numpy(), s_ref, rtol=1e-6, atol=1e-9)


def test_spectral_variance_matches_triangular_and_sklearn():
    """DeviceGP.enable_spectral_variance: |P k|^2 with the leading eigenpairs of K instead of |L^-1 k|^2 — same posterior
    std as sklearn (and as the triangular form) at a fraction of the rows; refused when the kernel is not low rank."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200 import DeviceGP
    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(-np.pi, np.pi, 2000)).reshape(-1, 1)
    y = 0.2 + 0.5 * np.cos(X[:, 0] + 0.3) + 0.09 * rng.standard_normal(2000)
    sk = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None).fit(X, y)
    gp = DeviceGP.from_sklearn(sk, "cuda:0")
    q = rng.uniform(-np.pi, np.pi, 3000)
    m0, s0 = gp.predict(q, True)
    rows = gp.enable_spectral_variance()
    assert 0 < rows <= 256 and rows % 32 == 0                    # ~80 significant eigenvalues at l = 0.2 on [-pi, pi]
    m1, s1 = gp.predict(q, True)
    mr, sr = sk.predict(q.reshape(-1, 1), return_std=True)
    assert float((m1 - m0).abs().max()) < 1e-11                  # fused kernel: another summation order (|mean| ~ 0.5)
    assert float(((s1 - s0).abs() / s0).max()) < 1e-9
    assert np.allclose(s1.cpu().numpy(), sr, rtol=1e-6, atol=1e-9) and np.allclose(m1.cpu().numpy(), mr, rtol=1e-8, atol=1e-8)
    # outside the training range too (the std grows to the prior there)
    far = np.linspace(-6, 6, 257)
    _, sf = gp.predict(far, True)
    assert np.allclose(sf.cpu().numpy(), sk.predict(far.reshape(-1, 1), return_std=True)[1], rtol=1e-6, atol=1e-9)
    # a rough kernel is not low rank: the projection is refused and the triangular form stays
    sk2 = GaussianProcessRegressor(kernel=RBF(0.003) + WhiteKernel(0.008), optimizer=None).fit(X, y)
    gp2 = DeviceGP.from_sklearn(sk2, "cuda:0")
    assert gp2.enable_spectral_variance() == 0
    _, s2 = gp2.predict(q, True)
    assert np.allclose(s2.cpu().numpy(), sk2.predict(q.reshape(-1, 1), return_std=True)[1], rtol=1e-6, atol=1e-9)


def test_learning_module_facade_error_and_predict(golden_gp):
    """Reference surface: error(vd) -> four (1,) arrays; predict(vd) -> (alpha, muX, muY, sigX, sigY)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200 import LearningModule
    g = golden_gp
    X = g["X"].reshape(-1, 1)
    gX = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None).fit(X, g["yx"])
    gY = GaussianProcessRegressor(kernel=RBF(0.25) + WhiteKernel(0.008), optimizer=None).fit(X, g["yy"])
    a0, freq, Dx, Dy = g["hyper"]
    lm = LearningModule(device="cuda:0")
    lm.set_models(gX, gY, a0, freq, Dx, Dy)
    for i in range(4):
        out = lm.error(g["vd"][i])
        assert all(o.shape == (1,) for o in out)
        assert np.allclose(np.concatenate(out), g["error"][i], rtol=1e-7, atol=1e-10)
    obj = [float(np.ravel(lm._objective(a, g["vd"][i]))[0]) for i, a in enumerate(g["obj_alpha"][:8])]
    assert np.allclose(obj, g["objective"][:8], rtol=1e-8, atol=1e-9)
    # predict: compare with the oracle objective minimised by the same scipy routine
    from scipy.optimize import minimize_scalar
    gx, gy = fitted_pair(g)
    vd = g["vd"][0]
    ref = minimize_scalar(lambda a: float(np.ravel(mo.lm_objective(a, a0, freq, vd, gx, gy, Dx, Dy))[0]),
                          method="Bounded", bounds=[-np.pi, np.pi])
    A, muX, muY, sigX, sigY = lm.predict(vd)
    assert abs(float(A) - ref.x) < 1e-6
    m_ref, s_ref = mo.gp_predict(gx, np.array([[ref.x]]))
    assert np.allclose(muX, m_ref, rtol=1e-6, atol=1e-8) and np.allclose(sigX, s_ref, rtol=1e-5, atol=1e-8)
    mx, my, sx, sy = lm.error_batch(torch.as_tensor(g["vd"], device="cuda:0"))
    assert np.allclose(mx.cpu().numpy(), g["error"][:, 0], rtol=1e-7, atol=1e-10)


def test_actor_forward_matches_torch_fp32_reference():
    from mr_rl_b200 import actor_forward, init_actor, pack_actor
    from mr_rl_b200.actor import torch_reference
    params = init_actor(0)
    # make BN non-trivial so the inference form is really exercised
    g = torch.Generator().manual_seed(1)
    for k in ("m1", "m2"):
        params[k] = 0.05 * torch.randn(64, generator=g)
    for k in ("v1", "v2"):
        params[k] = 0.5 + torch.rand(64, generator=g)
    for k in ("be1", "be2", "b1", "b2"):
        params[k] = 0.01 * torch.randn(64, generator=g)
    params["w3"] = 0.3 * torch.randn(64, 2, generator=g)
    packed = pack_actor(params, "cuda:0")
    n = 5000
    rng = np.random.default_rng(0)
    obs = np.zeros((n, 5))
    obs[:, :2] = rng.uniform(-200, 200, (n, 2))
    obs[:, 4] = np.hypot(obs[:, 0], obs[:, 1])
    ref = torch_reference(params, obs).numpy()
    for dt in (torch.float64, torch.float32):
        soa = torch.as_tensor(obs.T.copy(), dtype=dt, device="cuda:0")
        act = actor_forward(packed, soa).cpu().numpy()
        assert act.shape == (n, 2)
        assert np.allclose(act, ref, rtol=1e-4, atol=1e-4)        # fp32 kernel vs torch fp32
    assert np.allclose(ref, mo.actor_forward({k: v.numpy() for k, v in params.items()}, obs), rtol=1e-4, atol=1e-4)


def test_tensor_core_actor_matches_fp32_actor_actions():
    """One fused step with the tcgen05 actor vs the standalone fp32 actor kernel: the ACTIONS themselves
    (recovered from the displacement is indirect, so compare through a noise-free, a0 = 1 step from rest)."""
    from mr_rl_b200 import VecMREnv, actor_forward, init_actor, pack_actor
    from mr_rl_b200.actor import torch_reference
    params = init_actor(3)
    g = torch.Generator().manual_seed(4)
    params["w3"] = 0.4 * torch.randn(64, 2, generator=g)
    params["m2"] = 0.05 * torch.randn(64, generator=g); params["v2"] = 0.5 + torch.rand(64, generator=g)
    params["b2"] = 0.02 * torch.randn(64, generator=g); params["be1"] = 0.02 * torch.randn(64, generator=g)
    packed = pack_actor(params, "cuda:0")
    n = 1000                                           # not a multiple of 128: padding lanes take part in the MMA
    env = VecMREnv(n, device="cuda:0", noise="none")
    env.reset(init=None, noise_var=0.0, a0=1.0)
    obs0 = env.obs.cpu().numpy().copy()
    ref = torch_reference(params, obs0).numpy().astype(np.float64)          # [n, 2] fp32 torch forward
    res = env.rollout(policy=packed, k_steps=1, record=True)
    # noise-free first step from reset: the integrator restarts at h = 1e-6 with K0 = 0 (zero action) and
    # grows, so the displacement is dt * f * (cos a, sin a) up to ~1e-6 relative -> it exposes the actions
    disp = env.last_pos.cpu().numpy() - obs0[:, :2]
    pred = 0.03 * ref[:, :1] * np.stack([np.cos(ref[:, 1]), np.sin(ref[:, 1])], 1)
    assert np.allclose(disp, pred, rtol=2e-4, atol=1e-7)
    f_got = np.hypot(disp[:, 0], disp[:, 1]) / 0.03
    assert np.allclose(f_got, np.abs(ref[:, 0]), rtol=1e-4, atol=1e-6)          # the 1e-4 bar on the action magnitude


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_standalone_tensor_core_actor_matches_fp32_reference(dt):
    """mr_actor_forward_env (persistent CTAs, hidden layer on tcgen05) against the torch fp32 forward and the CUDA-core
    kernel on MR_Env observations, ragged n (several tiles per CTA and a partial last tile)."""
    from mr_rl_b200 import VecMREnv, actor_forward, init_actor, pack_actor
    from mr_rl_b200.actor import torch_reference
    params = init_actor(5)
    g = torch.Generator().manual_seed(6)
    params["w3"] = 0.4 * torch.randn(64, 2, generator=g)
    params["m1"] = 0.05 * torch.randn(64, generator=g); params["v1"] = 0.5 + torch.rand(64, generator=g)
    params["b2"] = 0.02 * torch.randn(64, generator=g); params["be2"] = 0.02 * torch.randn(64, generator=g)
    packed = pack_actor(params, "cuda:0")
    n = 128 * 700 + 37
    env = VecMREnv(n, device="cuda:0", dtype=dt, noise="none")
    env.reset(init=None, noise_var=0.0, a0=1.0)
    a_tc = actor_forward(packed, env._obs, n, env_obs=True)
    a_simt = actor_forward(packed, env._obs, n)
    ref = torch_reference(params, env.obs.cpu().numpy())
    hi = torch.tensor([20.0, 2 * np.pi])
    assert a_tc.shape == (n, 2) and a_tc.dtype == dt
    assert float(((a_tc.cpu().float() - ref).abs() / hi).max()) < 1e-4          # the 1e-4 bar, relative to the action range
    assert float(((a_tc - a_simt).abs().cpu().float() / hi).max()) < 1e-4


def test_actor_in_the_rollout_loop_matches_stepwise_composition():
    """Config-5 path at reduced size: fused rollout with the actor evaluated in-kernel ==
    actor_forward kernel + single-step kernel composed on the host (noise-free)."""
    from mr_rl_b200 import VecMREnv, actor_forward, init_actor, pack_actor
    params = init_actor(0)
    params["w3"] = 0.5 * torch.randn(64, 2, generator=torch.Generator().manual_seed(2))
    packed = pack_actor(params, "cuda:0")
    n, K = 777, 20
    e1 = VecMREnv(n, device="cuda:0", noise="none"); e2 = VecMREnv(n, device="cuda:0", noise="none")
    for e in (e1, e2):
        e.reset(init=None, noise_var=0.0, a0=1.0)
    e2._state.copy_(e1._state); e2._obs.copy_(e1._obs)
    res = e1.rollout(policy=packed, k_steps=K, record=True)
    xy = []
    for _ in range(K):
        a = actor_forward(packed, e2._obs, n)
        e2.step(a)
        xy.append(e2.last_pos.cpu().numpy().copy())
    xy = np.stack(xy)
    got = res["xy"].cpu().numpy().transpose(0, 2, 1)
    # the actor is an fp32 network (1e-4 on actions): in the fused kernel its hidden layer runs on the tensor
    # cores as 3xTF32, the standalone kernel uses fp32 FMAs -> actions agree to ~1e-6, positions accordingly
    assert rel_err(got, xy) < 1e-6
    assert rel_err(e1.last_pos.cpu().numpy(), e2.last_pos.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("path", ["default", "tf32", "simt"])
def test_in_rollout_actor_paths_cross_an_episode_boundary(path):
    """Every in-kernel actor implementation (3xFP16 tensor cores with both dense layers, 3xTF32 hidden layer, CUDA cores),
    with non-trivial batch-norm statistics, through AUTO RESETS: the fused rollout must equal actor_forward + single step
    composed on the host — in particular the first action of a new episode sees the new episode's observation (x, y AND
    distance), as a caller that does `s = env.reset()` would feed it (RL/MR_ddpg.py:270-278)."""
    from mr_rl_b200 import VecMREnv, _lib as L, actor_forward, init_actor, pack_actor
    params = init_actor(7)
    g = torch.Generator().manual_seed(8)
    params["w3"] = 0.5 * torch.randn(64, 2, generator=g)
    for k in ("m1", "m2"):
        params[k] = 0.05 * torch.randn(64, generator=g)
    for k in ("v1", "v2"):
        params[k] = 0.5 + torch.rand(64, generator=g)
    for k in ("be1", "be2", "b1", "b2"):
        params[k] = 0.02 * torch.randn(64, generator=g)
    packed = pack_actor(params, "cuda:0")
    n, K = 300, 14
    try:
        L.set_actor_path(path)
        e1 = VecMREnv(n, device="cuda:0", noise="philox", seed=3, auto_reset=True)
        e2 = VecMREnv(n, device="cuda:0", noise="philox", seed=3, auto_reset=True)
        for e in (e1, e2):
            e.max_timesteps = 5                                    # two episode boundaries inside K steps
            e.reset(init=None, noise_var=0.0, a0=1.0)
        res = e1.rollout(policy=packed, k_steps=K, record=True, record_done=True)
        xy, dn = [], []
        for _ in range(K):
            a = actor_forward(packed, e2._obs, n)
            _, _, d, _ = e2.step(a)
            # (with auto reset last_pos is already the new episode's start: compare the states after the whole run and the
            # done flags per step; positions per step only where no reset happened)
            xy.append(e2.last_pos.cpu().numpy().copy()); dn.append(d.cpu().numpy().copy())
    finally:
        L.set_actor_path("default")
    xy, dn = np.stack(xy), np.stack(dn).astype(bool)
    got = res["xy"].cpu().numpy().transpose(0, 2, 1)
    assert np.array_equal(res["done_traj"].cpu().numpy().astype(bool), dn) and dn.sum() >= 2 * n
    keep = ~dn
    assert rel_err(got[keep], xy[keep]) < 2e-6
    assert rel_err(e1.last_pos.cpu().numpy(), e2.last_pos.cpu().numpy()) < 2e-6
    assert np.array_equal(e1.counter.cpu().numpy(), e2.counter.cpu().numpy())


def test_mr_env_facade_matches_live_reference(golden_single):
    """The N = 1 drop-in class: numpy in / numpy out, python int reward, python bool done, dict info."""
    from mr_rl_b200 import MR_Env
    g = golden_single.case("c1_sigma1")
    env = MR_Env(device="cuda:0", noise="table", noise_table=g["z"][:, None])
    assert env.observation_space.shape[0] == 5 and env.action_space.shape[0] == 2
    assert np.allclose(env.action_space.high, [20, 2 * np.pi])
    obs = env.reset(init=np.asarray(g["init"]), noise_var=1, a0=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (5,) and obs.dtype == np.float64
    assert rel_err(obs, g["reset_obs"]) < 1e-9
    for k in range(60):
        obs, rew, done, info = env.step(g["actions"][k])
        assert rew == 10 and isinstance(done, bool) and info == {}
        assert rel_err(obs, g["obs"][k]) < 1e-9
        assert done == bool(g["done"][k]) and env.counter == int(g["counter"][k])
        assert rel_err(env.state_prime, g["state_prime"][k]) < 1e-9
        assert rel_err(np.asarray(env.last_pos), g["pos"][k]) < 1e-9


def test_mr_env_facade_reset_ordering_quirk(golden_single):
    from mr_rl_b200 import MR_Env
    g = golden_single.case("reset_after_mismatch")
    env = MR_Env(device="cuda:0", noise="table", noise_table=g["z"][:, None])
    env.simulator.is_mismatched = True              # left over from a previous mismatched episode
    env.reset(init=np.asarray(g["init"]), noise_var=1, a0=1, is_mismatched=False)
    assert int(env._vec._cursor[0]) == int(g["reset_cursor"]) == 6     # 2 evaluations x 3 draws
    for k in range(10):
        obs, *_ = env.step(g["actions"][k])
        assert rel_err(obs, g["obs"][k]) < 1e-9


def test_run_sim_matches_live_reference(golden_single):
    from mr_rl_b200 import run_sim
    g = golden_single.case("c1_mismatch_circle")
    X, Y, alpha, time, freq = run_sim(g["actions"], init_pos=np.array([0, 0]), noise_var=0.5, a0=1.5,
                                      is_mismatched=True, device="cuda:0", noise="table", noise_table=g["z"][:, None])
    assert rel_err(np.stack([X, Y], 1), g["pos"]) < 1e-9
    assert np.array_equal(alpha, g["actions"][:, 1]) and np.array_equal(freq, g["actions"][:, 0])
    assert np.allclose(time, np.linspace(0, (len(X) - 1) / 30.0, len(X)))


@pytest.mark.parametrize("surrogate", [True, False])
def test_device_heading_correction_matches_reference_predict(golden_gp, surrogate):
    """LearningModule.predict (bounded scalar minimisation with the GPs in the loop) on the device vs the
    live reference's outputs (alpha, muX, muY, sigX, sigY) and vs scipy's minimiser on the oracle objective.
    surrogate=True: the search runs on the verified Chebyshev interpolants of the two GP means (the default);
    False: on the direct kernel sums."""
    from scipy.optimize import minimize_scalar
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from mr_rl_b200 import LearningModule
    g = golden_gp
    X = g["X"].reshape(-1, 1)
    gX = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None).fit(X, g["yx"])
    gY = GaussianProcessRegressor(kernel=RBF(0.25) + WhiteKernel(0.008), optimizer=None).fit(X, g["yy"])
    a0, freq, Dx, Dy = g["hyper"]
    lm = LearningModule(device="cuda:0")
    lm.heading_surrogate = surrogate
    lm.set_models(gX, gY, a0, freq, Dx, Dy)
    assert (lm._cheb is not None) == surrogate
    if surrogate:
        assert 32 < lm._cheb.shape[1] < 600                   # ~exp(-(k l / pi)^2 / 2) decay at l = 0.2: a few hundred terms
    vd = g["vd"]
    alpha, mx, my, sx, sy, nfev = lm.predict_batch(vd, return_nfev=True)
    alpha = alpha.cpu().numpy()
    ref = g["predict"]                                   # [24, 5] from the unmodified reference
    # xatol = 1e-5 is the algorithm's own resolution; iterates agree far better when no branch flips
    assert np.max(np.abs(alpha[:24] - ref[:, 0])) < 2e-5
    got = np.stack([mx.cpu().numpy(), my.cpu().numpy(), sx.cpu().numpy(), sy.cpu().numpy()], 1)
    assert np.allclose(got[:24], ref[:, 1:], rtol=1e-4, atol=1e-6)
    gx, gy = fitted_pair(g)
    nf = nfev.cpu().numpy()
    same = 0
    for i in range(len(vd)):
        r = minimize_scalar(lambda a: float(np.ravel(mo.lm_objective(a, a0, freq, vd[i], gx, gy, Dx, Dy))[0]),
                            method="Bounded", bounds=[-np.pi, np.pi])
        assert abs(alpha[i] - r.x) < 2e-5
        same += int(nf[i] == r.nfev)
    assert same >= int(0.9 * len(vd))                    # identical iteration path (same number of objective calls)
    # the scalar facade goes through the same kernel
    A, muX, muY, sigX, sigY = lm.predict(vd[0])
    assert abs(float(A) - ref[0, 0]) < 2e-5 and muX.shape == (1,)


def test_cheb_heading_entry_point_argument_errors_and_constant_series():
    """mr_gp_correct_heading_cheb: argument checks, and a closed-form case — with constant GP means (one coefficient each)
    the objective is (a0 f)^2 + ex^2 + ey^2 + 2 a0 f (ex cos a + ey sin a), minimised at a = atan2(-ey, -ex)."""
    import ctypes as C
    from mr_rl_b200 import _lib as L
    lib = L.load()
    cx = torch.tensor([0.3], dtype=torch.float64, device="cuda:0")
    cy = torch.tensor([-0.2], dtype=torch.float64, device="cuda:0")
    vd = torch.tensor([[2.0, 1.0], [-1.0, 0.5], [0.1, -3.0]], dtype=torch.float64, device="cuda:0")
    out = torch.empty(3, dtype=torch.float64, device="cuda:0")
    nf = torch.empty(3, dtype=torch.int32, device="cuda:0")
    assert lib.mr_gp_correct_heading_cheb(None, cy.data_ptr(), 1, vd.data_ptr(), 3, 1.5, 4.0, 0.0, 0.0, out.data_ptr(), None, None) != 0
    assert lib.mr_gp_correct_heading_cheb(cx.data_ptr(), cy.data_ptr(), 0, vd.data_ptr(), 3, 1.5, 4.0, 0.0, 0.0, out.data_ptr(), None, None) != 0
    assert b"n_coef" in lib.mr_last_error()
    assert lib.mr_gp_correct_heading_cheb(cx.data_ptr(), cy.data_ptr(), 1, None, 3, 1.5, 4.0, 0.0, 0.0, out.data_ptr(), None, None) != 0
    rc = lib.mr_gp_correct_heading_cheb(cx.data_ptr(), cy.data_ptr(), 1, vd.data_ptr(), 3, 1.5, 4.0, 0.05, -0.02, out.data_ptr(),
                                        nf.data_ptr(), None)
    assert rc == 0
    ex = 0.3 + 0.05 - vd[:, 0].cpu().numpy()
    ey = -0.2 - 0.02 - vd[:, 1].cpu().numpy()
    want = np.arctan2(-ey, -ex)
    assert np.max(np.abs(np.angle(np.exp(1j * (out.cpu().numpy() - want))))) < 2e-5
    assert int(nf.min()) >= 5 and int(nf.max()) < 60


def test_gp_posterior_at_config3_size_against_sklearn_spot_check():
    """BASELINE configs[3] at its stated size — 262 144 queries against 2000 training points, both forms of the variance —
    checked against sklearn's GaussianProcessRegressor.predict(return_std=True) on every 64th query (4096 of them)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel

    from mr_rl_b200 import DeviceGP
    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(-np.pi, np.pi, 2000))
    y = 0.2 + 0.5 * np.cos(X + 0.3) + 0.09 * rng.standard_normal(2000)
    gpr = GaussianProcessRegressor(kernel=RBF(0.2) + WhiteKernel(0.008), optimizer=None, alpha=1e-10).fit(X[:, None], y)
    q = rng.uniform(-np.pi, np.pi, 262144)
    mu_ref, sd_ref = gpr.predict(q[::64, None], return_std=True)
    gp = DeviceGP.fit(X, y, 0.2, 0.008, device="cuda:0")
    qd = torch.as_tensor(q, device="cuda:0")
    mu, sd = gp.predict(qd, True)                                        # triangular form = sklearn's algorithm
    assert np.allclose(mu.cpu().numpy()[::64], mu_ref, rtol=1e-8, atol=1e-10)
    assert np.allclose(sd.cpu().numpy()[::64], sd_ref, rtol=1e-6, atol=1e-12)
    assert gp.enable_spectral_variance() > 0 and gp.spectral_build_ms > 0
    mu2, sd2 = gp.predict(qd, True)                                      # low-rank form, all 262 144 queries against the triangular one
    assert float((mu2 - mu).abs().max()) < 1e-10
    assert float(((sd2 - sd).abs() / sd).max()) < 1e-8
