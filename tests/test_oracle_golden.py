"""CPU: the oracle restatement (oracle/mr_oracle.py) against the golden vectors recorded
from the live reference (tests/golden/*.npz, written by oracle/gen_golden.py)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import mr_oracle as mo

SINGLE_CASES = ["c1_sigma0", "c1_sigma1", "c1_mismatch_circle", "c1_mismatch_idle", "c1_default_sim_params",
                "near_goal", "out_of_bounds", "float32_init", "reset_after_mismatch", "reset_into_mismatch",
                "noisefree_mismatch"]


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_single_env_restatement_matches_live_reference(golden_single, name):
    g = golden_single.case(name)
    sig, a0, mism, prior_mism = g["params"]
    r = mo.rollout(g["actions"], g["init"], sig, a0, bool(mism), g["z"], mism_before_reset=bool(prior_mism))
    # integer / boolean quantities: bit-exact
    assert np.array_equal(r["done"], g["done"])
    assert np.array_equal(r["counter"], g["counter"])
    assert np.array_equal(r["cursor"], g["cursor"])          # identical number of noise draws
    assert np.array_equal(r["attempts"], g["attempts"])
    assert int(r["reset_cursor"]) == int(g["reset_cursor"])
    assert np.all(r["rew"] == 10) and np.all(g["rew"] == 10)
    # floating point: 1e-9 relative (north_star); the restatement is in fact ~1e-13
    for k in ("pos", "obs", "state_prime", "carry_f", "carry_h", "t", "reset_obs", "reset_carry_h"):
        assert rel_err(r[k], g[k]) < 1e-9, k
    assert (r["status"] == 0).all()


def test_golden_files_record_versions(golden_single, golden_batch, golden_gp):
    for g in (golden_single, golden_batch, golden_gp):
        v = [str(s) for s in g["versions"]]
        assert any(s.startswith("scipy") for s in v) and any(s.startswith("numpy") for s in v)


@pytest.mark.parametrize("tag", ["sigma0", "sigma1", "mismatch"])
def test_batch_restatement(golden_batch, tag):
    sig, a0, mism, _ = golden_batch[f"{tag}/params"]
    acts, init, z = golden_batch["actions"], golden_batch["init"], golden_batch["z"]
    for e in range(0, acts.shape[1], 7):
        r = mo.rollout(acts[:, e], init[e], sig, a0, bool(mism), z[e])
        assert np.array_equal(r["done"], golden_batch[f"{tag}/done"][:, e])
        assert np.array_equal(r["cursor"], golden_batch[f"{tag}/cursor"][:, e])
        assert rel_err(r["pos"], golden_batch[f"{tag}/pos"][:, e]) < 1e-9
        assert rel_err(r["obs"], golden_batch[f"{tag}/obs"][:, e]) < 1e-9


def test_stale_action_blend_noise_free():
    """y_{k+1} = y_k + dt*[(35/384) v(a_{k-1}) + (349/384) v(a_k)] once h = dt."""
    rng = np.random.default_rng(3)
    acts = np.stack([rng.uniform(1, 20, 30), rng.uniform(0, 2 * np.pi, 30)], 1)
    r = mo.rollout(acts, [110.0, 105.0], 0.0, 1.0, False, None)
    v = acts[:, :1] * np.stack([np.cos(acts[:, 1]), np.sin(acts[:, 1])], 1)
    for k in range(2, 30):
        if r["attempts"][k] != 1:
            continue
        pred = r["pos"][k - 1] + (r["t"][k] - r["t"][k - 1]) * ((35 / 384) * v[k - 1] + (349 / 384) * v[k])
        assert np.allclose(pred, r["pos"][k], rtol=1e-12, atol=0)


def test_done_truth_table():
    assert mo.is_done(mo.convert_state(100.0, 100.0), 50) is False
    assert mo.is_done(mo.convert_state(100.0, 100.0), 51) is True           # counter > max_timesteps
    assert mo.is_done(mo.convert_state(10.0, 10.0), 1) is True              # d < 30
    assert mo.is_done(mo.convert_state(5000.5, 0.0), 1) is True             # outside observation_space
    assert mo.is_done(mo.convert_state(-5000.0, 100.0), 1) is False         # bound is inclusive
    assert mo.is_done(mo.convert_state(float("nan"), 0.0), 1) is True       # NaN is not contained
    assert mo.shaped_reward(mo.convert_state(10.0, 10.0), 1) == 100.0
    assert mo.shaped_reward(mo.convert_state(100.0, 100.0), 51) == -100.0
    assert mo.shaped_reward(mo.convert_state(100.0, 100.0), 3) == -0.1


def test_t_table_matches_accumulated_time(golden_single):
    g = golden_single.case("c1_sigma1")
    tt = mo.t_table(len(g["t"]))
    assert np.array_equal(tt[1:], g["t"])


def test_gp_restatement_matches_reference_learning_module(golden_gp):
    g = golden_gp
    gx = mo.fit_fixed_gp(g["X"], g["yx"], float(g["lsx"]), float(g["noise"]))
    gy = mo.fit_fixed_gp(g["X"], g["yy"], float(g["lsy"]), float(g["noise"]))
    assert rel_err(gx.alpha, g["alpha_x"]) < 1e-6
    mx, sx = mo.gp_predict(gx, g["grid"].reshape(-1, 1))
    my, sy = mo.gp_predict(gy, g["grid"].reshape(-1, 1))
    assert np.allclose(mx, g["grid_mx"], rtol=1e-8, atol=1e-10)
    assert np.allclose(my, g["grid_my"], rtol=1e-8, atol=1e-10)
    assert np.allclose(sx, g["grid_sx"], rtol=1e-7, atol=1e-10)
    assert np.allclose(sy, g["grid_sy"], rtol=1e-7, atol=1e-10)
    err = np.array([np.concatenate(mo.lm_error(gx, gy, v)) for v in g["vd"]])
    assert np.allclose(err, g["error"], rtol=1e-7, atol=1e-10)
    a0, freq, Dx, Dy = g["hyper"]
    obj = np.array([float(np.ravel(mo.lm_objective(a, a0, freq, g["vd"][i], gx, gy, Dx, Dy))[0])
                    for i, a in enumerate(g["obj_alpha"])])
    assert np.allclose(obj, g["objective"], rtol=1e-8, atol=1e-9)


def test_actor_forward_shape_and_bounds():
    p = mo.actor_init(0)
    obs = np.array([[110.0, 105.0, 0, 0, 152.0], [0, 0, 0, 0, 0]])
    a = mo.actor_forward(p, obs)
    assert a.shape == (2, 2) and a.dtype == np.float32
    assert np.all(np.abs(a) <= np.array(mo.ACTION_HIGH, dtype=np.float32))


@pytest.mark.parametrize("name", ["c1_sigma1", "c1_mismatch_circle", "reset_after_mismatch", "near_goal"])
def test_scipy_env_port_matches_live_reference(golden_single, name):
    """The CPU-baseline port built on the real scipy RK45 reproduces the live reference."""
    from oracle.scipy_env import ScipyEnv
    g = golden_single.case(name)
    sig, a0, mism, prior = g["params"]
    cur = [0]

    def normal(mu, s, n):
        v = mu + s * g["z"][cur[0]]
        cur[0] += 1
        return np.array([v])

    env = ScipyEnv(normal=normal)
    env.mism = bool(prior)
    env.reset(g["init"], noise_var=sig, a0=a0, is_mismatched=bool(mism))
    assert cur[0] == int(g["reset_cursor"])
    T = min(len(g["actions"]), 120)
    for k in range(T):
        obs, rew, done, _ = env.step(g["actions"][k])
        assert rel_err(obs, g["obs"][k]) < 1e-12
        assert done == bool(g["done"][k]) and cur[0] == int(g["cursor"][k]) and rew == 10
