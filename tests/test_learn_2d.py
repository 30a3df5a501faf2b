"""Learning_module_2d.py (GPs over heading and frequency) through mr_rl_b200.learning_module_2d against the live
reference (golden learn_2d.npz, oracle/gen_golden.py:gen_learn_2d)."""
import numpy as np
import pytest

from conftest import Golden, rel_err


@pytest.fixture(scope="module")
def g2():
    return Golden("learn_2d.npz")


def test_2d_preprocessing_matches_reference_on_cpu(g2):
    from mr_rl_b200.learning_module_2d import LearningModule2D
    lm = LearningModule2D.__new__(LearningModule2D)
    lm.Dx = lm.Dy = 0
    lm.estimateDisturbance(g2["px_idle"], g2["py_idle"], g2["t_idle"])
    assert rel_err(lm.Dx, g2["Dx"]) < 1e-12 and rel_err(lm.Dy, g2["Dy"]) < 1e-12        # mean over ALL frames (:58-59)
    a0, X, Yx, Yy = lm._targets(g2["px"], g2["py"], g2["alpha"], g2["freq"], g2["time"].copy())
    assert rel_err(a0, g2["a0"]) < 1e-12
    assert np.array_equal(X, g2["X"]) and rel_err(Yx, g2["Yx"]) < 1e-10 and rel_err(Yy, g2["Yy"]) < 1e-10


@pytest.mark.gpu
def test_2d_module_fit_error_predict_match_the_live_reference(g2):
    from mr_rl_b200.learning_module_2d import LearningModule2D
    lm = LearningModule2D(device="cuda:0", fit="device")
    lm.estimateDisturbance(g2["px_idle"], g2["py_idle"], g2["t_idle"])
    np.random.seed(int(g2["seed"]))
    a0 = lm.learn(g2["px"], g2["py"], g2["alpha"], g2["freq"], g2["time"].copy())
    assert rel_err(a0, g2["a0"]) < 1e-12
    assert np.allclose(lm.gprX.kernel_.theta, g2["theta_x"], atol=2e-3) and np.allclose(lm.gprY.kernel_.theta, g2["theta_y"], atol=2e-3)
    assert abs(lm.gprX.log_marginal_likelihood_value_ - g2["lml_x"]) < 1e-6 * abs(g2["lml_x"])
    assert abs(lm.gprY.log_marginal_likelihood_value_ - g2["lml_y"]) < 1e-6 * abs(g2["lml_y"])
    scale = np.abs(g2["error"][:, :2]).max()
    for vd, e_ref, p_ref in zip(g2["vd"], g2["error"], g2["predict"]):
        mx, my, sx, sy = lm.error(vd)                       # Learning_module_2d.py:226-237
        assert mx.shape == (1,) and abs(mx[0] - e_ref[0]) < 2e-3 * scale and abs(my[0] - e_ref[1]) < 2e-3 * scale
        assert abs(sx[0] - e_ref[2]) < 2e-2 * e_ref[2] and abs(sy[0] - e_ref[3]) < 2e-2 * e_ref[3]
        X, pmx, pmy, psx, psy = lm.predict(vd)              # :239-268, scipy minimize over (alpha, f)
        assert abs(np.angle(np.exp(1j * (X[0] - p_ref[0])))) < 2e-2 and abs(X[1] - p_ref[1]) < 2e-2
        # the minimiser's objective value is what has to agree; the argmin is flat along the valley
        vd_ = np.asarray(vd)
        assert abs(lm._objective(X, vd_) - lm._objective(p_ref[:2], vd_)) < 1e-4 * (1 + abs(lm._objective(p_ref[:2], vd_)))
