"""Drop-in module for ``from utils import run_sim, find_alpha_corrected`` (main.py:3)."""
from mr_rl_b200.utils import run_sim  # noqa: F401


def find_alpha_corrected(v_desired, gp):
    """utils.py:194-196."""
    alpha_corrected, muX, muY, sigX, sigY = gp.predict(v_desired)
    return alpha_corrected, muX, muY, sigX, sigY


def test_gp(gp, X, Y, a0, alpha, freq, time):
    """utils.py:175-192: the per-frame ``gp.predict`` loop as ONE batched device call (LearningModule.predict_batch).
    Returns alpha_pred [T, 1] and (v_desired, v_error, v_stdv), each [T, 2]; the reference's tuple also names vx, vy,
    which are undefined there (NameError), so they are not reproduced."""
    import numpy as np
    alpha = np.asarray(alpha, float)[:len(time)]
    v_desired = np.stack([a0 * freq * np.cos(alpha), a0 * freq * np.sin(alpha)], 1)
    a_t, mux, muy, sigx, sigy = gp.predict_batch(v_desired)
    v_error = np.stack([mux.cpu().numpy(), muy.cpu().numpy()], 1)
    v_stdv = np.stack([sigx.cpu().numpy(), sigy.cpu().numpy()], 1)
    return a_t.cpu().numpy().reshape(-1, 1), (v_desired, v_error, v_stdv)


test_gp.__test__ = False      # a reference function name, not a pytest test
