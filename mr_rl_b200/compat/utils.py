"""Drop-in module for ``from utils import run_sim, find_alpha_corrected`` (main.py:3)."""
from mr_rl_b200.utils import run_sim  # noqa: F401


def find_alpha_corrected(v_desired, gp):
    """utils.py:194-196."""
    alpha_corrected, muX, muY, sigX, sigY = gp.predict(v_desired)
    return alpha_corrected, muX, muY, sigX, sigY
