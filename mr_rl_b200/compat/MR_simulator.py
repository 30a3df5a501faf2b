"""Drop-in module for ``from MR_simulator import Simulator`` (MR_env.py:11)."""
from mr_rl_b200.mr_env import Simulator  # noqa: F401
