"""Drop-in module for ``from MR_data import MRExperiment`` (MR_env.py:10): the host logger with the
reference's method names and pickle layout (plotting helpers not reproduced)."""
from mr_rl_b200.recording import MRExperiment  # noqa: F401
