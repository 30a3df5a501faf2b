"""Drop-in module for ``import Learning_module as GP`` (main.py:2): LearningModule with GP
inference on the device."""
from mr_rl_b200.learning_module import LearningModule  # noqa: F401
