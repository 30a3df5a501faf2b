"""Drop-in module for ``import Learning_module_2d as GP`` (main_2d.py): GP.LearningModule is the 2-D variant with the GPs
on the device."""
from mr_rl_b200.learning_module_2d import LearningModule2D as LearningModule  # noqa: F401
