"""Drop-in module: put ``mr_rl_b200/compat`` ahead of the reference on sys.path and
``from MR_env import MR_Env`` (utils.py:11, RL/MR_ddpg.py, RL/read_data.py:13) resolves to the
CUDA-backed class."""
from mr_rl_b200.mr_env import MR_Env  # noqa: F401


def save_frames_as_gif(frames, path="./", filename="gym_animation.gif"):
    raise NotImplementedError("rendering (MR_env.py:232-244) is outside the hot path")
