"""Import-path shim for RL/MR_ddpg.py: the learner pieces under the reference's module name.

``ReplayBuffer`` and ``OUNoise`` keep the reference's constructor arguments (plus a device); the ActorNetwork /
CriticNetwork pair, whose TensorFlow sessions are driven op by op in the reference's ``train`` (:285-305), is one
object here (``DDPGLearner``) because the whole update block is a single kernel launch; ``train`` is the vectorised
loop.  See mr_rl_b200/ddpg.py.
"""
from mr_rl_b200.ddpg import DDPGLearner, OUNoise, ReplayBuffer, init_critic, pack_critic, train  # noqa: F401
from mr_rl_b200.actor import init_actor, pack_actor  # noqa: F401
