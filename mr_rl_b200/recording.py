"""Trajectory recording in the reference's on-disk format (MR_data.py:9-85).

``MRExperiment`` pickles its ``__dict__`` (protocol 2): ``iterations`` (index of the last episode),
``states`` / ``observations`` / ``actions`` / ``rewards`` (dict episode -> array with one row per
transition, row 0 written by ``new_iter`` at reset: state, obs, zero action, reward [0]) and ``steps``
(dict episode -> number of transitions), plus the constant fields.  ``experiment_dict`` builds exactly
that from the arrays a recorded fused rollout returns (one episode per env), so the reference's
analysis tooling (``MRExperiment.load_from_experiment``, ``RL/evaluate_learning.py``-style scripts)
keeps working on device-generated data.
"""
from __future__ import annotations

import pickle

import numpy as np


def _to_np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def experiment_dict(init_xy, actions, xy, done=None, reward=10, info=None, until_done=False):
    """init_xy [N,2]; actions [K,N,2] or [K,>=2] (same row for every env); xy [K,2,N] positions after
    each step (``VecMREnv.rollout(..., record=True)['xy']``); done [K,N] optional.  One episode
    ("iteration") per env; with ``until_done`` an episode stops at its first terminal step, as a
    caller that resets on ``done`` would log it."""
    init_xy, actions, xy = _to_np(init_xy).astype(np.float64), _to_np(actions).astype(np.float64), _to_np(xy).astype(np.float64)
    K, _, N = xy.shape
    init_xy = np.broadcast_to(init_xy.reshape(-1, 2), (N, 2))
    if actions.ndim == 2:
        actions = np.broadcast_to(actions[:, None, :2], (K, N, 2))
    done = _to_np(done).astype(bool) if done is not None else np.zeros((K, N), bool)
    exp = {"iterations": N - 1, "states": {}, "observations": {}, "actions": {}, "rewards": {}, "steps": {},
           "info": info, "viewer": None, "scream": None, "obs_states_str": {}, "time_step": 10}
    for e in range(N):
        n_steps = K
        if until_done and done[:, e].any():
            n_steps = int(np.argmax(done[:, e])) + 1
        pos = np.vstack([init_xy[e][None, :], xy[:n_steps, :, e]])                     # row 0 = reset state
        d = np.hypot(pos[:, 0], pos[:, 1])
        obs = np.column_stack([pos, np.zeros((n_steps + 1, 2)), d])                      # MR_env.py:100-116
        exp["states"][e] = pos
        exp["observations"][e] = obs
        exp["actions"][e] = np.vstack([np.zeros((1, 2)), actions[:n_steps, e, :2]])      # new_iter logs a zero action
        exp["rewards"][e] = np.vstack([np.zeros((1, 1)), np.full((n_steps, 1), reward)])  # ... and reward [0]
        exp["steps"][e] = n_steps
    return exp


def save_experiment(exp, path):
    """MRExperiment.save_experiment's pickle (protocol 2, MR_data.py:67-75) at an explicit path."""
    with open(path, "wb") as f:
        pickle.dump(exp, f, 2)


def load_experiment(path):
    with open(path, "rb") as f:
        return pickle.load(f)


class MRExperiment:
    """Host-side transition logger with the reference's interface and pickle layout (MR_data.py:9-85):
    ``new_iter`` at reset, ``new_transition`` per step, ``save_experiment`` / ``load_from_experiment``.
    The plotting helpers of the reference class are not reproduced."""

    def __init__(self, info=None):
        self.iterations = -1
        self.states, self.observations, self.actions, self.rewards, self.steps = {}, {}, {}, {}, {}
        self.info = info
        self.viewer = None
        self.scream = None
        self.obs_states_str = {}
        self.time_step = 10

    def new_iter(self, s0, obs0, a0, r0):
        self.iterations += 1
        it = self.iterations
        self.steps[it] = 0
        self.states[it], self.observations[it], self.actions[it], self.rewards[it] = s0, obs0, a0, r0

    def new_transition(self, s, obs, a, r):
        it = self.iterations
        self.steps[it] += 1
        self.states[it] = np.vstack([self.states[it], s])
        self.observations[it] = np.vstack([self.observations[it], obs])
        self.actions[it] = np.vstack([self.actions[it], a])
        self.rewards[it] = np.vstack([self.rewards[it], r])

    def save_experiment(self, descr="_experiment", directory="_experiments"):
        import datetime
        import os
        import time as t
        st = datetime.datetime.fromtimestamp(t.time()).strftime("%Y-%m-%d-%H")
        os.makedirs(directory, exist_ok=True)
        path = os.path.join(directory, st + descr)
        save_experiment(self.__dict__, path)
        return path

    def load_from_experiment(self, name, directory="_experiments"):
        import os
        self.__dict__.update(load_experiment(os.path.join(directory, name)))
