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


def experiment_from_rollout(res, info=None):
    """The reference's MRExperiment dict from a device-recorded rollout (``VecMREnv.rollout(..., record_episodes=True)``),
    episodes split at the terminal steps the kernel flagged — any number of episodes per env, auto resets included.
    Episodes ("iterations") are numbered env-major: all episodes of env 0 in time order, then env 1, ...  Row 0 of an
    episode is what ``new_iter`` logs at reset (start state, its observation, a zero action, reward [0],
    MR_env.py:190-198), the following rows what ``new_transition`` logs per step (MR_env.py:94-95).
    Vectorised: the per-row arrays are assembled with numpy index arithmetic over all N x K transitions at once and cut
    into per-episode views with one ``np.split`` (no Python loop over envs or steps)."""
    xy = _to_np(res["xy"]).astype(np.float64)                       # [K, 2, N] position after step k
    K, _, N = xy.shape
    done = _to_np(res["done_traj"]).astype(bool)                    # [K, N]
    acts = _to_np(res["actions_traj"]).astype(np.float64)           # [K, N, 2]
    rew = _to_np(res["rew_traj"]).astype(np.float64)                # [K, N]
    start = _to_np(res["start_xy"]).astype(np.float64)              # [N, 2]
    reset_xy = _to_np(res["reset_xy"]).astype(np.float64)           # [K, 2, N]
    # slot layout per env: for every step k an optional "new_iter" row (present when an episode starts at k) followed
    # by the transition row -> [N, K, 2 slots]; flattening env-major puts every episode's rows in order
    starts = np.zeros((K, N), bool)
    starts[0] = True
    starts[1:] = done[:-1]                                          # an episode starts after every terminal step
    begin_xy = np.empty((K, 2, N))
    begin_xy[0] = start.T
    begin_xy[1:] = reset_xy[:-1]                                    # only read where starts is set
    pos = np.stack([begin_xy, xy], axis=1)                          # [K, slot, 2, N]
    pos = pos.transpose(3, 0, 1, 2).reshape(N * K * 2, 2)           # env-major rows
    act = np.stack([np.zeros_like(acts), acts], axis=1).transpose(2, 0, 1, 3).reshape(N * K * 2, 2)
    rw = np.stack([np.zeros_like(rew), rew], axis=1).transpose(2, 0, 1).reshape(N * K * 2, 1)
    present = np.stack([starts, np.ones_like(starts)], axis=1).transpose(2, 0, 1).reshape(N * K * 2)
    is_start = np.stack([starts, np.zeros_like(starts)], axis=1).transpose(2, 0, 1).reshape(N * K * 2)
    pos, act, rw, is_start = pos[present], act[present], rw[present], is_start[present]
    obs = np.column_stack([pos, np.zeros((len(pos), 2)), np.hypot(pos[:, 0], pos[:, 1])])     # MR_env.py:100-116, goal (0, 0)
    cuts = np.flatnonzero(is_start)
    n_ep = len(cuts)
    lengths = np.diff(np.append(cuts, len(pos))) - 1
    ids = range(n_ep)
    return {"iterations": n_ep - 1,
            "states": dict(zip(ids, np.split(pos, cuts[1:]))), "observations": dict(zip(ids, np.split(obs, cuts[1:]))),
            "actions": dict(zip(ids, np.split(act, cuts[1:]))), "rewards": dict(zip(ids, np.split(rw, cuts[1:]))),
            "steps": dict(zip(ids, lengths.tolist())),
            "info": info, "viewer": None, "scream": None, "obs_states_str": {}, "time_step": 10}


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
