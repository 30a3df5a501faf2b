"""Single-env façades with the reference's exact class surface, backed by the CUDA kernels.

``MR_Env``    (MR_env.py:21-229)      gym-style env returning numpy, usable by main.py / RL/MR_ddpg.py
``Simulator`` (MR_simulator.py:8-94)  the bare simulator object MR_Env owns

Both hold a one-env ``VecMREnv``; every call is a kernel launch plus a device->host read, so they
exist for drop-in compatibility, not speed — batch with ``VecMREnv`` for throughput.
"""
from __future__ import annotations

import numpy as np
import torch

from .spaces import Box
from .vec_env import VecMREnv


class Simulator:
    """Reference attribute names: a0, noise_var, is_mismatched, time_span, number_iterations,
    last_state, current_action, state_prime; methods reset_start_pos, step, get_state."""

    def __init__(self, device="cuda", noise="philox", seed=0, noise_table=None, _vec=None):
        self._vec = _vec or VecMREnv(1, device=device, dtype=torch.float64, noise=noise, seed=seed,
                                     noise_table=noise_table, host_mapped_aux=True)
        self._a_pin = torch.zeros(1, 2, dtype=torch.float64).pin_memory()
        self.time_span = self._vec.time_span          # MR_simulator.py:12
        self.number_iterations = 100                  # :13
        self.a0 = 0                                   # :16
        self.noise_var = 0                            # :18
        self.is_mismatched = False                    # :19
        self.last_state = None
        self.current_action = None
        self.state_prime = None
        self._mism_live = False

    def reset_start_pos(self, state_vector):
        """MR_simulator.py:21-34 — builds the integrator with whatever is_mismatched currently is."""
        v = self._vec
        v.params.is_mismatched = 1 if self.is_mismatched else 0
        v.reset(init=np.asarray(state_vector, dtype=np.float64)[:2], noise_var=self.noise_var, a0=self.a0,
                is_mismatched=self.is_mismatched)
        x0, y0 = state_vector[0], state_vector[1]
        self.last_state = np.array([x0, y0])
        self.current_action = np.zeros(2)
        self.state_prime = v.state_prime[0].cpu().numpy().copy()

    def step(self, f_t, alpha_t):
        """MR_simulator.py:36-52."""
        v = self._vec
        v.params.is_mismatched = 1 if self.is_mismatched else 0
        v.params.a0 = float(self.a0)
        v.params.noise_var = float(self.noise_var)
        self.current_action = np.array([f_t, alpha_t])
        self._a_pin[0, 0] = float(f_t)                  # same one-launch host step as MR_Env.step
        self._a_pin[0, 1] = float(alpha_t)
        obs_h, _, _, _ = v.step_host(self._a_pin)
        if v._status[0] != 0:
            v.check_status()
        self.last_state = obs_h[0, :2].copy()
        self.state_prime = v._sp[:, 0].numpy().copy() if v._sp.device.type == "cpu" else v.state_prime[0].cpu().numpy().copy()
        return self.last_state

    def get_state(self):
        return self.last_state


class MR_Env:
    """Drop-in for the reference ``MR_Env`` (gym.Env is not required: the reference uses only the
    Box spaces from gym)."""

    def __init__(self, type="continuous", action_dim=2, device="cuda", noise="philox", seed=0, noise_table=None):
        self.type = type
        self.action_dim = action_dim
        self._vec = VecMREnv(1, device=device, dtype=torch.float64, noise=noise, seed=seed, noise_table=noise_table,
                             host_mapped_aux=True)
        v = self._vec
        self._a_pin = torch.zeros(1, 2, dtype=torch.float64).pin_memory()      # the action, read by the kernel in place
        self.action_space = v.action_space
        self.observation_space = v.observation_space
        self.init_space = v.init_space
        self.init_goal_space = Box(low=np.array([-31, -31]), high=np.array([-32, -32]))   # MR_env.py:43-45
        self.borders = [[-510, 510], [-510, -510], [510, -510], [510, 510]]
        self.simulator = Simulator(_vec=v)
        self.test_performance = False
        self.last_pos = np.zeros(2)
        self.init_goal = np.zeros(2)
        self.last_action = np.zeros(self.action_dim)
        self.number_loop = 0
        self.counter = 0
        self.max_timesteps = 50
        self.min_dist2goal = 30
        self.viewer = None
        self.MR_data = None
        self.name_experiment = None
        self.state_prime = None

    def _sync_limits(self):
        p = self._vec.params
        p.max_timesteps = int(self.max_timesteps)
        p.min_dist2goal = float(self.min_dist2goal)

    def step(self, action):
        """MR_env.py:70-98 -> (obs (5,) float64, 10, bool, {})."""
        v = self._vec
        sim = self.simulator
        self._sync_limits()
        f_t, alpha_t = action[0], action[1]
        v.params.is_mismatched = 1 if sim.is_mismatched else 0
        v.params.a0 = float(sim.a0)
        v.params.noise_var = float(sim.noise_var)
        # one launch, one stream synchronisation: the kernel reads the action from and writes obs / reward / done /
        # state_prime / status to page-locked host memory itself (mr_env_step_host, direct mode)
        self._a_pin[0, 0] = float(f_t)
        self._a_pin[0, 1] = float(alpha_t)
        obs_h, _, done_h, _ = v.step_host(self._a_pin)
        if v._status[0] != 0:
            v.check_status()
        self.counter += 1
        obs = obs_h[0].copy()
        self.state_prime = sim.state_prime = v._sp[:, 0].numpy().copy()
        sim.last_state = obs[:2].copy()
        sim.current_action = np.array([f_t, alpha_t])
        done = bool(done_h[0])
        rew = 10                                              # MR_env.py:89
        self.last_pos = [obs[0], obs[1]]
        self.last_action = np.array([f_t, alpha_t])
        if self.MR_data is not None:
            self.MR_data.new_transition(sim.last_state, obs, self.last_action, rew)     # MR_env.py:94-95
            if done and (not self.observation_space.contains(obs) or self.counter > self.max_timesteps) \
                    and self.MR_data.iterations > 0:
                self.MR_data.save_experiment(self.name_experiment)                      # MR_env.py:145-147
        return obs, rew, done, dict()

    def convert_state(self, state, goal_loc):
        """MR_env.py:100-116 (host helper; the kernels compute the same row)."""
        x, y, gx, gy = state[0], state[1], goal_loc[0], goal_loc[1]
        d = np.linalg.norm(np.array((gx, gy)) - np.array((x, y)))
        return np.array([x, y, gx, gy, d])

    def calculate_reward(self, obs):
        """MR_env.py:118-134 (unused by step, as in the reference)."""
        d = obs[4]
        if d < self.min_dist2goal:
            return 100
        if not self.observation_space.contains(obs) or self.counter > self.max_timesteps:
            return -100
        return -0.1

    def end(self, state, obs):
        """MR_env.py:136-152."""
        d = obs[4]
        if not self.observation_space.contains(obs) or self.counter > self.max_timesteps:
            return True
        return bool(d < self.min_dist2goal)

    def set_init_space(self, low, high):
        self.init_space = Box(low=np.array(low), high=np.array(high))
        self._vec.init_space = self.init_space                 # property: writes through to the launch parameters

    def set_goal(self, init):
        return self.init_goal                                  # MR_env.py:157-162 (no-op)

    def reset(self, init=None, noise_var=1, a0=1, is_mismatched=False):
        """MR_env.py:164-201."""
        if init is None:
            init = self.init_space.sample()
        v = self._vec
        sim = self.simulator
        self._sync_limits()
        sim.noise_var = noise_var
        sim.a0 = a0
        v.params.is_mismatched = 1 if sim.is_mismatched else 0   # stale flag is what the integrator sees
        obs_t = v.reset(init=np.asarray(init, dtype=np.float64)[:2], noise_var=noise_var, a0=a0,
                        is_mismatched=is_mismatched)
        self.goal_loc = self.init_space.sample()
        sim.is_mismatched = is_mismatched
        self.last_pos = init
        self.counter = 0
        obs = obs_t[0].cpu().numpy().copy()
        v.check_status()
        sim.last_state = np.array([init[0], init[1]])
        sim.current_action = np.zeros(2)
        sim.state_prime = v.state_prime[0].cpu().numpy().copy()
        if self.MR_data is not None:
            if self.MR_data.iterations > 0:
                self.MR_data.save_experiment(self.name_experiment)
            self.MR_data.new_iter(sim.last_state, obs, np.zeros(len(self.last_action)), np.array([0]))
        return obs

    def render(self, mode="human"):
        return None

    def close(self):
        return None

    def set_save_experice(self, name="experiment_ssn_ddpg_10iter"):
        assert type(name) == type(""), "name must be a string"
        from .recording import MRExperiment
        self.MR_data = MRExperiment()
        self.name_experiment = name

    def set_test_performace(self):
        self.test_performance = True
