"""TEST INFRASTRUCTURE / CPU BASELINE — the env restated on top of the real third-party
``scipy.integrate.RK45`` (the dependency the reference calls at MR_simulator.py:4,31-34,42-50,90-91).

It keeps the reference's cost model — a Python RHS callback, a fresh RK45 object after every env
step, numpy small-array work — so timing it on the host cores is the honest "reference numpy env"
baseline where /root/reference itself cannot travel (the GPU box).  Pinned against the golden
vectors by tests/test_oracle_golden.py.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import RK45

from .mr_oracle import ATOL, MAX_TIMESTEPS, MIN_DIST2GOAL, OBS_HIGH, OBS_LOW, RTOL, TIME_SPAN

_LOW = np.array(OBS_LOW, dtype=np.float32)
_HIGH = np.array(OBS_HIGH, dtype=np.float32)


class ScipyEnv:
    """One env: MR_Env.reset/step semantics (MR_env.py:70-98,164-201) around scipy's RK45."""

    def __init__(self, normal=None):
        self.normal = normal or np.random.normal     # noise source, patched in parity tests
        self.a0, self.sigma, self.mism = 0.0, 0.0, False
        self.act = np.zeros(2)
        self.solver = None
        self.state_prime = None
        self.counter = 0
        self.y = np.zeros(2)

    def _rhs(self, t, y):
        f, al = self.act[0], self.act[1]
        s = self.sigma
        if self.mism:                                # MR_simulator.py:56,78-80
            a0 = self.a0 + (f / 4) * 0.8 + self.normal(0, s / 4, 1)[0]
            out = np.array([a0 * f * np.cos(al + 0.1) + self.normal(0, s, 1)[0] + 0.2,
                            a0 * f * np.sin(al - 0.15) + self.normal(0, s, 1)[0] - 0.1])
        else:                                        # MR_simulator.py:82-83
            out = np.array([self.a0 * f * np.cos(al) + self.normal(0, s, 1)[0],
                            self.a0 * f * np.sin(al) + self.normal(0, s, 1)[0]])
        self.state_prime = out
        return out

    def _new_solver(self, t0):
        self.solver = RK45(self._rhs, t0, self.y, t0 + TIME_SPAN, rtol=RTOL, atol=ATOL)

    def _obs(self):
        x, y = self.y[0], self.y[1]
        return np.array([x, y, 0.0, 0.0, np.linalg.norm(np.array((0.0, 0.0)) - np.array((x, y)))])

    def reset(self, init, noise_var=1, a0=1, is_mismatched=False):
        self.sigma, self.a0 = noise_var, a0
        self.y = np.array([init[0], init[1]])
        self.act = np.zeros(2)
        self._new_solver(0)                          # built with the OLD mismatch flag (MR_env.py:181)
        self.mism = is_mismatched                    # MR_env.py:183
        self.counter = 0
        return self._obs()

    def step(self, action):
        self.counter += 1
        self.act = np.array([action[0], action[1]])
        while self.solver.status != "finished":
            self.solver.step()
        self.y = self.solver.y
        self._new_solver(self.solver.t)
        obs = self._obs()
        inside = bool(np.all(obs >= _LOW) and np.all(obs <= _HIGH))
        done = (not inside or self.counter > MAX_TIMESTEPS) or bool(obs[4] < MIN_DIST2GOAL)
        return obs, 10, done, {}


def _worker(args):
    """Step ``n_env`` envs ``n_steps`` times with random actions; returns (env_steps, seconds)."""
    import time
    seed, n_env, n_steps, sigma = args
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    envs = [ScipyEnv() for _ in range(n_env)]
    for e in envs:
        e.reset(rng.uniform(100, 120, 2).astype(np.float32), noise_var=sigma, a0=1)
    acts = np.stack([rng.uniform(0, 20, (n_steps, n_env)), rng.uniform(0, 2 * np.pi, (n_steps, n_env))], -1)
    t0 = time.perf_counter()
    for k in range(n_steps):
        for i, e in enumerate(envs):
            _, _, done, _ = e.step(acts[k, i])
            if done:
                e.reset(rng.uniform(100, 120, 2).astype(np.float32), noise_var=sigma, a0=1)
    return n_env * n_steps, time.perf_counter() - t0


def throughput(n_procs, envs_per_proc, n_steps, sigma=1.0):
    """Aggregate env-steps/s of ``n_procs`` worker processes (one per host core)."""
    import multiprocessing as mp
    import time
    jobs = [(1000 + r, envs_per_proc, n_steps, sigma) for r in range(n_procs)]
    t0 = time.perf_counter()
    if n_procs == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {"env_steps": steps, "wall_s": wall, "busy_s": busy, "steps_per_s": steps / busy}
