"""TEST INFRASTRUCTURE — run the unmodified reference in the build container.

Imports ``MR_env.MR_Env`` / ``utils.run_sim`` / ``Learning_module`` straight from
``/root/reference`` (never copied) behind stub modules for the packages the image
lacks (gym, matplotlib, turtle, tkinter), and forces identical inputs by patching
``numpy.random.normal`` — which the reference looks up at call time
(MR_simulator.py:56,79-83) — to pop from a caller-supplied standard-normal stream.

Only ``oracle/gen_golden.py`` and the container-only tests use this module; it
cannot run on the GPU box (no ``/root/reference`` there).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MR_RL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "MR_env.py"))


class _Anything:
    """Attribute/ call sink for plotting and GUI stubs."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class Box:
    """gym.spaces.Box as the reference uses it (MR_env.py:34-45): float32 bounds,
    uniform ``sample`` and a bounds-only ``contains`` (last bullet)."""

    def __init__(self, low, high, dtype=np.float32, seed=None):
        self.low = np.asarray(low).astype(dtype)
        self.high = np.asarray(high).astype(dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Env:  # gym.Env base
            pass

        spaces.Box = Box
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.ticker",
                 "matplotlib.patches", "matplotlib.cm", "turtle", "tkinter"):
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "animation", "ticker", "patches", "cm"):
        if isinstance(mpl, _StubModule):
            mpl.__dict__[sub] = sys.modules["matplotlib." + sub]


_loaded = {}


def load():
    """Return the reference modules {MR_env, MR_simulator, utils, Learning_module}."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    with contextlib.redirect_stdout(io.StringIO()):
        for name in ("MR_simulator", "MR_env", "utils", "Learning_module"):
            _loaded[name] = importlib.import_module(name)
    return _loaded


def load_2d():
    """Learning_module_2d.py (the (alpha, f) variant of the disturbance GPs) from the reference tree."""
    if "Learning_module_2d" not in _loaded:
        load()
        import importlib
        with contextlib.redirect_stdout(io.StringIO()):
            _loaded["Learning_module_2d"] = importlib.import_module("Learning_module_2d")
    return _loaded["Learning_module_2d"]


def load_ddpg():
    """RL/MR_ddpg.py imported from the reference tree behind stubs for tensorflow / tflearn (absent from the image):
    its ReplayBuffer (:16-56) and OUNoise (:58-78) are plain Python / numpy and run as they are; the TensorFlow
    networks cannot (so the learner's graph stays unpinned, see oracle/ddpg_oracle.py)."""
    if "MR_ddpg" in _loaded:
        return _loaded["MR_ddpg"]
    load()
    for name in ("tensorflow", "tflearn", "tflearn.layers", "tflearn.layers.normalization", "tflearn.activations",
                 "tflearn.initializations"):
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    import importlib.util
    spec = importlib.util.spec_from_file_location("MR_ddpg", os.path.join(REFERENCE_ROOT, "RL", "MR_ddpg.py"))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    _loaded["MR_ddpg"] = mod
    return mod


class NoiseStream:
    """Replacement for ``numpy.random.normal`` popping from a fixed z-stream."""

    def __init__(self, z):
        self.z = np.asarray(z, dtype=np.float64)
        self.cursor = 0

    def __call__(self, mu=0.0, sigma=1.0, size=None):
        n = 1 if size is None else int(np.prod(size))
        out = np.empty(n)
        for i in range(n):
            out[i] = mu + sigma * self.z[self.cursor]
            self.cursor += 1
        return out if size is not None else out[0]


@contextlib.contextmanager
def patched_noise(z):
    stream = NoiseStream(z)
    orig = np.random.normal
    np.random.normal = stream
    try:
        yield stream
    finally:
        np.random.normal = orig


def new_env():
    mods = load()
    with contextlib.redirect_stdout(io.StringIO()):
        return mods["MR_env"].MR_Env()


def rollout(actions, init, noise_var, a0, is_mismatched, z, env=None, record=True):
    """reset + step through ``actions`` on the live reference, straight through ``done``.

    Returns a dict of per-step arrays: pos, obs, done, counter,
    state_prime, cursor (after the step), attempts, plus the carried integrator state
    (f, h_abs, t) after each step and everything after reset under ``reset_*``.
    """
    env = env or new_env()
    T = len(actions)
    out = {
        "pos": np.zeros((T, 2)), "obs": np.zeros((T, 5)), "done": np.zeros(T, np.uint8),
        "counter": np.zeros(T, np.int32), "state_prime": np.zeros((T, 2)),
        "cursor": np.zeros(T, np.int64), "attempts": np.zeros(T, np.int32),
        "carry_f": np.zeros((T, 2)), "carry_h": np.zeros(T), "t": np.zeros(T),
        "rew": np.zeros(T),
    }
    with patched_noise(z) as ns, contextlib.redirect_stdout(io.StringIO()):
        obs0 = env.reset(init=None if init is None else np.array(init), noise_var=noise_var,
                         a0=a0, is_mismatched=is_mismatched)
        integ = env.simulator.integrator
        out["reset_obs"] = np.asarray(obs0, dtype=np.float64)
        out["reset_cursor"] = np.int64(ns.cursor)
        out["reset_carry_f"] = np.array(integ.f, dtype=np.float64)
        out["reset_carry_h"] = np.float64(integ.h_abs)
        out["reset_state_prime"] = np.array(env.simulator.state_prime, dtype=np.float64)
        for k in range(T):
            integ = env.simulator.integrator
            obs, rew, done, _ = env.step(actions[k])
            nfev = integ.nfev  # the integrator that just finished (ctor's 2 evals included)
            new = env.simulator.integrator
            out["pos"][k] = env.last_pos
            out["obs"][k] = obs
            out["done"][k] = done
            out["rew"][k] = rew
            out["counter"][k] = env.counter
            out["state_prime"][k] = env.state_prime
            out["cursor"][k] = ns.cursor
            out["attempts"][k] = (nfev - 2) // 6
            out["carry_f"][k] = new.f
            out["carry_h"][k] = new.h_abs
            out["t"][k] = new.t
    out["env"] = env
    return out
