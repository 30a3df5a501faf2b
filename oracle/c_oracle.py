"""TEST INFRASTRUCTURE — ctypes loader for the plain-C oracle (oracle/mr_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmr_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "mr_oracle.c")):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.mr_oracle_rollout.restype = C.c_int
        _lib.mr_oracle_max_threads.restype = C.c_int
    return _lib


def rollout(init, actions, sigma, a0, mism=False, mism_at_reset=False, z=None, seed=0, auto_reset=False,
            want_pos=True, want_attempts=False):
    """init [n,2]; actions [T,n,2]; z [n,zlen] or None (internal generator).  Returns dict."""
    lib = load()
    init = np.ascontiguousarray(init, dtype=np.float64)
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    T, n = actions.shape[:2]
    zp, zlen = None, 0
    if z is not None:
        z = np.ascontiguousarray(z, dtype=np.float64)
        zp, zlen = z.ctypes.data_as(C.c_void_p), z.shape[1]
    pos = np.zeros((T, n, 2)) if want_pos else None
    done = np.zeros((T, n), np.uint8)
    cursor = np.zeros(n, np.int64)
    att = np.zeros((T, n), np.int32) if want_attempts else None
    final = np.zeros((n, 5))
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    bad = lib.mr_oracle_rollout(C.c_int(n), C.c_int(T), P(init), P(actions), C.c_double(sigma), C.c_double(a0),
                                C.c_int(int(mism)), C.c_int(int(mism_at_reset)), zp, C.c_int64(zlen),
                                C.c_uint64(seed), C.c_int(int(auto_reset)), P(pos), P(done), P(cursor), P(att), P(final))
    return {"pos": pos, "done": done, "cursor": cursor, "attempts": att, "final": final, "bad": bad}


def throughput(n_env=65536, n_steps=64, sigma=1.0, seed=0):
    """env-steps/s of the C port on all host threads (internal noise generator, auto reset)."""
    import time
    rng = np.random.default_rng(seed)
    init = rng.uniform(100, 120, (n_env, 2)).astype(np.float32).astype(np.float64)
    acts = np.stack([rng.uniform(0, 20, (n_steps, n_env)), rng.uniform(0, 2 * np.pi, (n_steps, n_env))], -1)
    rollout(init[:1024], acts[:4, :1024], sigma, 1.0, seed=seed, auto_reset=True, want_pos=False)   # warm up threads
    t0 = time.perf_counter()
    rollout(init, acts, sigma, 1.0, seed=seed, auto_reset=True, want_pos=False)
    dt = time.perf_counter() - t0
    return {"env_steps": n_env * n_steps, "seconds": dt, "steps_per_s": n_env * n_steps / dt,
            "threads": load().mr_oracle_max_threads()}
