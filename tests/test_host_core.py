"""CPU: the per-env algorithm header the CUDA kernels are built from (mr_core.cuh), compiled
for the host with g++, against the golden vectors of the live reference.  This is the only
way to exercise that code in the GPU-less build container; the kernels themselves are
checked on the B200 by tests/test_gpu_*.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err
from test_oracle_golden import SINGLE_CASES


def _build(tmp_path_factory, name, extra):
    out = tmp_path_factory.mktemp(name) / f"lib{name}.so"
    src = os.path.join(ROOT, "tests", "host_core_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *extra, src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="session")
def host_core(tmp_path_factory):
    return _build(tmp_path_factory, "hostcore", [])


@pytest.fixture(scope="session")
def host_core_exact(tmp_path_factory):
    """Same header with the division-free shortcuts disabled (margin so large they never fire)."""
    return _build(tmp_path_factory, "hostcore_exact", ["-DMR_MARGIN=1e300"])


def run_host(lib, g):
    acts = np.ascontiguousarray(g["actions"][:, :2], dtype=np.float64)
    T = len(acts)
    sig, a0, mism, prior = g["params"]
    z = np.ascontiguousarray(g["z"], dtype=np.float64)
    pos = np.zeros((T, 2)); d = np.zeros(T); done = np.zeros(T, np.uint8); cnt = np.zeros(T, np.int32)
    cur = np.zeros(T, np.int64); att = np.zeros(T, np.int32); carry = np.zeros((T, 3)); sp = np.zeros((T, 2))
    rst = np.zeros(4)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.host_core_rollout.restype = ctypes.c_int
    st = lib.host_core_rollout(P(acts), ctypes.c_int(T), ctypes.c_double(float(g["init"][0])), ctypes.c_double(float(g["init"][1])),
                               ctypes.c_double(sig), ctypes.c_double(a0), ctypes.c_int(int(mism)), ctypes.c_int(int(prior)),
                               P(z), ctypes.c_int(len(z)), P(pos), P(d), P(done), P(cnt), P(cur), P(att), P(carry), P(sp), P(rst))
    return dict(status=st, pos=pos, d=d, done=done, counter=cnt, cursor=cur, attempts=att, carry=carry, sp=sp, reset=rst)


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_core_header_matches_live_reference(host_core, golden_single, name):
    g = golden_single.case(name)
    r = run_host(host_core, g)
    assert r["status"] == 0
    assert np.array_equal(r["done"], g["done"])
    assert np.array_equal(r["counter"], g["counter"])
    assert np.array_equal(r["cursor"], g["cursor"])
    assert np.array_equal(r["attempts"], g["attempts"])
    assert int(r["reset"][3]) == int(g["reset_cursor"])
    assert rel_err(r["pos"], g["pos"]) < 1e-9
    assert rel_err(r["d"], g["obs"][:, 4]) < 1e-9
    assert rel_err(r["carry"][:, :2], g["carry_f"]) < 1e-9
    assert rel_err(r["carry"][:, 2], g["carry_h"]) < 1e-9
    assert rel_err(r["sp"], g["state_prime"]) < 1e-9
    assert rel_err(r["reset"][2], g["reset_carry_h"]) < 1e-9


def test_division_free_shortcuts_equal_the_exact_formulas(host_core, host_core_exact):
    """The shortcuts in ctor()/sim_step() skip divisions, square roots and pow only when the exact
    scipy expression is decided by a 1e-9 margin; results must be BIT-identical to evaluating
    everything, across regimes (near the origin, far away, tiny/huge actions, with/without noise)."""
    rng = np.random.default_rng(7)
    T = 40
    for case in range(300):
        scale = 10.0 ** rng.uniform(-6, 3.6)
        init = rng.uniform(-1, 1, 2) * scale
        acts = np.stack([rng.uniform(0, 20, T) * 10.0 ** rng.uniform(-4, 0), rng.uniform(0, 2 * np.pi, T)], 1)
        sig = float(rng.choice([0.0, 0.01, 1.0, 5.0]))
        mism = int(rng.integers(0, 2))
        g = {"actions": acts, "init": init, "params": (sig, float(rng.uniform(0.5, 2)), mism, int(rng.integers(0, 2))),
             "z": rng.standard_normal(T * 1500)}
        a = run_host(host_core, g)
        b = run_host(host_core_exact, g)
        assert a["status"] == b["status"]
        for k in ("pos", "d", "done", "counter", "cursor", "attempts", "carry", "sp", "reset"):
            assert np.array_equal(a[k], b[k]), (case, k)
