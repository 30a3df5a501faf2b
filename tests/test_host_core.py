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


@pytest.fixture(scope="session")
def host_core(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcore") / "libhostcore.so"
    src = os.path.join(ROOT, "tests", "host_core_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


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
