"""The C ABI used from plain C (tests/c_abi_demo.c): no Python objects, no torch types across the boundary."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err


def build_demo(tmp):
    exe = os.path.join(tmp, "c_abi_demo")
    lib_dir = os.path.join(ROOT, "mr_rl_b200", "_lib")
    cmd = ["gcc", "-std=c99", "-Wall", os.path.join(ROOT, "tests", "c_abi_demo.c"), "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", "-L" + lib_dir, "-lmr_rl_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_c_client_compiles_and_links(tmp_path):
    """CPU: the header is valid C99 and every symbol the client uses resolves against the library."""
    assert os.path.exists(build_demo(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True])
def test_c_client_reproduces_reference_trajectory(tmp_path, golden_single, host):
    """host=True: every step through mr_env_step_host with page-locked host buffers from cudaHostAlloc (direct mode)."""
    g = golden_single.case("c1_sigma0")
    exe = build_demo(str(tmp_path))
    T = 60
    args = [repr(float(g["init"][0])), repr(float(g["init"][1])), "1.0", str(T)]
    for k in range(T):
        args += [repr(float(g["actions"][k, 0])), repr(float(g["actions"][k, 1]))]
    env = dict(os.environ, MR_DEMO_HOST="1" if host else "0")
    out = subprocess.run([exe, *args], capture_output=True, text=True, check=True, env=env).stdout
    rows = np.array([[float(v) for v in line.split()] for line in out.strip().splitlines()])
    assert rows.shape == (T, 5)
    assert rel_err(rows[:, :2], g["pos"][:T]) < 1e-9 and rel_err(rows[:, 2], g["obs"][:T, 4]) < 1e-9
    assert np.all(rows[:, 3] == 10.0) and np.array_equal(rows[:, 4].astype(np.uint8), g["done"][:T])
