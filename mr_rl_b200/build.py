"""Build libmr_rl_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m mr_rl_b200.build [--force] [--verbose]

One object per translation unit (the kernels are split per storage dtype x noise mode so
they compile in parallel), linked into mr_rl_b200/_lib/libmr_rl_b200.so.  nvcc
cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmr_rl_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-diag-suppress", "20013,20015"]


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inl")) + \
        glob.glob(os.path.join(INCLUDE, "*.h"))


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src, obj, verbose):
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(-4, "-Xptxas")
        cmd.insert(-4, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    return r.stderr if verbose else ""


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = _deps()
    todo, objs = [], []
    for src in sources:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *deps]):
            todo.append((src, obj))
    jobs = jobs or min(len(todo), os.cpu_count() or 4) or 1
    log = []
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            for out in ex.map(lambda so: _compile(so[0], so[1], verbose), todo):
                log.append(out)
    if todo or force or _stale(LIB_PATH, objs):
        r = subprocess.run([NVCC, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
