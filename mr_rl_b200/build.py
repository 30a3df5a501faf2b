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


def _compile(src, obj, verbose, extra_flags=()):
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, *extra_flags, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(-4, "-Xptxas")
        cmd.insert(-4, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    return r.stderr if verbose else ""


def build(force=False, verbose=False, jobs=None, extra_flags=(), variant=""):
    """variant: build an experimental copy (extra -D flags) as libmr_rl_b200<variant>.so; load it with
    MR_LIB_PATH.  The default build takes no extra flags."""
    global OBJ_DIR, LIB_PATH
    obj_dir = OBJ_DIR + variant
    lib_path = LIB_PATH[:-3] + variant + ".so"
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    return _build(force, verbose, jobs, list(extra_flags), obj_dir, lib_path)


def _build(force, verbose, jobs, extra_flags, OBJ_DIR, LIB_PATH):
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
            for out in ex.map(lambda so: _compile(so[0], so[1], verbose, extra_flags), todo):
                log.append(out)
    if todo or force or _stale(LIB_PATH, objs):
        r = subprocess.run([NVCC, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    variant = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, extra_flags=flags, variant=variant)
    print(path)
