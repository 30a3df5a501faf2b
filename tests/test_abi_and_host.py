"""CPU: the C-ABI library loads and exports every symbol include/mr_rl_b200.h declares, the ctypes
structures match the C layout, argument errors are reported without touching a GPU, and the
multi-process (world_size 2, gloo) sharding / statistics logic is correct."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mr_rl_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mr_rl_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert {"mr_env_reset", "mr_env_step", "mr_env_rollout", "mr_gp_predict", "mr_actor_forward"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.EXPORTS) == set(names)
    assert lib.mr_abi_version() == _lib.ABI_VERSION == 2
    assert lib.mr_actor_param_count() == 5 * 64 + 5 * 64 + 64 * 64 + 5 * 64 + 64 * 2 + 2


def lib_counts():
    from mr_rl_b200 import _lib
    lib = _lib.load()
    return lib.mr_actor_param_count(), lib.mr_critic_param_count()


def test_ctypes_structs_match_c_layout(tmp_path):
    from mr_rl_b200 import _lib
    prog = tmp_path / "sizes.c"
    structs = ["mr_sim_params", "mr_env_state", "mr_noise", "mr_time_table", "mr_step_out", "mr_rollout_io", "mr_gp_model",
               "mr_host_step_io", "mr_ddpg_state", "mr_replay", "mr_ddpg_hyper", "mr_reset_params"]
    body = "\n".join(f'printf("{s} %zu\\n", sizeof({s}));' for s in structs)
    prog.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\nint main(void){{{body}\n'
                    'printf("off_action_high %zu\\n", offsetof(mr_sim_params, action_high));\n'
                    'printf("off_stats %zu\\n", offsetof(mr_rollout_io, stats));\n'
                    'printf("off_noise %zu\\n", offsetof(mr_gp_model, noise_level));\n'
                    'printf("off_skip %zu\\n", offsetof(mr_step_out, skip_goal_rows));\n'
                    'printf("off_bound %zu\\n", offsetof(mr_ddpg_hyper, action_bound));\n'
                    'printf("off_cap %zu\\n", offsetof(mr_replay, capacity));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(prog), "-o", str(exe)], check=True)   # the header is plain C
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    py = {"mr_sim_params": _lib.SimParams, "mr_env_state": _lib.EnvState, "mr_noise": _lib.Noise, "mr_time_table": _lib.TimeTable,
          "mr_step_out": _lib.StepOut, "mr_rollout_io": _lib.RolloutIO, "mr_gp_model": _lib.GPModel,
          "mr_host_step_io": _lib.HostStepIO, "mr_ddpg_state": _lib.DDPGState, "mr_replay": _lib.Replay,
          "mr_ddpg_hyper": _lib.DDPGHyper, "mr_reset_params": _lib.ResetParams}
    for name, cls in py.items():
        assert C.sizeof(cls) == int(out[name]), name
    assert _lib.SimParams.action_high.offset == int(out["off_action_high"])
    assert _lib.RolloutIO.stats.offset == int(out["off_stats"])
    assert _lib.GPModel.noise_level.offset == int(out["off_noise"])
    assert _lib.StepOut.skip_goal_rows.offset == int(out["off_skip"])
    assert _lib.DDPGHyper.action_bound.offset == int(out["off_bound"])
    assert _lib.Replay.capacity.offset == int(out["off_cap"])
    assert lib_counts() == (5186, 2849)


def test_default_params_mirror_reference_constants():
    from mr_rl_b200 import _lib
    p = _lib.default_params()
    assert (p.time_span, p.atol, p.max_timesteps, p.min_dist2goal) == (0.030, 1e-4, 50, 30.0)   # MR_simulator.py:12, MR_env.py:62-63
    assert p.rtol == 0.030 / 100                                                                 # MR_simulator.py:91
    assert (p.a0, p.noise_var) == (1.0, 1.0)                                                     # MR_env.py:167-168
    assert list(p.init_low) == [100.0, 100.0] and list(p.init_high) == [120.0, 120.0]
    assert p.action_high[0] == 20.0 and abs(p.action_high[1] - 2 * np.pi) < 1e-15
    assert (p.bound_xy, p.bound_d) == (5000.0, 80000.0)


def test_time_table_is_the_accumulated_sum():
    from mr_rl_b200 import _lib
    from oracle import mr_oracle as mo
    lib = _lib.load()
    t = np.zeros(600)
    lib.mr_fill_time_table_host(t.ctypes.data_as(C.c_void_p), 600, 0.030)
    assert np.array_equal(t, mo.t_table(599))


def test_argument_errors_need_no_gpu():
    from mr_rl_b200 import _lib
    lib = _lib.load()
    p = _lib.default_params()
    assert lib.mr_env_step(None, 4, 0, C.byref(p), None, None, None, None, None) == -1
    assert b"null state" in lib.mr_last_error()
    st = _lib.EnvState()
    assert lib.mr_env_step(C.byref(st), -1, 0, C.byref(p), None, None, None, None, None) == -1
    assert lib.mr_gp_predict(None, None, 1, None, None, None, 0, None) == -1
    assert lib.mr_actor_forward(None, None, 0, 1, 0, (C.c_double * 2)(20.0, 6.28), None, None) == -1


def test_no_cpu_fallback_exists():
    """The product must fail loudly without a CUDA device / library — never route through a CPU path."""
    import mr_rl_b200
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            mr_rl_b200.VecMREnv(4)
    with pytest.raises(mr_rl_b200.MRLibraryError):
        mr_rl_b200.VecMREnv(4, device="cpu")
    src = "".join(open(os.path.join(ROOT, "mr_rl_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "mr_rl_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("mr_oracle", "").lower() or "import oracle" not in src     # product never imports the oracle
    assert "from oracle" not in src and "import oracle" not in src


def test_spaces_and_shard_ranges():
    from mr_rl_b200 import Box, shard_range
    b = Box(low=np.array([100, 100]), high=np.array([120, 120]), seed=0)
    s = b.sample()
    assert s.dtype == np.float32 and b.contains(s) and not b.contains(np.array([99.0, 100.0]))
    for total, world in [(1 << 20, 8), (1000, 3), (7, 8), (0, 2)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1


def _dist_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from mr_rl_b200 import dist as mrd
    r, w, _ = mrd.init_from_env(backend="gloo")
    start, count = mrd.shard_range(1000, r, w)
    # every rank accumulates statistics for its own shard; merged statistics must equal the serial answer
    stats = torch.zeros(8, dtype=torch.float64)
    lengths = (np.arange(start, start + count) % 51) + 1
    stats[0] = count; stats[1] = float(lengths.sum()); stats[2] = 10.0 * lengths.sum(); stats[5] = count; stats[6] = float(lengths.sum())
    merged = mrd.merge_stats(stats)
    q.put((r, start, count, merged, stats.tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert [r[1:3] for r in res] == [(0, 500), (500, 500)]
    all_len = (np.arange(1000) % 51) + 1
    for r in res:
        m = r[3]
        assert m["episodes"] == 1000 and m["sum_length"] == float(all_len.sum()) and m["sum_reward"] == 10.0 * all_len.sum()
        assert abs(m["mean_episode_length"] - all_len.mean()) < 1e-12
    assert res[0][4][0] == 500      # local accumulators are not overwritten by the merge
