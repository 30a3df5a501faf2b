#!/usr/bin/env bash
# round 2, call B: full GPU tests, reset-draw debug dump, bench line, per-rank sizes of the strong-scaling run on one GPU
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/b_pytest.log
python tools/dbg_reset_draws.py > gpurun_out/b_dbg.log 2>&1
python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
for n in 131072 262144 524288; do
  for l in graph direct; do
    python bench.py --steps 20 --warmup 5 --envs $n --launch $l --no-extras --cpu-seconds 0.2 --e2e-steps 3 2>> gpurun_out/b_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print($n,'$l',d['ms_per_step']*1e3,'us/step',d['value']/1e9,'Genv-steps/s', d['repeat_ms_per_step'])" >> gpurun_out/b_sizes.log 2>&1
  done
done
tail -5 gpurun_out/b_pytest.log; tail -2 gpurun_out/b_smoke.log; cat gpurun_out/b_sizes.log; cut -c1-1500 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
