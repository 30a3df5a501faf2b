#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/j_stepbench.log
for v in "" "_nopf"; do
  echo "== lib${v}" >> gpurun_out/j_stepbench.log
  MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so python tools/stepbench.py --steps 400 >> gpurun_out/j_stepbench.log 2>&1
  MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2 --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench', round(d['ms_per_step']*1e3,2),'us/step frac', round(d['roofline']['frac'],3), [round(x*1e3,2) for x in d['repeat_ms_per_step']])" >> gpurun_out/j_stepbench.log
done
python -u -m pytest tests/test_gpu_env.py -m gpu -q -k "variants or trajectory_parity or statistics" > gpurun_out/j_pytest.log 2>&1
cat gpurun_out/j_stepbench.log; tail -3 gpurun_out/j_pytest.log
