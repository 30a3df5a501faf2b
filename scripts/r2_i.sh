#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/i_sizes.log
for n in 131072 262144; do
  for p in tma vec scalar; do
    MR_STEP_PATH=$p python bench.py --steps 20 --warmup 5 --envs $n --launch graph --no-extras --cpu-seconds 0.2 --e2e-steps 3 2>> gpurun_out/i_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print($n,'$p',round(d['ms_per_step']*1e3,2),'us/step',round(d['value']/1e9,2),'Genv-steps/s', [round(x*1e3,2) for x in d['repeat_ms_per_step']])" >> gpurun_out/i_sizes.log 2>&1
  done
done
cat gpurun_out/i_sizes.log
