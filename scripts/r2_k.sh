#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests/test_gpu_env.py tests/test_c_client.py -m gpu -q -x > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_pytest.log
rm -f gpurun_out/k_stepbench.log
for path in tma tmap; do
  echo "== path=${path}" >> gpurun_out/k_stepbench.log
  MR_STEP_PATH=$path python tools/stepbench.py --steps 400 >> gpurun_out/k_stepbench.log 2>&1
  MR_STEP_PATH=$path python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2 --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench', round(d['ms_per_step']*1e3,2),'us/step frac', round(d['roofline']['frac'],3), [round(x*1e3,2) for x in d['repeat_ms_per_step']], 'e2e', round(d['e2e']['ms_per_step'],3))" >> gpurun_out/k_stepbench.log
  MR_STEP_PATH=$path python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2 --e2e-steps 3 --envs 131072 --launch graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench 131072 graph', round(d['ms_per_step']*1e3,2),'us/step')" >> gpurun_out/k_stepbench.log
done
python tools/tablebench.py >> gpurun_out/k_stepbench.log 2>&1
tail -5 gpurun_out/k_pytest.log; cat gpurun_out/k_stepbench.log
