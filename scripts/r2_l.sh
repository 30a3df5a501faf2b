#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/l_log.txt
for l2 in 0 2 3; do
  echo "== tmap L2 promotion $l2" >> gpurun_out/l_log.txt
  MR_TMAP_L2=$l2 python tools/stepbench.py --steps 400 >> gpurun_out/l_log.txt 2>&1
done
for path in tma tmap; do
  MR_STEP_PATH=$path python bench.py --steps 100 --warmup 5 --cpu-seconds 0.2 --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$path bench sigma1', round(d['ms_per_step']*1e3,2), 'sigma0', round(d['noise_free_sigma0']['ms_per_step']*1e3,2), 'sp', round(d['with_state_prime_rows']['ms_per_step']*1e3,2), 'table', round(d['parity_mode_table_noise']['ms_per_step']*1e3,2))" >> gpurun_out/l_log.txt
done
cat gpurun_out/l_log.txt
