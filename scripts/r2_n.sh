#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/n_log.txt
for pad in 0 16 256 4112; do
  for path in tma tmap; do
    echo "== row pad $pad path $path" >> gpurun_out/n_log.txt
    MR_ROW_PAD=$pad MR_STEP_PATH=$path python tools/stepbench.py --steps 400 2>&1 | grep "float64" >> gpurun_out/n_log.txt
  done
done
cat gpurun_out/n_log.txt
