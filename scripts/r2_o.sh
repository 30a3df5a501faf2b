#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/o_log.txt
for v in "" "_s4" "_s2"; do
  echo "== lib${v}" >> gpurun_out/o_log.txt
  MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so python tools/stepbench.py --steps 400 2>&1 | grep float >> gpurun_out/o_log.txt
done
cat gpurun_out/o_log.txt
