#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/m_log.txt
for v in "" "_l1simt"; do
  echo "== lib${v}" >> gpurun_out/m_log.txt
  MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so python -m pytest tests/test_gpu_gp_actor.py -m gpu -q -k "actor" 2>&1 | tail -2 >> gpurun_out/m_log.txt
  MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so python tools/actorbench.py --paths default --launches 4 >> gpurun_out/m_log.txt 2>&1
done
cat gpurun_out/m_log.txt
