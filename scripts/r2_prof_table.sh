#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TCMD="python tools/tablebench.py --steps 8 --paths tmap"
$TCMD > gpurun_out/p_plaint.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 10 -c 1 -f -o gpurun_out/prof_step_table $TCMD > gpurun_out/p_ncu_t.log 2>&1
tail -2 gpurun_out/p_ncu_t.log
