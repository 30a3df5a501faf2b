#!/usr/bin/env bash
# round 2 evidence on one B200: bench line, ncu launch list of the bench command, ncu full captures of the step kernels
# (generated noise, noise-free, table noise) and of the actor rollout.  Summaries: python tools/ncu_summary.py <rep>.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/p_bench_1gpu.json 2> gpurun_out/p_bench_1gpu.err
python tools/tablebench.py > gpurun_out/p_table.log 2>&1
CMD="python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 0.5 --e2e-steps 3 --preroll-s 0.002"
$CMD > gpurun_out/p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/p_launches.csv $CMD > gpurun_out/p_ncu_launches.log 2>&1
$CMD > gpurun_out/p_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 12 -c 1 -f -o gpurun_out/prof_step_sigma1 $CMD > gpurun_out/p_ncu_s1.log 2>&1
$CMD --sigma 0 > gpurun_out/p_plain0.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 12 -c 1 -f -o gpurun_out/prof_step_sigma0 $CMD --sigma 0 > gpurun_out/p_ncu_s0.log 2>&1
TCMD="python tools/tablebench.py --steps 8 --paths tmap"
$TCMD > gpurun_out/p_plaint.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 10 -c 1 -f -o gpurun_out/prof_step_table $TCMD > gpurun_out/p_ncu_t.log 2>&1
ACMD="python tools/actorbench.py --paths default --launches 1 --k 20"
$ACMD > gpurun_out/p_plaina.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_actor16 $ACMD > gpurun_out/p_ncu_a.log 2>&1
cat gpurun_out/p_table.log; cut -c1-400 gpurun_out/p_bench_1gpu.json; ls -la gpurun_out/*.ncu-rep
