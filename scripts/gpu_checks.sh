#!/usr/bin/env bash
# Regenerate the round's GPU evidence on a B200 box (run from the repo root, e.g. through gpurun):
#   tests -> bench line -> ncu launch list -> ncu full captures of the step kernel and of the other paths.
# Outputs land in gpurun_out/; summaries are produced afterwards (no GPU needed) with
#   python tools/ncu_summary.py gpurun_out/prof_step_sigma1.ncu-rep > profiles/rNN_ncu_step_sigma1.txt
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
CMD="python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 0.5 --e2e-steps 3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 8 -c 1 -f -o gpurun_out/prof_step_sigma1 $CMD > gpurun_out/ncu_s1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_tma -s 8 -c 1 -f -o gpurun_out/prof_step_sigma0 $CMD --sigma 0 > gpurun_out/ncu_s0.log 2>&1
PCMD="python tools/profile_paths.py"
$PCMD > gpurun_out/paths_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"env_rollout_kernel|gp_var_kernel|gp_kq_mean|gp_posterior_spectral" -c 8 -f -o gpurun_out/prof_paths $PCMD > gpurun_out/ncu_paths.log 2>&1
NCMD="python tools/next_rows_once.py"
$NCMD > gpurun_out/next_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/next_rows_launches.csv $NCMD > gpurun_out/ncu_next_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ddpg_update|gp_lml_grad" -c 2 -f -o gpurun_out/prof_next_rows $NCMD > gpurun_out/ncu_next.log 2>&1
python tools/gpfitbench.py > gpurun_out/gpfit.log 2>&1
python tools/ddpgbench.py > gpurun_out/ddpg.log 2>&1
python tools/e2ebench.py > gpurun_out/e2e.log 2>&1
ls -la gpurun_out
