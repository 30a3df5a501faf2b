#!/usr/bin/env bash
# round 2, call A: GPU tests after the ADVICE fixes, step-kernel variants (plain TMA vs warp-specialised, Philox-10 vs -7),
# a short bench line, one ncu capture of the warp-specialised kernel
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/a_pytest.log
python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1
for v in "" "_p7"; do
  for path in tma ws; do
    echo "== lib${v} path=${path}" >> gpurun_out/a_stepbench.log
    MR_LIB_PATH=$PWD/mr_rl_b200/_lib/libmr_rl_b200${v}.so MR_STEP_PATH=$path python tools/stepbench.py --steps 400 >> gpurun_out/a_stepbench.log 2>&1
  done
done
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
CMD="python tools/stepbench.py --steps 30"
$CMD > gpurun_out/a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_step_tma_ws -s 20 -c 1 -f -o gpurun_out/prof_ws $CMD > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_pytest.log; tail -2 gpurun_out/a_smoke.log; cat gpurun_out/a_stepbench.log; cat gpurun_out/a_bench.json | cut -c1-600
