#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests -m gpu -v > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
python tools/actorbench.py --paths default tf32 simt --launches 2 > gpurun_out/e_actor.log 2>&1
CMD="python tools/actorbench.py --paths default --launches 1 --k 20"
$CMD > gpurun_out/e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_actor16 $CMD > gpurun_out/e_ncu.log 2>&1
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/e_pytest.log | tail -20; cat gpurun_out/e_actor.log; tail -3 gpurun_out/e_ncu.log
