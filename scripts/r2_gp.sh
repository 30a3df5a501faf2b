#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_gp_fit.py tests/test_learn_flow.py tests/test_learn_2d.py -m gpu -q > gpurun_out/gp_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gp_pytest.log
python tools/gpfitbench.py > gpurun_out/gp_fit.log 2>&1
CMD="python tools/gpfit_once.py"
$CMD > gpurun_out/gp_once_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/gp_launches.csv $CMD > gpurun_out/gp_ncu.log 2>&1
tail -3 gpurun_out/gp_pytest.log; cat gpurun_out/gp_fit.log | cut -c1-200; python tools/launch_agg.py gpurun_out/gp_launches.csv | head -8
