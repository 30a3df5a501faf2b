#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests/test_gpu_gp_actor.py -m gpu -q > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
python tools/actorbench.py --paths default tf32 --launches 4 > gpurun_out/f_actor.log 2>&1
python tools/actorbench.py --paths default --launches 4 --sigma 0 >> gpurun_out/f_actor.log 2>&1
tail -4 gpurun_out/f_pytest.log; cat gpurun_out/f_actor.log
