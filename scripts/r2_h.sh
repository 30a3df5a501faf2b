#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1
python tools/actorbench.py --paths default --launches 4 > gpurun_out/h_actor.log 2>&1
python tools/actorbench.py --paths default --launches 10 --k 20 >> gpurun_out/h_actor.log 2>&1
tail -6 gpurun_out/h_pytest.log; tail -1 gpurun_out/h_smoke.log; cat gpurun_out/h_actor.log
