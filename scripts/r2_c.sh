#!/usr/bin/env bash
# round 2, call C: verbose GPU tests with the fault handler on (a silent crash in call B), smoke
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests -m gpu -x -v > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1
tail -30 gpurun_out/c_pytest.log; tail -2 gpurun_out/c_smoke.log
