#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests -m gpu -v > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
python __graft_entry__.py smoke > gpurun_out/d_smoke.log 2>&1
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/d_pytest.log | tail -30; tail -2 gpurun_out/d_smoke.log
