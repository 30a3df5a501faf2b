#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -u -X faulthandler -m pytest tests -m gpu -q -x > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_1gpu.json 2> gpurun_out/z_bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err
tail -4 gpurun_out/z_pytest.log; tail -1 gpurun_out/z_smoke.log; cut -c1-300 gpurun_out/z_bench_1gpu.json; cut -c1-300 gpurun_out/z_bench_ref.json; tail -3 gpurun_out/z_bench_1gpu.err
