#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/san_plain.log 2>&1 &&
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/san_memcheck.log
tail -3 gpurun_out/san_plain.log; tail -12 gpurun_out/san_memcheck.log
