#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -x -q -k "long_context" > gpurun_out/pytest_long.log 2>&1
echo "pytest long rc=$?" >> gpurun_out/pytest_long.log
tail -n 5 gpurun_out/pytest_long.log
timeout 400 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
grep -h '^{' gpurun_out/bench_quick.log | sort -u | cut -c1-200; tail -n 2 gpurun_out/bench_quick.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
