#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "gemv or decode" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
grep -h '^{' gpurun_out/bench_quick.log | sort -u | cut -c1-200; tail -n 2 gpurun_out/bench_quick.log
