#!/bin/bash
# v3 GEMV core (warp-owned items, FIFO ring of small slots): parity, microbench, decode probe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1
echo "microbench rc=$?" >> gpurun_out/microbench.log
grep -h '^{' gpurun_out/microbench.log | cut -c1-200; tail -3 gpurun_out/microbench.log
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
grep -h '^{' gpurun_out/bench_quick.log | cut -c1-220; tail -3 gpurun_out/bench_quick.log
