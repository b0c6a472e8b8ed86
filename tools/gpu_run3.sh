#!/bin/bash
# run 3: persistent decode kernel + new GEMV core
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then
  GL_MEGA=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_nomega.log 2>&1
  echo "pytest(nomega) rc=$?" >> gpurun_out/pytest_gpu_nomega.log
fi
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1
echo "microbench rc=$?" >> gpurun_out/microbench.log
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
tail -4 gpurun_out/pytest_gpu.log
tail -3 gpurun_out/bench_quick.log
