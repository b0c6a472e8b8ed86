#!/bin/bash
# first GPU bring-up: parity tests, GEMV micro-benchmarks, 8B decode probe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then
  GL_PDL=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_nopdl.log 2>&1
  echo "pytest(nopdl) rc=$?" >> gpurun_out/pytest_gpu_nopdl.log
  GL_PDL=0 GL_GRAPH=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_nograph.log 2>&1
  echo "pytest(nograph) rc=$?" >> gpurun_out/pytest_gpu_nograph.log
  GL_PDL=0 GL_GRAPH=0 GL_FUSE=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_nofuse.log 2>&1
  echo "pytest(nofuse) rc=$?" >> gpurun_out/pytest_gpu_nofuse.log
fi
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1
echo "microbench rc=$?" >> gpurun_out/microbench.log
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/microbench.log; tail -3 gpurun_out/bench_quick.log
