#!/bin/bash
# sampler parity tests + cost of the samplers inside a request
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sampler or sampled or generate" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 400 python tools/sample_probe.py > gpurun_out/sample_probe.log 2>&1
grep -h '^{' gpurun_out/sample_probe.log
