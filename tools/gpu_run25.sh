#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "prefill or decode or service or smoke" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 600 python tools/prefill_probe.py > gpurun_out/prefill_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/prefill_probe.log
tail -n 6 gpurun_out/prefill_probe.log
