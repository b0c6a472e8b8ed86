#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
grep -E '"ctx": 576|"ctx": 1,' gpurun_out/bench_quick.log
