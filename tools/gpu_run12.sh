#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log

timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_quick.log",):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)
            if "type" in d: print(d["type"], d["rows"], d["cols"], d["GBps"], d["frac_of_measured_peak"], d["cfg"])
            elif "ctx" in d: print(d["ctx"], d["ms_per_token"], d["tok_s"], d["cfg"])
        else: print(l.strip()[:200])
PY
