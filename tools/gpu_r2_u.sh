#!/bin/bash
# fused prompt attention: the whole GPU suite on it, A/B against the three-launch path, launch list of a 2048-token prompt pass, config-4 line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2u}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -16 gpurun_out/${TAG}_pytest_gpu.log
if grep -q "failed\|error" gpurun_out/${TAG}_pytest_gpu.log; then
  echo "== same suite on the three-launch path"
  ( time GL_PREFILL_FLASH=0 timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu_noflash.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_noflash.log
fi
timeout 400 python tools/prefill_attn_probe.py 512 2048 > gpurun_out/${TAG}_prefill_attn_probe.log 2>&1; grep "^{" gpurun_out/${TAG}_prefill_attn_probe.log || tail -5 gpurun_out/${TAG}_prefill_attn_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_prefill2048.csv python tools/prefill_attn_probe.py 2048 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_prefill2048.csv 100000 | head -16
( time GL_BENCH_WATCHDOG_S=300 timeout 600 python bench.py --workload config4 --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
tail -3 gpurun_out/${TAG}_bench_c4.err; cut -c1-600 gpurun_out/${TAG}_bench_c4.json
