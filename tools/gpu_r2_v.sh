#!/bin/bash
# CTA-pair GEMM (cta_group::2): the whole GPU suite, A/B against the one-CTA kernel at 512 / 2048 prompt tokens, bench lines of configs 4, 3, 5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2v}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=6 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -14 gpurun_out/${TAG}_pytest_gpu.log
if grep -q "failed\|error" gpurun_out/${TAG}_pytest_gpu.log; then
  echo "== same suite without CTA pairs"
  ( time GL_TC5_PAIR=0 timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu_nopair.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_nopair.log
  export GL_TC5_PAIR=0
  echo "== continuing with GL_TC5_PAIR=0"
fi
PROBE_VARIANTS="GL_TC5_PAIR=1;GL_TC5_PAIR=0" timeout 400 python tools/prefill_attn_probe.py 512 2048 > gpurun_out/${TAG}_prefill_pair_probe.log 2>&1; grep "^{" gpurun_out/${TAG}_prefill_pair_probe.log || tail -5 gpurun_out/${TAG}_prefill_pair_probe.log
PROBE_VARIANTS="GL_TC5_PAIR=1" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_prefill2048.csv python tools/prefill_attn_probe.py 2048 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_prefill2048.csv 100000 2>/dev/null | head -14
for W in config4 config3 config5; do
  ( time GL_BENCH_WATCHDOG_S=300 timeout 600 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  tail -2 gpurun_out/${TAG}_bench_$W.err; cut -c1-330 gpurun_out/${TAG}_bench_$W.json; echo
done
