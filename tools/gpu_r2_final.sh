#!/bin/bash
# the tree as the driver will run it at round end: smoke(), the full GPU suite (timed), the default bench line (config 2, both legs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2final}
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 4 gpurun_out/${TAG}_smoke.log
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=5 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -12 gpurun_out/${TAG}_pytest_gpu.log
if grep -q "failed\|error" gpurun_out/${TAG}_pytest_gpu.log; then
  echo "== the suite again with the mma.sync prompt attention"
  ( time GL_PREFILL_ATTN_TC5=0 timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu_notc5attn.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu_notc5attn.log
  export GL_PREFILL_ATTN_TC5=0
fi
( time GL_BENCH_WATCHDOG_S=300 timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err; cut -c1-1500 gpurun_out/${TAG}_bench.json
PROBE_VARIANTS="GL_PREFILL_ATTN_TC5=1;GL_PREFILL_ATTN_TC5=0" timeout 200 python tools/prefill_attn_probe.py 512 2048 > gpurun_out/${TAG}_prefill_attn_tc5_probe.log 2>&1; grep "^{" gpurun_out/${TAG}_prefill_attn_tc5_probe.log | cut -c1-250
for W in config4 config3; do
  ( time GL_BENCH_WATCHDOG_S=200 timeout 400 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  tail -1 gpurun_out/${TAG}_bench_$W.err; cut -c1-330 gpurun_out/${TAG}_bench_$W.json; echo
done
