#!/bin/bash
# the WHOLE GPU suite (timed: the driver's limit is 1200 s), the quantised-GEMM timeline, the probe, and the bench lines of configs 3, 4, 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2o}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=12 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -22 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/qgemm_trace.py 32 576 > gpurun_out/${TAG}_qgemm_trace.txt 2>&1; tail -7 gpurun_out/${TAG}_qgemm_trace.txt
for B in 8 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/${TAG}_probe_q.log 2>&1
grep "^{" gpurun_out/${TAG}_probe_q.log
( time GL_BENCH_WATCHDOG_S=200 timeout 420 python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
tail -4 gpurun_out/${TAG}_bench_c3.err; cut -c1-400 gpurun_out/${TAG}_bench_c3.json
( time GL_BENCH_WATCHDOG_S=200 timeout 420 python bench.py --workload config4 --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
tail -3 gpurun_out/${TAG}_bench_c4.err; cut -c1-400 gpurun_out/${TAG}_bench_c4.json
( time GL_BENCH_WATCHDOG_S=200 timeout 420 python bench.py --steps 3 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
tail -3 gpurun_out/${TAG}_bench_c2.err; cut -c1-400 gpurun_out/${TAG}_bench_c2.json
