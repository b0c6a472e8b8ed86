#!/bin/bash
# parity of the batched paths, then the in-kernel timeline of the quantised GEMMs in a replayed step, then the probe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2l}
mkdir -p gpurun_out
T() { ( time timeout "$1" python -m pytest "${@:2}" -q -p no:cacheprovider --durations=4 ) ; }
T 600 tests/test_gpu_batch.py tests/test_gpu_8b_shape.py -x > gpurun_out/${TAG}_t_batch.log 2>&1; tail -5 gpurun_out/${TAG}_t_batch.log
if ! grep -q " passed" gpurun_out/${TAG}_t_batch.log || grep -q "failed" gpurun_out/${TAG}_t_batch.log; then echo "batched path failed: stopping"; exit 1; fi
timeout 300 python tools/qgemm_trace.py 32 576 > gpurun_out/${TAG}_qgemm_trace.txt 2>&1; cat gpurun_out/${TAG}_qgemm_trace.txt | tail -22
for B in 8 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/${TAG}_probe_q.log 2>&1
grep "^{" gpurun_out/${TAG}_probe_q.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_q.csv python tools/batch_probe.py 32 576 1 2 > gpurun_out/${TAG}_ncu_q.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_q.csv 261
