#!/bin/bash
# quick pass: quantised-path parity, then the batched-step probe with and without programmatic dependent launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2h}
mkdir -p gpurun_out
T() { ( time timeout "$1" python -m pytest "${@:2}" -q -p no:cacheprovider --durations=4 ) ; }
T 300 tests/test_gpu_batch.py -k "2]" -x > gpurun_out/${TAG}_t_batch_m2.log 2>&1; tail -5 gpurun_out/${TAG}_t_batch_m2.log
if ! grep -q " passed" gpurun_out/${TAG}_t_batch_m2.log || grep -q "failed" gpurun_out/${TAG}_t_batch_m2.log; then echo "quantised path failed: stopping"; exit 1; fi
for B in 8 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/${TAG}_probe_q.log 2>&1
for B in 8 32 64; do GL_BATCH_PDL=0 timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/${TAG}_probe_q_nopdl.log 2>&1
grep "^{" gpurun_out/${TAG}_probe_q.log; echo "-- GL_BATCH_PDL=0"; grep "^{" gpurun_out/${TAG}_probe_q_nopdl.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_q.csv python tools/batch_probe.py 32 576 1 2 > gpurun_out/${TAG}_ncu_q.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_q.csv 261
