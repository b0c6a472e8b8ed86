#!/bin/bash
# quick pass: batched parity (both weight paths) + 8B shape, batched-step probe, launch list, full capture of the attention kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2j}
mkdir -p gpurun_out
T() { ( time timeout "$1" python -m pytest "${@:2}" -q -p no:cacheprovider --durations=4 ) ; }
T 400 tests/test_gpu_batch.py -x > gpurun_out/${TAG}_t_batch.log 2>&1; tail -5 gpurun_out/${TAG}_t_batch.log
if ! grep -q " passed" gpurun_out/${TAG}_t_batch.log || grep -q "failed" gpurun_out/${TAG}_t_batch.log; then echo "batched path failed: stopping"; exit 1; fi
T 400 tests/test_gpu_8b_shape.py > gpurun_out/${TAG}_t_8b.log 2>&1; tail -4 gpurun_out/${TAG}_t_8b.log
for B in 8 16 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/${TAG}_probe_q.log 2>&1
grep "^{" gpurun_out/${TAG}_probe_q.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_q.csv python tools/batch_probe.py 32 576 1 2 > gpurun_out/${TAG}_ncu_q.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_q.csv 261
if [ -z "$NO_FULL" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:batch_attn_mma -s 40 -c 1 -o gpurun_out/${TAG}_battn_full python tools/batch_probe.py 32 576 1 2 > gpurun_out/${TAG}_ncu_full_attn.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
fi
