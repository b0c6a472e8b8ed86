#!/bin/bash
# run 2: batched prefill parity + bench line + ncu evidence for the decode GEMV
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
# launch list: 3 warm-up + 4 timed decode steps = 7 x 163 launches (model file is reused from bench)
cp /dev/shm/gridllm_llama3_8b_q4km_synth_seed1234.gguf /dev/shm/prof_llama3_8b.gguf 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_decode.csv python tools/profile_decode.py 4 > gpurun_out/prof1.log 2>&1
echo "ncu1 rc=$?" >> gpurun_out/prof1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 330 -c 6 -o gpurun_out/gemv_full -f python tools/profile_decode.py 2 > gpurun_out/prof2.log 2>&1
echo "ncu2 rc=$?" >> gpurun_out/prof2.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 1500; tail -2 gpurun_out/prof1.log gpurun_out/prof2.log
