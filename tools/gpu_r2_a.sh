#!/bin/bash
# round 2, first GPU pass: the whole -m gpu suite (new: batching, 8B-shape parity, cancel, rope variants), then the bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt; df -h /dev/shm >> gpurun_out/r2a_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/r2a_pytest.log 2>&1
tail -40 gpurun_out/r2a_pytest.log
( time timeout 600 python bench.py --steps 4 --warmup 3 ) > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
tail -3 gpurun_out/r2a_bench_c2.err; cut -c1-600 gpurun_out/r2a_bench_c2.json
( time timeout 900 python bench.py --workload config3 --steps 2 --warmup 1 --batch-weights 1 ) > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
tail -5 gpurun_out/r2a_bench_c3.err; cut -c1-900 gpurun_out/r2a_bench_c3.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -3 gpurun_out/r2a_bench_ref.err; cut -c1-600 gpurun_out/r2a_bench_ref.json
