#!/bin/bash
# 2-GPU check of the bench contract (torchrun, one rank per GPU, nccl barrier + max-reduce of the times)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu2.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?" >> gpurun_out/bench_2gpu.err
cat gpurun_out/bench_2gpu.json | cut -c1-700
tail -n 5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
echo "rc=$?" >> gpurun_out/bench_ref_2gpu.err
cat gpurun_out/bench_ref_2gpu.json | cut -c1-500
tail -n 3 gpurun_out/bench_ref_2gpu.err
