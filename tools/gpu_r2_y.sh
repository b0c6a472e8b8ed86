#!/bin/bash
# two GPUs: N engines in ONE process behind the scheduler rules (the north_star's multi-GPU shape), and the bench contract under torchrun
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2y}
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_service.py -x -q -m gpu -p no:cacheprovider -k "in_process_workers or shard" -rs ) > gpurun_out/${TAG}_pytest_2gpu.log 2>&1; tail -8 gpurun_out/${TAG}_pytest_2gpu.log
( time GL_BENCH_WATCHDOG_S=300 timeout 600 python bench.py --workload config3 --gpus 2 --steps 2 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c3_inproc2.json 2> gpurun_out/${TAG}_bench_c3_inproc2.err
tail -2 gpurun_out/${TAG}_bench_c3_inproc2.err; cut -c1-400 gpurun_out/${TAG}_bench_c3_inproc2.json; echo
( time GL_BENCH_WATCHDOG_S=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 1 --no-cpu ) > gpurun_out/${TAG}_bench_c2_torchrun2.json 2> gpurun_out/${TAG}_bench_c2_torchrun2.err
tail -2 gpurun_out/${TAG}_bench_c2_torchrun2.err; grep "^{" gpurun_out/${TAG}_bench_c2_torchrun2.json | cut -c1-400
