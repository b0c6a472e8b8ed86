#!/bin/bash
# round 2, second GPU pass: new library (batched GEMM on quantised weights, 128x256 prefill tiles, packed embeddings).
# Order: cheap targeted tests first (each group in its own process: a trapped kernel poisons only its group), then config3 with a watchdog.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T() { ( time timeout "$1" python -m pytest "${@:2}" -q -p no:cacheprovider --durations=12 ) ; }
T 600 tests/test_gpu_batch.py -k "not 2]" > gpurun_out/r2b_t_batch_m1.log 2>&1; tail -25 gpurun_out/r2b_t_batch_m1.log
T 600 tests/test_gpu_batch.py -k "2]" > gpurun_out/r2b_t_batch_m2.log 2>&1; tail -30 gpurun_out/r2b_t_batch_m2.log
T 900 tests/test_gpu_decode.py tests/test_gpu_service.py > gpurun_out/r2b_t_decode.log 2>&1; tail -25 gpurun_out/r2b_t_decode.log
( time GL_BENCH_WATCHDOG_S=100 timeout 420 python bench.py --workload config3 --steps 2 --warmup 1 --batch-weights 1 --no-cpu ) > gpurun_out/r2b_bench_c3_w16.json 2> gpurun_out/r2b_bench_c3_w16.err
tail -40 gpurun_out/r2b_bench_c3_w16.err; cut -c1-1500 gpurun_out/r2b_bench_c3_w16.json
( time GL_BENCH_WATCHDOG_S=100 timeout 420 python bench.py --workload config3 --steps 2 --warmup 1 --batch-weights 2 --no-cpu ) > gpurun_out/r2b_bench_c3_q.json 2> gpurun_out/r2b_bench_c3_q.err
tail -40 gpurun_out/r2b_bench_c3_q.err; cut -c1-1500 gpurun_out/r2b_bench_c3_q.json
T 600 tests/test_gpu_8b_shape.py > gpurun_out/r2b_t_8b.log 2>&1; tail -25 gpurun_out/r2b_t_8b.log
