#!/bin/bash
# evidence run: full GPU parity suite, bench (native + reference arm), launch list of the bench command, full ncu of one
# layer + lm_head, device timeline, prefill launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json | cut -c1-400
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 300 python tools/perop_trace.py 576 > gpurun_out/perop_trace_576.log 2>&1
timeout 300 python tools/sample_probe.py > gpurun_out/sample_probe.log 2>&1
grep -h '^{' gpurun_out/sample_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -s 342 -c 5 -o gpurun_out/layer_full -f python tools/profile_decode.py 3 576 > gpurun_out/prof_layer.log 2>&1
echo "ncu layer rc=$?" >> gpurun_out/prof_layer.log
timeout 900 ncu --set full --clock-control none -k regex:"gemv_kernel" -s 515 -c 1 -o gpurun_out/head_full -f python tools/profile_decode.py 3 576 > gpurun_out/prof_head.log 2>&1
echo "ncu head rc=$?" >> gpurun_out/prof_head.log
# prefill: launch list of one 512-token prefill (first engine only: tcgen05 path), then full metrics of the gate/up GEMM
GL_PREFILL_TC5=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_prefill.csv python tools/prefill_probe.py > gpurun_out/prefill_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc5" -s 10 -c 4 -o gpurun_out/tc5_full -f python tools/prefill_probe.py > gpurun_out/prof_tc5.log 2>&1
echo "ncu tc5 rc=$?" >> gpurun_out/prof_tc5.log
timeout 300 python tools/prefill_probe.py > gpurun_out/prefill_probe.log 2>&1
cat gpurun_out/prefill_probe.log
tail -n 2 gpurun_out/prof_layer.log gpurun_out/prof_head.log gpurun_out/prof_tc5.log
