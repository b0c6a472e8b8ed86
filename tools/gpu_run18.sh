#!/bin/bash
# re-entry baseline: parity tests, decode probe over configs, GEMV microbench, launch list + full ncu of one layer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
echo "bench_quick rc=$?" >> gpurun_out/bench_quick.log
grep -h '^{' gpurun_out/bench_quick.log | cut -c1-200
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1
echo "microbench rc=$?" >> gpurun_out/microbench.log
grep -h '^{' gpurun_out/microbench.log | cut -c1-200
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ctx576.csv python tools/profile_decode.py 3 576 > gpurun_out/prof_ctx576.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemv_kernel|attn_decode" -s 326 -c 7 -o gpurun_out/perop_full -f python tools/profile_decode.py 3 576 > gpurun_out/prof2.log 2>&1
echo "ncu2 rc=$?" >> gpurun_out/prof2.log
tail -2 gpurun_out/prof_ctx576.log gpurun_out/prof2.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
