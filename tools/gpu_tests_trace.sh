#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or service or long or gemv" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 300 python tools/perop_trace.py 576 > gpurun_out/perop_trace_576.log 2>&1
grep -E "ATTN tail|^O |^ATTN|ms/token" gpurun_out/perop_trace_576.log
timeout 500 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
grep -h '^{' gpurun_out/bench_quick.log | sort -u | cut -c1-200
