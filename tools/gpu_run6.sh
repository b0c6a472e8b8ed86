#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
rm -f gpurun_out/mega_trace.log
for cfg in "GL_WARPS=8" "GL_WARPS=8 GL_ACT_BITS=8"; do
  echo "=== $cfg" >> gpurun_out/mega_trace.log
  env $cfg timeout 300 python tools/mega_trace.py 576 >> gpurun_out/mega_trace.log 2>&1
done
tail -4 gpurun_out/pytest_gpu.log
cat gpurun_out/mega_trace.log

grep -E '"GL_WARPS": "8"|ACT_BITS' gpurun_out/microbench.log | head -12
