#!/bin/bash
# bufferbloat probe: does the prologue / barrier time of the persistent kernel scale with bytes in flight?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
for cfg in "2 36" "3 36" "5 36" "2 72" "6 18" "4 9"; do
  set -- $cfg
  echo "=== SLOTS=$1 SLOT_KB=$2" >> gpurun_out/mega_trace_inflight.log
  GL_MEGA=1 GL_MEGA_SLOTS=$1 GL_MEGA_SLOT_KB=$2 timeout 300 python tools/mega_trace.py 576 >> gpurun_out/mega_trace_inflight.log 2>&1
done
cat gpurun_out/mega_trace_inflight.log
