#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
for ctx in 1 576; do
timeout 300 python tools/perop_trace.py $ctx > gpurun_out/perop_trace_$ctx.log 2>&1
cat gpurun_out/perop_trace_$ctx.log
done
