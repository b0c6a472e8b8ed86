#!/bin/bash
# short A/B run: synthetic Llama-3-8B q4_K_M, per-op timeline at ctx 576, tools/bench_quick.py configurations
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 500 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
grep -h '^{' gpurun_out/bench_quick.log | sort -u | cut -c1-200
