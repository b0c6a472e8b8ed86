#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
rm -f gpurun_out/mega_trace.log
for nw in 8 12; do
  echo "=== GL_WARPS=$nw" >> gpurun_out/mega_trace.log
  GL_MEGA=1 GL_WARPS=$nw timeout 300 python tools/mega_trace.py 576 >> gpurun_out/mega_trace.log 2>&1
done
cat gpurun_out/mega_trace.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ctx576.csv python tools/profile_decode.py 3 576 > gpurun_out/prof_ctx576.log 2>&1
tail -n 2 gpurun_out/prof_ctx576.log
