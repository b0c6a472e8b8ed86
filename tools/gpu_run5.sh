#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
for cfg in "GL_WARPS=8" "GL_WARPS=8 GL_ACT_BITS=8" "GL_WARPS=12 GL_MEGA_SLOT_KB=54"; do
  echo "=== $cfg" >> gpurun_out/mega_trace.log
  env $cfg timeout 300 python tools/mega_trace.py 576 >> gpurun_out/mega_trace.log 2>&1
  env $cfg timeout 300 python tools/mega_trace.py 1 >> gpurun_out/mega_trace.log 2>&1
done
GL_WARPS=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 1 -o gpurun_out/mega_full -f python tools/profile_decode.py 2 > gpurun_out/prof_mega.log 2>&1
echo "ncu rc=$?" >> gpurun_out/prof_mega.log
tail -40 gpurun_out/mega_trace.log
