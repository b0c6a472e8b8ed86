#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_decode.csv python tools/profile_decode.py 4 > gpurun_out/prof1.log 2>&1
echo "ncu1 rc=$?" >> gpurun_out/prof1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemv_kernel|attn_decode" -s 326 -c 7 -o gpurun_out/perop_full -f python tools/profile_decode.py 2 > gpurun_out/prof2.log 2>&1
echo "ncu2 rc=$?" >> gpurun_out/prof2.log
tail -2 gpurun_out/prof1.log gpurun_out/prof2.log
