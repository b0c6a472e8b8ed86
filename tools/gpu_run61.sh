#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_|embed_kernel" -s 790 -c 45 --csv --log-file gpurun_out/launches_sampler.csv python tools/sample_probe.py > gpurun_out/sampler_under_ncu.log 2>&1
grep -c sample gpurun_out/launches_sampler.csv
