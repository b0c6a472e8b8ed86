#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "prefill or embed or generate" > gpurun_out/pytest_prefill.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_prefill.log
tail -n 3 gpurun_out/pytest_prefill.log
python - <<'PY' > gpurun_out/build_model.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle import gguf_synth as S
S.build_model('/dev/shm/prof_llama3_8b.gguf', S.LLAMA3_8B, 'q4_k_m', seed=1234, mode='random', with_vocab=False)
PY
timeout 300 python tools/prefill_probe.py > gpurun_out/prefill_probe.log 2>&1
cat gpurun_out/prefill_probe.log
timeout 400 python tools/bench_quick.py > gpurun_out/bench_quick.log 2>&1
grep -h '^{' gpurun_out/bench_quick.log | sort -u | cut -c1-200
