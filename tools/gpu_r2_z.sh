#!/bin/bash
# tcgen05 prompt attention (opt-in): parity against the mma.sync kernel, then A/B of the prompt pass at 512 / 2048 tokens
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2z}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_decode.py -x -q -m gpu -p no:cacheprovider -k "tcgen05_prompt_attention" ) > gpurun_out/${TAG}_pytest_attn_tc5.log 2>&1; tail -12 gpurun_out/${TAG}_pytest_attn_tc5.log
PROBE_VARIANTS="GL_PREFILL_ATTN_TC5=1;GL_PREFILL_ATTN_TC5=0" timeout 300 python tools/prefill_attn_probe.py 512 2048 > gpurun_out/${TAG}_prefill_attn_tc5_probe.log 2>&1; grep "^{" gpurun_out/${TAG}_prefill_attn_tc5_probe.log; grep -v "^{" gpurun_out/${TAG}_prefill_attn_tc5_probe.log | tail -4
