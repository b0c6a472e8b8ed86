#!/bin/bash
# RoPE / split fused into the QKV GEMM's epilogue: GPU suite, A/B, then ncu --set full of one CTA-pair gate/up GEMM and one fused attention launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2w}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=4 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -12 gpurun_out/${TAG}_pytest_gpu.log
if grep -q "failed\|error" gpurun_out/${TAG}_pytest_gpu.log; then
  echo "== same suite with the stand-alone RoPE kernel"
  ( time GL_PREFILL_FUSE_ROPE=0 timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu_nofuse.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu_nofuse.log
  export GL_PREFILL_FUSE_ROPE=0
fi
PROBE_VARIANTS="GL_PREFILL_FUSE_ROPE=1;GL_PREFILL_FUSE_ROPE=0" timeout 400 python tools/prefill_attn_probe.py 512 2048 > gpurun_out/${TAG}_prefill_rope_probe.log 2>&1; grep "^{" gpurun_out/${TAG}_prefill_rope_probe.log || tail -5 gpurun_out/${TAG}_prefill_rope_probe.log
# one gate/up GEMM (third tcgen05 launch of a layer) of the SECOND prompt pass, and one attention launch, 2048 tokens
PROBE_VARIANTS="GL_NONE=1" timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_tc5 --launch-skip 130 --launch-count 1 -o gpurun_out/${TAG}_gateup_pair -f python tools/prefill_attn_probe.py 2048 > gpurun_out/${TAG}_ncu_gemm.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_gemm.log
PROBE_VARIANTS="GL_NONE=1" timeout 400 ncu --set full --import-source on --clock-control none -k regex:flash_prefill --launch-skip 40 --launch-count 1 -o gpurun_out/${TAG}_flash -f python tools/prefill_attn_probe.py 2048 > gpurun_out/${TAG}_ncu_flash.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_flash.log
ls -la gpurun_out/*.ncu-rep
