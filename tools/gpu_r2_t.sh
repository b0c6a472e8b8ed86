#!/bin/bash
# A/B of the batched step's switches on one box: folded RMSNorm, cluster mode, dependent launch, fused RoPE
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2t}
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 200 python tools/batch_probe.py 32 576 8 2 2>&1 | grep "^{"; env $1 timeout 200 python tools/batch_probe.py 64 576 8 2 2>&1 | grep "^{"; }
{
run "GL_NONE=1"
run "GL_BATCH_FOLD_NORM=0"
run "GL_QGEMM_CLUSTER=0"
run "GL_QGEMM_CLUSTER=0 GL_BATCH_FOLD_NORM=0"
run "GL_BATCH_PDL=0"
run "GL_BATCH_FUSE_ROPE=0"
run "GL_NONE=2"
} > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_ab.log
