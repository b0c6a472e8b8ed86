#!/bin/bash
# batch-1 decode A/B: attention splits as one cluster (DSMEM merge) against the ticket path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2x2}
mkdir -p gpurun_out
DECODE_VARIANTS="GL_NONE=1;GL_ATTN_CLUSTER=1,GL_ATTN_SPLITS=16;GL_ATTN_CLUSTER=1,GL_ATTN_SPLITS=8;GL_ATTN_SPLITS=16;GL_NONE=2" timeout 500 python tools/decode_ab.py > gpurun_out/${TAG}_decode_ab.log 2>&1; grep "^{" gpurun_out/${TAG}_decode_ab.log; grep -v "^{" gpurun_out/${TAG}_decode_ab.log | tail -5
