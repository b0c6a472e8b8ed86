#!/bin/bash
# decode attention with the splits of a KV head as one cluster (DSMEM merge): parity tests, then batch-1 decode A/B against the ticket path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2x}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -m gpu -p no:cacheprovider -k "cluster_attention or long_context or decode_steps" ) > gpurun_out/${TAG}_pytest_cluster.log 2>&1; tail -8 gpurun_out/${TAG}_pytest_cluster.log
DECODE_VARIANTS="GL_NONE=1;GL_ATTN_CLUSTER=1,GL_ATTN_SPLITS=16;GL_ATTN_CLUSTER=1,GL_ATTN_SPLITS=8;GL_ATTN_SPLITS=16;GL_NONE=2" timeout 500 python tools/decode_ab.py > gpurun_out/${TAG}_decode_ab.log 2>&1; grep "^{" gpurun_out/${TAG}_decode_ab.log || tail -8 gpurun_out/${TAG}_decode_ab.log
