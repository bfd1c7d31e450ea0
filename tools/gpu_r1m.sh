#!/bin/bash
# N-GPU: sharded parity, then z-block sweep with wait diagnostics
set -u
N=${1:-2}
mkdir -p gpurun_out
{
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
CLAPCA_DIAG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    tools/multi_knobs.py 2048 50 "16;32;64;128" 2>&1 | grep -E "^N=|clapca diag \[slab rank 0|Error|error|Traceback" | uniq
} 2>&1 | tee gpurun_out/r1m_n$N.txt
