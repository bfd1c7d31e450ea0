#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
{
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    tools/multi_knobs.py 2048 50 "16;32;64;128;16,CLAPCA_GHOST_EAGER=1;16,CLAPCA_GHOST_SCATTER=1" 2>&1 | grep -E "^N=|Error|error|Traceback"
} 2>&1 | tee gpurun_out/r1n_n$N.txt
