#!/bin/bash
# N-GPU diagnostics of the sharded sweep: where the persistent warps wait (CLAPCA_DIAG), per z-block size
set -u
N=${1:-2}
mkdir -p gpurun_out
CLAPCA_DIAG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    tools/multi_knobs.py 2048 50 "16;128;64,CLAPCA_TEAM=0" 2>&1 | grep -E "^N=|clapca diag|Error|error|Traceback" | tee gpurun_out/diag_multi_team_n$N.txt
