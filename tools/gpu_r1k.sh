#!/bin/bash
# N-GPU knob sweep of the sharded sweep: z-block size x team mode
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    tools/multi_knobs.py 2048 50 "16;32;64;128;64,CLAPCA_TEAM=0;32,CLAPCA_TEAM=0;64,CLAPCA_EDGE_FLAG_ROWS=2;32,CLAPCA_TEAM=8" 2>&1 | grep -E "^N=|Error|error|Traceback" | tee gpurun_out/knobs_multi_team_n$N.txt
