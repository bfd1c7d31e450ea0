#!/bin/bash
# N-GPU pass: sharded parity, then the scaling bench at N
set -u
N=${1:-2}
mkdir -p gpurun_out
{
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/scale_n$N.json | cut -c1-1600
} 2>&1 | tee gpurun_out/r1j_n$N.txt
