#!/bin/bash
set -u
N=${1:-2}
export CLAPCA_DIAG=1
echo "== N=1"; timeout 300 python bench.py --workload ca3d_2048 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -E "diag|^\{" | cut -c1-200 | tail -3
echo "== N=1 z1024"; timeout 300 python bench.py --workload ca3d_2048_z1024 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -E "diag|^\{" | cut -c1-200 | tail -3
for b in 16 64; do
echo "== N=$N B=$b"; CLAPCA_BLOCK_PLANES=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --workload ca3d_2048 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -E "diag|^\{" | cut -c1-220 | tail -$((2*N+1))
done
