#!/bin/bash
# multi-GPU pass: parity (sharded == single GPU == oracle), then the scaling bench at N = 1 and N = all
set -u
N=${1:-2}
W=${2:-ca3d_2048}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
echo "== bench N=1"; timeout 600 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | tee gpurun_out/scale_${W}_n1.json | cut -c1-400
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus $n --workload $W --steps 3 --warmup 2 --no-cpu 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/scale_${W}_n$n.json | cut -c1-600
  fi
done
