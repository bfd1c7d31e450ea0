#!/bin/bash
# 2-GPU pass: the whole GPU suite (streamed / fused / mesh / multi-GPU parity), mesh bench + launch list, scaling bench N=1,2
set -u
N=${1:-2}
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench terrain_mesh_8192"; timeout 300 python bench.py --workload terrain_mesh_8192 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_terrain_mesh_8192.json | cut -c1-1800
echo "== ncu launch list of the mesh bench"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_terrain_mesh_1024.csv python bench.py --workload terrain_mesh_8192 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_mesh.log 2>&1
grep -c . gpurun_out/launches_terrain_mesh_1024.csv
echo "== bench N=1"; timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_n1.json | cut -c1-300
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus $n --steps 3 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/scale_n$n.json | cut -c1-1500
  fi
done
} 2>&1 | tee gpurun_out/r1f.txt
