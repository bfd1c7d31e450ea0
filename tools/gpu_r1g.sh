#!/bin/bash
# 1-GPU pass: whole GPU suite, mesh bench
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench terrain_mesh_8192"; timeout 300 python bench.py --workload terrain_mesh_8192 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_terrain_mesh_8192.json | cut -c1-1300
} 2>&1 | tee gpurun_out/r1g.txt
