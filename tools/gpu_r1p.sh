#!/bin/bash
# final 1-GPU pass of the round: whole GPU suite, smoke, then one bench line per BASELINE config
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in ca3d_128 ca2d_16384 ca2d_16384_cavetest terrain_8192 noise_256 terrain_mesh_8192; do
  echo "== bench $w"; timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_$w.json | cut -c1-700
done
} 2>&1 | tee gpurun_out/r1p.txt
