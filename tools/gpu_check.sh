#!/bin/bash
# First GPU pass: smoke, GPU parity tests, a small and the full bench, and an ncu launch list.
# Run under gpurun from the repo root; everything is written to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench 512"; timeout 300 python bench.py --workload ca3d_512 --steps 3 --warmup 2 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_512.json
echo "== bench 1024"; timeout 300 python bench.py --workload ca3d_1024 --steps 3 --warmup 2 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_1024.json
echo "== bench 2048"; timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 8 2>&1 | tail -3 | tee gpurun_out/bench_2048.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_512.csv \
    python bench.py --workload ca3d_512 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches_512.csv | cut -c1-300
