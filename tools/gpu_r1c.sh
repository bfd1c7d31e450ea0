#!/bin/bash
# Re-entry pass on one B200: GPU parity tests, the default bench line, the field / 2D workloads, the ncu launch
# list of the default bench command and one full capture of the team sweep kernel (roofline.traffic).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu.txt
echo "== bench default"; timeout 600 python bench.py --cpu-seconds 10 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-700
for w in terrain_8192 noise_256 ca2d_16384; do
  echo "== bench $w"; timeout 300 python bench.py --workload $w --cpu-seconds 8 2>&1 | tail -1 | tee gpurun_out/bench_$w.json | cut -c1-900
done
echo "== bench terrain_8192 direct"; CLAPCA_TERRAIN_DIRECT=1 timeout 300 python bench.py --workload terrain_8192 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_terrain_8192_direct.json | cut -c1-500
echo "== launch list (default command, fewer steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launches_default.log 2>&1
grep -c "ca3d_" gpurun_out/launches_default.csv
echo "== ncu full (team sweep kernel, 2048^3 x 50)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ca3d_team -c 1 -f -o gpurun_out/prof_team_2048 \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full_team.log 2>&1
tail -2 gpurun_out/ncu_full_team.log
echo "== ncu full (terrain kernels)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:terrain_ -c 3 -f -o gpurun_out/prof_terrain_8192 \
    python bench.py --workload terrain_8192 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full_terrain.log 2>&1
tail -2 gpurun_out/ncu_full_terrain.log
echo "== ncu full (noise kernel)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:noise_bake -c 1 -f -o gpurun_out/prof_noise_256 \
    python bench.py --workload noise_256 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full_noise.log 2>&1
tail -2 gpurun_out/ncu_full_noise.log
ls -la gpurun_out/*.ncu-rep
