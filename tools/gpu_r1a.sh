#!/bin/bash
# round-1 re-entry check: parity, default bench line, low-parallelism knob sweep, launch list + full ncu capture
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default bench"; timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 | tee gpurun_out/bench_default.json
echo "== low-parallelism sweep"
timeout 600 python tools/tune_lowpar.py 2048 "50 25 12 6" "0 3 2" "8 2" 2>&1 | grep -v Warning | tee gpurun_out/tune_lowpar.txt
echo "== launch list (default workload)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ca3d_2048.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
grep -c sweep gpurun_out/launches_ca3d_2048.csv
echo "== ncu full (sweep kernel, default order)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ca3d_sweep -c 1 -f -o gpurun_out/prof_sweep_ca3d_2048 \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_sweep_ca3d_2048.ncu-rep --page raw --csv > gpurun_out/prof_sweep_ca3d_2048_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_sweep_ca3d_2048_raw.csv | tee gpurun_out/ncu_full_sweep_ca3d_2048.txt
ls -la gpurun_out/
