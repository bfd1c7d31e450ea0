#!/bin/bash
# ncu captures of the dominant kernel + a launch list of the headline bench command.
set -u
mkdir -p gpurun_out
W=${1:-ca3d_1024}
echo "== quick parity"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench $W"; timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_$W.json
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launches_$W.log 2>&1
grep -c sweep gpurun_out/launches_$W.csv
echo "== ncu full (sweep kernel)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:ca3d_sweep -c 1 -f -o gpurun_out/prof_sweep_$W \
    python bench.py --workload $W --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full_$W.log 2>&1
tail -3 gpurun_out/ncu_full_$W.log
ls -la gpurun_out/*.ncu-rep
