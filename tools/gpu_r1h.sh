#!/bin/bash
# 1-GPU pass: whole GPU suite, then the ncu launch list (time + DRAM bytes) of the default bench command
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== ncu launch list of the default bench (1 warm-up + 1 step)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv \
    --log-file gpurun_out/launches_bench_ca3d_2048.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
grep -c . gpurun_out/launches_bench_ca3d_2048.csv
} 2>&1 | tee gpurun_out/r1h.txt
