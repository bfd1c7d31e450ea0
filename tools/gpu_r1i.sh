#!/bin/bash
# 1-GPU pass: instantiator tests, then the ncu launch list (time + DRAM bytes) of the library's kernels in the default bench
set -u
mkdir -p gpurun_out
{
echo "== pytest instantiators / seeding"; timeout 600 python -m pytest tests -m gpu -q -k "instantiators or seeding" 2>&1 | tail -4
echo "== ncu launch list of the default bench (1 warm-up + 1 step), library kernels only"
timeout 600 ncu -k regex:clapca --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv \
    --log-file gpurun_out/launches_bench_ca3d_2048.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
grep -c . gpurun_out/launches_bench_ca3d_2048.csv
} 2>&1 | tee gpurun_out/r1i.txt
