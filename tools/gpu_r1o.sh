#!/bin/bash
# 1-GPU verification: whole GPU suite, smoke, default bench line, ncu launch list of the library's kernels
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_ca3d_2048.json | cut -c1-2500
echo "== ncu launch list (library kernels, 1 warm-up + 1 step)"
timeout 400 ncu -k regex:"ca3d_|max_u8" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv \
    --log-file gpurun_out/launches_bench_ca3d_2048.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
grep -c . gpurun_out/launches_bench_ca3d_2048.csv
} 2>&1 | tee gpurun_out/r1o.txt
