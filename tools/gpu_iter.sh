#!/bin/bash
# quick iteration pass: GPU parity, benches at 1024^3 / 2048^3, light ncu metrics (time, DRAM bytes, issue) of the sweep kernel
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for W in ca3d_1024 ca3d_2048; do
  echo "== bench $W"; timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/iter_bench_$W.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'GCUPS; kernel ms', d['roofline']['kernel_ms'], 'step ms', d['ms_per_step'], 'pop', d['config']['population'])"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
for W in ca3d_1024 ca3d_2048; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:ca3d_sweep -c 1 --csv --log-file gpurun_out/iter_ncu_$W.csv \
      python bench.py --workload $W --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/iter_ncu_$W.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/iter_ncu_$W.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print('$W', r[h.index('Metric Name')], r[h.index('Metric Unit')], r[h.index('Metric Value')])
PY
done
