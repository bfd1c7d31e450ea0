#!/bin/bash
# tuning experiment: kernel time of the fused sweep vs claim order, generation batch, counter batching
W=${1:-ca3d_2048}
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.1f GCUPS total, pop %d' % (d['roofline']['kernel_ms'], d['value'], d['config']['population']))"); echo "$@ : $r"; }
{
run CLAPCA_ORDER=0 CLAPCA_FLAG_ROWS=8
for gb in 4 8 12 16 25 50; do for fr in 4 8; do run CLAPCA_ORDER=2 CLAPCA_GEN_BATCH=$gb CLAPCA_FLAG_ROWS=$fr; done; done
run CLAPCA_ORDER=2 CLAPCA_GEN_BATCH=12 CLAPCA_FLAG_ROWS=2
run CLAPCA_ORDER=2 CLAPCA_GEN_BATCH=12 CLAPCA_FLAG_ROWS=16
} | tee gpurun_out/knobs3_$W.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
for gb in 8 16; do
CLAPCA_GEN_BATCH=$gb timeout 900 ncu --metrics $M --clock-control none -k regex:ca3d_sweep -c 1 --csv --log-file gpurun_out/knobs3_ncu_gb$gb.csv \
      python bench.py --workload $W --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/knobs3_ncu.log 2>&1
grep -E "dram__|gpu__time|hit_rate|issue_active" gpurun_out/knobs3_ncu_gb$gb.csv | awk -F'","' -v g=$gb '{print "gb" g, $(NF-2), $(NF-1), $NF}'
done
