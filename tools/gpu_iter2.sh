#!/bin/bash
set -u
W=ca3d_2048
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.1f GCUPS total, workers %d pop %d' % (d['roofline']['kernel_ms'], d['value'], d['config']['workers'], d['config']['population']))"); echo "$@ : $r"; }
{
for t in 64 128 256; do for gb in 16 25; do run CLAPCA_CTA_THREADS=$t CLAPCA_GEN_BATCH=$gb; done; done
run CLAPCA_CTA_THREADS=128 CLAPCA_ORDER=0
W=ca3d_1024
for t in 64 128 256; do run CLAPCA_CTA_THREADS=$t CLAPCA_GEN_BATCH=16;  done
run CLAPCA_CTA_THREADS=128 CLAPCA_ORDER=0
} | tee gpurun_out/knobs4.txt
W=ca3d_2048
CLAPCA_GEN_BATCH=16 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:ca3d_sweep -c 1 -f -o gpurun_out/prof_sweep_2048_gb16 \
    python bench.py --workload $W --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full_2048.log 2>&1
tail -2 gpurun_out/ncu_full_2048.log | cut -c1-300
