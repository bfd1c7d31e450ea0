#!/bin/bash
set -u
W=ca3d_2048
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.1f GCUPS total, workers %d pop %d' % (d['roofline']['kernel_ms'], d['value'], d['config']['workers'], d['config']['population']))"); echo "$W $@ : $r"; }
{
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for pf in 0 4 8 16 32; do run CLAPCA_PREFETCH_ROWS=$pf; done
run CLAPCA_PREFETCH_ROWS=8 CLAPCA_CTA_THREADS=64
run CLAPCA_PREFETCH_ROWS=8 CLAPCA_FLAG_ROWS=4
run CLAPCA_PREFETCH_ROWS=8 CLAPCA_FLAG_ROWS=16
W=ca3d_1024
for pf in 0 8; do run CLAPCA_PREFETCH_ROWS=$pf; done
W=ca3d_512
run CLAPCA_PREFETCH_ROWS=0
} 2>&1 | tee gpurun_out/knobs5.txt
