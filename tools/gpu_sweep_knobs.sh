#!/bin/bash
# tuning experiment: kernel time of the fused sweep vs claim order, counter batching, fence
W=${1:-ca3d_2048}
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.1f GCUPS total, pop %d' % (d['roofline']['kernel_ms'], d['value'], d['config']['population']))"); echo "$@ : $r"; }
{
for fr in 1 2 4 8; do run CLAPCA_ORDER=0 CLAPCA_FLAG_ROWS=$fr; done
run CLAPCA_ORDER=0 CLAPCA_FLAG_ROWS=1 CLAPCA_NOFENCE=1
run CLAPCA_ORDER=0 CLAPCA_FLAG_ROWS=4 CLAPCA_NOFENCE=1
for fr in 2 4; do for seg in 512 2048; do run CLAPCA_ORDER=1 CLAPCA_FLAG_ROWS=$fr CLAPCA_SEG_ROWS=$seg; done; done
run CLAPCA_ORDER=1 CLAPCA_FLAG_ROWS=4 CLAPCA_SEG_ROWS=2048 CLAPCA_NOFENCE=1
} | tee gpurun_out/knobs2_$W.txt
