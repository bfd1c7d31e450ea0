#!/bin/bash
set -u
echo "== pytest gpu (2D)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ca2d" 2>&1 | tail -8
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms kernel, %.3f ms step, %.1f GCUPS total, planes %d workers %d pop %d' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value'], d['config']['planes'], d['config']['workers'], d['config']['population']))"); echo "$W $@ : $r"; }
{
W=ca2d_16384
for wpl in 1 2 4; do for fr in 2 4 8; do run CLAPCA_2D_WPL=$wpl CLAPCA_2D_FLAG_ROWS=$fr; done; done
W=ca2d_16384_cavetest
for wpl in 1 2 4; do run CLAPCA_2D_WPL=$wpl; done
W=ca2d_4096
for wpl in 1 2 4; do run CLAPCA_2D_WPL=$wpl; done
} | tee gpurun_out/knobs_2d.txt
