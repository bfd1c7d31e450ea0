#!/bin/bash
set -u
run() { r=$(env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.2f ms step, %.1f GCUPS total, workers %d pop %d' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value'], d['config']['workers'], d['config']['population']))"); echo "$W $@ : $r"; }
for W in ca3d_2048 ca3d_2048_z1024 ca3d_2048_z256; do run CLAPCA_ORDER=0; done
