#!/bin/bash
# tuning: sharded sweep time vs z-block size, claim order and counter batching
set -u
N=${1:-2}; W=${2:-ca3d_2048}; BS=${3:-"16 32 64"}
run() { r=$(env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --workload $W --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -E '^\{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms kernel, %.2f ms step, %.1f GCUPS, pop %d' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value'], d['config']['population']))"); echo "N=$N $@ : $r"; }
{
for b in $BS; do run CLAPCA_BLOCK_PLANES=$b; done
} | tee -a gpurun_out/knobs_multi_${W}_n$N.txt
