#!/bin/bash
# layout items / streamed run: parity first, then the end-to-end probe, the default bench line, the full GPU suite
set -u
mkdir -p gpurun_out
{
echo "== pytest streamed / fused"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed or fused" 2>&1 | tail -15
echo "== e2e probe 1024"; timeout 300 python tools/e2e_probe.py 1024 50 2>&1 | grep -v "^$"
echo "== e2e probe 2048"; timeout 600 python tools/e2e_probe.py 2048 50 2>&1 | grep -v "^$"
echo "== bench default"; timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -3
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r1e.txt
