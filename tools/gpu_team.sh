#!/bin/bash
# team mode (CLAPCA_TEAM): parity, then the low-parallelism sweep on one GPU (gens 50/12/6 ~ N = 1/4/8)
set -u
mkdir -p gpurun_out
{
echo "== pytest team"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "team" 2>&1 | tail -3
echo "== lowpar"
timeout 900 python tools/tune_lowpar.py 2048 "50 12 6" "0" "8" "CLAPCA_TEAM=0 CLAPCA_EDGE_FLAG_ROWS=0;CLAPCA_TEAM=16 CLAPCA_EDGE_FLAG_ROWS=0;CLAPCA_TEAM=16 CLAPCA_EDGE_FLAG_ROWS=2;CLAPCA_TEAM=16 CLAPCA_EDGE_FLAG_ROWS=1;CLAPCA_TEAM=12 CLAPCA_EDGE_FLAG_ROWS=2;CLAPCA_TEAM=8 CLAPCA_EDGE_FLAG_ROWS=2" 2>&1 | grep -v "^$"
echo "== lowpar flag_rows 4/16 with teams"
timeout 600 python tools/tune_lowpar.py 2048 "50 6" "0" "4 16" "CLAPCA_TEAM=16 CLAPCA_EDGE_FLAG_ROWS=2" 2>&1 | grep -v "^$"
echo "== diag"
CLAPCA_DIAG=1 timeout 600 python tools/tune_lowpar.py 2048 "50 6" "0" "8" "CLAPCA_TEAM=16 CLAPCA_EDGE_FLAG_ROWS=2" 2>&1 | grep -v "^$" | tail -8
} 2>&1 | tee gpurun_out/team_lowpar.txt
