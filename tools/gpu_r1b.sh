#!/bin/bash
# publisher-warp mode vs self-publishing workers on one GPU: parity first, then the low-parallelism sweep
set -u
mkdir -p gpurun_out
echo "== parity (default mode)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== parity (publisher mode)"; CLAPCA_PUB_WORKERS=9 CLAPCA_FLAG_ROWS=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
echo "# self-publishing workers (flag_rows 8)"
timeout 300 python tools/tune_lowpar.py 2048 "50 12 6" "0" "8" "CLAPCA_PUB_WORKERS=0"
echo "# publisher mode: N workers per CTA, mailbox every flag_rows rows"
timeout 600 python tools/tune_lowpar.py 2048 "50 12 6" "0" "1 2 4" "CLAPCA_PUB_WORKERS=4;CLAPCA_PUB_WORKERS=9;CLAPCA_PUB_WORKERS=19"
echo "# publisher mode, capped CTAs per SM"
timeout 300 python tools/tune_lowpar.py 2048 "12 6" "1" "1 2" "CLAPCA_PUB_WORKERS=9;CLAPCA_PUB_WORKERS=12"
timeout 300 python tools/tune_lowpar.py 2048 "12 6" "2" "1 2" "CLAPCA_PUB_WORKERS=4;CLAPCA_PUB_WORKERS=6"
} 2>&1 | grep -v Warning | tee gpurun_out/tune_pub.txt
