#!/bin/bash
# Local helper (runs in the build container, not on the GPU box): gpurun answers 3 when no box / slot is free right
# now (nothing is charged) -- retry until the call goes through.   usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 45
done
exit 3
