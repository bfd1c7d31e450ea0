#!/bin/bash
# tools/gpu.sh -- everything that is run on the GPU box, as ONE parameterised script:
#     gpurun [--gpus N] --timeout S -- 'bash tools/gpu.sh <task> [args...] ; bash tools/gpu.sh <task> ...'
# Tasks write under gpurun_out/ (merged back by gpurun); summaries worth keeping are copied to profiles/ by hand.
#   check [pytest-args]          the -m gpu suite (+ smoke)
#   bench [bench.py args]        one bench line -> gpurun_out/bench_<tag>.json   (TAG=... names the file)
#   golden                       default bench line + tests/golden/plane_hashes_ca3d_2048.json (copied to gpurun_out/)
#   knobs "<side>" "<gens>" "<env;env;...>"     single-GPU knob sweep of the ca3d sweep (tools/tune_lowpar.py)
#   launches [bench.py args]     ncu launch list (time + DRAM bytes per launch) of a bench command
#   ncufull <kernel-regex> [bench.py args]      ncu --set full of one kernel -> .ncu-rep + text summary
#   ab2d [side [gens [steps]]]   row engine vs diagonal engine on BASELINE config 3, one process (tools/ab_ca2d.py)
#   multi N [bench.py args]      torchrun bench on N GPUs -> gpurun_out/bench_<tag>_n<N>.json
#   multiknobs N "<env;env;...>" [steps]        N-GPU knob sweep inside one set of processes (tools/multi_knobs.py)
set -u
mkdir -p gpurun_out
task=${1:-check}; shift || true
TAG=${TAG:-default}
case "$task" in
check)
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
    ;;
bench)
    timeout 1200 python bench.py "$@" 2> gpurun_out/bench_$TAG.err | tail -1 | tee gpurun_out/bench_$TAG.json
    tail -5 gpurun_out/bench_$TAG.err
    ;;
golden)
    timeout 1500 python bench.py --write-plane-hashes "$@" 2> gpurun_out/bench_$TAG.err | tail -1 | tee gpurun_out/bench_$TAG.json
    tail -5 gpurun_out/bench_$TAG.err
    cp tests/golden/plane_hashes_ca3d_2048.json gpurun_out/ 2>/dev/null
    ;;
ab2d)
    timeout 200 python tools/ab_ca2d.py "$@" 2>&1 | tee gpurun_out/ab_ca2d_$TAG.txt
    ;;
knobs)
    timeout 1500 python tools/tune_lowpar.py "${1:-2048}" "${2:-50}" "0" "8" "${3:-}" 2>&1 | tee -a gpurun_out/knobs_$TAG.txt
    ;;
launches)
    # only the library's own kernels (namespace clapca): the seed generator alone launches thousands of torch kernels
    timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \
        --kernel-name-base demangled -k regex:clapca \
        --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary "$@" \
        > gpurun_out/launches_$TAG.log 2>&1
    tail -2 gpurun_out/launches_$TAG.log; wc -l gpurun_out/launches_$TAG.csv
    ;;
ncufull)
    k=$1; shift
    timeout 1800 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -c 1 -f -o gpurun_out/prof_$TAG \
        python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-secondary "$@" > gpurun_out/ncufull_$TAG.log 2>&1
    tail -2 gpurun_out/ncufull_$TAG.log
    ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/prof_$TAG.details.txt 2>&1
    ls -la gpurun_out/prof_$TAG.ncu-rep
    ;;
multi)
    n=$1; shift
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29617 \
        bench.py --gpus $n "$@" 2> gpurun_out/bench_${TAG}_n$n.err | tail -1 | tee gpurun_out/bench_${TAG}_n$n.json
    grep -E "clapca diag|Error|error" gpurun_out/bench_${TAG}_n$n.err | tail -20
    ;;
multiknobs)
    n=$1; shift
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29618 \
        tools/multi_knobs.py "$@" 2>&1 | grep -v "^W\|^\[W" | tee -a gpurun_out/multiknobs_${TAG}_n$n.txt
    ;;
*)
    echo "unknown task $task"; exit 2;;
esac
