#!/bin/bash
# compute-sanitizer passes over the kernel tests and one forward per architecture.  GPU box:
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh'      -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, timeout, command...
    local name=$1 tool=$2 to=$3; shift 3
    timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 77 --log-file gpurun_out/sanitizer_${name}.log "$@" > gpurun_out/sanitizer_${name}.out 2>&1
    echo "$name rc=$? $(grep -c '=========' gpurun_out/sanitizer_${name}.log) log lines; $(tail -n 1 gpurun_out/sanitizer_${name}.log)"
    tail -n 3 gpurun_out/sanitizer_${name}.out
}
run memcheck_kernels memcheck 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu
run memcheck_forward memcheck 900 python tools/sanitize_forward.py
run racecheck_forward racecheck 900 python tools/sanitize_forward.py vit_base
run synccheck_forward synccheck 600 python tools/sanitize_forward.py vit_base
run racecheck_kernels racecheck 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention or gemm_epilogues or pixel_shuffle"
