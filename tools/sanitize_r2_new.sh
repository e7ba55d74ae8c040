#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (attention_dma.cu, gemm_b2b.cu, head_tail.cu) and the forwards that use them
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local name=$1 tool=$2 to=$3; shift 3
    timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 77 --log-file gpurun_out/sanitizer2_${name}.log "$@" > gpurun_out/sanitizer2_${name}.out 2>&1
    echo "$name rc=$? $(tail -n 1 gpurun_out/sanitizer2_${name}.log)"; tail -n 2 gpurun_out/sanitizer2_${name}.out; }
run memcheck_new_kernels memcheck 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "b2b or head_tail or dma"
run memcheck_forward memcheck 900 python tools/sanitize_forward.py
run synccheck_forward synccheck 600 python tools/sanitize_forward.py vit_base
run initcheck_forward initcheck 600 python tools/sanitize_forward.py vit_base
