#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (gemm_gn.cu incl. the pixel-shuffle TMA store, gemm_res.cu MODE_TAB,
# gemm_ln.cu) and the forwards that use them
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local name=$1 tool=$2 to=$3; shift 3
    timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 77 --log-file gpurun_out/sanitizer3_${name}.log "$@" > gpurun_out/sanitizer3_${name}.out 2>&1
    echo "$name rc=$? $(tail -n 1 gpurun_out/sanitizer3_${name}.log)"; tail -n 2 gpurun_out/sanitizer3_${name}.out; }
run memcheck_new_kernels memcheck 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "table or pixel_shuffle or gemm_layernorm or bf16_residual"
run memcheck_forward memcheck 900 python tools/sanitize_forward.py
run synccheck_forward synccheck 600 python tools/sanitize_forward.py vit_base
run initcheck_forward initcheck 600 python tools/sanitize_forward.py vit_base
