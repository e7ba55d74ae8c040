#!/bin/bash
# Round-2 GPU captures in one gpurun call (one B200):  gpurun --timeout 1500 -- 'bash tools/profile_r2.sh'
#   gpurun_out/launches_r2.csv       ncu launch list of the bench command (gpu__time_duration per launch)
#   gpurun_out/gemm_traffic_r2.csv   DRAM bytes of every GEMM-class launch
#   gpurun_out/full_r2/*.csv         ncu --set full (raw + source pages) of the kernels new in round 2
#   gpurun_out/bench_r2.json         the bench line of the same build (not under a profiler)
mkdir -p gpurun_out/full_r2
BARGS="--no-e2e --no-cpu-baseline --no-noc --no-eager --profile-steps 0"
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm --csv \
    --log-file gpurun_out/gemm_traffic_r2.csv python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
BENCH_ARGS="--no-noc --no-eager" bash tools/ncu_full.sh gpurun_out/full_r2 "dma_attention_tc_kernel:10" "dma_attention_tc_kernel:11" \
    "dma_attention_tc_kernel:12" "gemm_b2b_kernel:4" "head_tail_kernel:1" "window_attention_tc_kernel:10" "gemm_tc2_kernel:120"
python tools/ncu_raw_pick.py gpurun_out/full_r2/*[0-9].csv > gpurun_out/full_r2/summary.txt 2>&1
tail -5 gpurun_out/bench_r2.err
