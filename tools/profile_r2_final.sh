#!/bin/bash
# Final round-2 captures in one gpurun call (one B200):  gpurun --timeout 1500 -- 'bash tools/profile_r2_final.sh'
#   gpurun_out/bench_r2.json          the bench line of this build (not under a profiler)
#   gpurun_out/launches_r2.csv        ncu launch list of the bench command (gpu__time_duration per launch)
#   gpurun_out/gemm_traffic_r2.csv    DRAM bytes of every launch of the bench's `gemm.*` classes (gemm_ln_kernel is its own class)
mkdir -p gpurun_out
BARGS="--no-e2e --no-cpu-baseline --no-noc --no-eager --profile-steps 0"
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_(tc|tc2|res|gn|b2b)_kernel" --csv \
    --log-file gpurun_out/gemm_traffic_r2.csv python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
tail -3 gpurun_out/bench_r2.err
grep -c gemm gpurun_out/gemm_traffic_r2.csv
