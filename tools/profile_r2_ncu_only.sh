BARGS="--no-e2e --no-cpu-baseline --no-noc --no-eager --profile-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_(tc|tc2|res|gn|b2b)_kernel" -c 400 --csv \
    --log-file gpurun_out/gemm_traffic_r2.csv python bench.py --steps 1 --warmup 3 $BARGS > /dev/null 2>&1
grep -c gemm gpurun_out/gemm_traffic_r2.csv
