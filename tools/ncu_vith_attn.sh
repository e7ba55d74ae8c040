ncu --set full --clock-control none -k regex:attention_kernel -s 3 -c 1 -f -o /tmp/vith_attn python bench.py --arch vit_huge --batch 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profile-steps 0 > /dev/null 2>&1
ncu -i /tmp/vith_attn.ncu-rep --page raw --csv > gpurun_out/vith_attn.csv 2>/dev/null
wc -l gpurun_out/vith_attn.csv
