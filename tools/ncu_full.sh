#!/bin/bash
# One `ncu --set full` capture per kernel of interest (second forward of a 2-forward run), exported as raw CSV on the box
# (the .ncu-rep files stay there).  usage: tools/ncu_full.sh out_dir "regex1:skip" "regex2:skip" ...
OUT=$1; shift
mkdir -p $OUT
for spec in "$@"; do
  re=${spec%%:*}; skip=${spec##*:}
  name=$(echo $re | tr -c 'a-zA-Z0-9_' '_')_$skip
  ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o /tmp/full_$name \
      python bench.py $BENCH_ARGS --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profile-steps 0 > /dev/null 2>&1
  ncu -i /tmp/full_$name.ncu-rep --page raw --csv > $OUT/$name.csv 2>/dev/null
  ncu -i /tmp/full_$name.ncu-rep --page source --csv > $OUT/${name}_source.csv 2>/dev/null
  echo "$name: $(wc -l < $OUT/$name.csv) lines"
done
