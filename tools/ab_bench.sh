#!/bin/bash
# A/B of several builds of the library on the bench workload, alternating runs (the boards are power-managed: clocks
# drift by ~10 % between runs, so single runs do not compare).  usage: ROUNDS=3 tools/ab_bench.sh libA.so libB.so ...
for i in $(seq ${ROUNDS:-3}); do
  for L in "$@"; do
    VPU_LIB_PATH=$L python bench.py --no-e2e --no-cpu-baseline --no-noc --no-eager --profile-steps 0 --steps 40 2>/dev/null > /tmp/ab_line.json
    python - "$L" <<'PY'
import json, sys
d = json.loads(open('/tmp/ab_line.json').read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], 'ms/step %.3f  sm_mhz %s  power %s' % (d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')), flush=True)
PY
  done
done
