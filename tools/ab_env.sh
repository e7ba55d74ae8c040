#!/bin/bash
# A/B of environment settings on the bench workload, alternating runs.  usage: ROUNDS=3 tools/ab_env.sh "VPU_LN_FOLD=0" "VPU_LN_FOLD=1"
for i in $(seq ${ROUNDS:-3}); do
  for E in "$@"; do
    env $E python bench.py --no-e2e --no-cpu-baseline --no-noc --no-eager --profile-steps 0 --steps 40 2>/dev/null > /tmp/ab_line.json
    python - "$E" <<'PY'
import json, sys
d = json.loads(open('/tmp/ab_line.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step %.3f  sm_mhz %s  power %s' % (d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')), flush=True)
PY
  done
done
