#!/bin/bash
# GPU box: deferred solves (PGN_SOLVE_CAP = ADMM iterations of one QP per launch inside the simulate loops, 0 = off) on configs 2, 3 and the cold start of config 1
for cfg in 2 3 1; do
  for cap in 0 200 500 1000; do
    PGN_SOLVE_CAP=$cap python bench.py --config $cfg --steps 30 --warmup 3 --no-cpu --no-latency --other-configs none 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('config $cfg cap $cap: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'cold', round(d['cold_start']['value']), '8d', round(d['survey_8d_timing']['value']), d['admm'])"
  done
done
