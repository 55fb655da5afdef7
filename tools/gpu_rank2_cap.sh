for cap in 0 200 400 1000; do
  PGN_SOLVE_CAP=$cap python bench.py --seed-shift 2000 --steps 50 --warmup 5 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('rank-2 batch cap $cap: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'cold', round(d['cold_start']['value']), '8d', round(d['survey_8d_timing']['value']), d['by_rank'])"
done
