for cap in 200 400 -1; do
  PGN_SOLVE_CAP=$cap python bench.py --steps 20 --warmup 3 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cap $cap: value', round(d['value']), 'cold', round(d['cold_start']['value']), '8d', round(d['survey_8d_timing']['value']), d['by_rank'])"
done
