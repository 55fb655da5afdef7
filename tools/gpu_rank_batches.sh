#!/bin/bash
# GPU box: the config-1 batches of ranks 0..7 (seed shift 1000 r), each alone on one GPU: which batch sets the max-over-ranks time at N = 8?
for r in 0 1 2 3 4 5 6 7; do
  python bench.py --seed-shift $((1000*r)) --steps 50 --warmup 5 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('rank $r batch: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['admm'], d['by_rank'])"
done
