#!/bin/bash
# GPU box: pipeline-part count with the tensor-memory ADMM build (296 resident CTAs): config 1 (B = 1024), value + e2e
for parts in 1 2 3 4 5 6 7 8; do
  python bench.py --parts $parts --steps 40 --warmup 5 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('parts $parts: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'per-step calls', round(d['per_step_calls']['value']), 'e2e', round(d['e2e']['value']), 'joined', round(d['e2e']['joined']['value']), 'cold', round(d['cold_start']['value']))"
done
