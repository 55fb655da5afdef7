#!/bin/bash
# GPU box: threads per QP of the tensor-memory ADMM build (two CTAs per SM in every case)
for nt in 256 384 512; do
  PGN_TMEM_THREADS=$nt python bench.py --steps 40 --warmup 5 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); p=d['details']['qp']['program']; print('threads $nt:', p['admm_threads'], p['admm_variant'], p['admm_smem_bytes'], 'value', round(d['value']), 'admm ms', round(d['stage_ms_per_step']['admm'],4), 'e2e', round(d['e2e']['value']), 'joined', round(d['e2e']['joined']['value']), 'cold', round(d['cold_start']['value']), {k: round(v,3) for k,v in d['admm_phase_share'].items()})"
done
