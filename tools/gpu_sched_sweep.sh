#!/bin/bash
# GPU box: sweep of the task-scheduling cost model (PGN_SCHED="a,b,c": lat = a + b K + c sh) for the tensor-memory ADMM build
for s in "120,22,30" "200,22,30" "300,22,30" "120,35,30" "120,15,30" "120,22,60" "60,22,30" "200,35,45"; do
  PGN_SCHED=$s python bench.py --steps 40 --warmup 5 --other-configs none --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$s', round(d['value']), round(d['stage_ms_per_step']['admm'],4), d['details']['qp']['program']['l_slots'], d['details']['qp']['program']['admm_variant'], {k: round(v,3) for k,v in d['admm_phase_share'].items() if k in ('solve','factor')})"
done
