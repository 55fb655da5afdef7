#!/usr/bin/env python
"""GPU box: '1 cold + 100 warm steps' of config 1 as one simulate call, with / without round graphs and deferred solves: time and catch-up rounds."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import pigeon.jl_b200 as p
trajs, tid, state, control, t0, other = bench.make_workload(1, 1024, 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
m.set_stream(stream.cuda_stream)
d = torch.tensor(t0, dtype=torch.float64, device="cuda")
for cap in (-1, 0, 200, 1000):
    for rep in range(2):
        m.set_solve_cap(cap)
        m.reset_solver(); m.reset_solved(); m.set_state(state, control, other)
        m.stage_ms(reset=True)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t = time.perf_counter()
        e0.record(stream); m.simulate_device_async(d.data_ptr(), 0.01, 101, k0=0); e1.record(stream); m.synchronize(); e2.record(stream); torch.cuda.synchronize()
        print(f"cap {cap} rep {rep} graphs {'off' if os.environ.get('PGN_NO_GRAPHS') else 'on'}: rounds enqueued {e0.elapsed_time(e1):.1f} ms, incl. catch-up {e0.elapsed_time(e2):.1f} ms, wall {1e3*(time.perf_counter()-t):.1f} ms, catch-up rounds {m.stage_ms(reset=True)['catchup_rounds']}, steps/s {1024*101/e0.elapsed_time(e2)*1e3:.0f}")
m.close()
