#!/usr/bin/env python
"""GPU box: how many vehicles of the config-1 workload (no guards, as `simulate` has none) go NaN in closed loop, per seed shift (= rank), and
what their QPs cost (status / iterations)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pigeon.jl_b200 as p
for shift in (0, 1000, 2000, 3000):
    trajs, tid, state, control, t0, other = bench.make_workload(1, 1024, shift)
    m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
    m.set_state(state, control, other)
    out = []
    for n in (1, 2, 3, 10, 30, 85):
        m.simulate_device(t0, 0.01, n) if n == 1 else None
    m.reset_solver(); m.reset_solved(); m.set_state(state, control, other)
    done = 0
    for n in (1, 2, 3, 10, 30, 85):
        m._lib.pgn_simulate_device  # noqa
        import torch
        d = torch.tensor(t0, dtype=torch.float64, device="cuda")
        m.simulate_device_async(d.data_ptr(), 0.01, n - done, k0=done); m.synchronize(); done = n
        q, u = m.get_state(); st = m.stats()
        bad = ~np.isfinite(q).all(axis=1)
        out.append((n, int(bad.sum()), int((~np.isfinite(u).all(axis=1)).sum()), {int(s): int((st["status"] == s).sum()) for s in np.unique(st["status"])}, int(st["iters"].max()),
                    int(st["iters"][bad].max()) if bad.any() else None))
    print("seed shift", shift, out)
    m.close()
