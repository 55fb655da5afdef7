#!/usr/bin/env python
"""GPU box: small pass over every entry point (step, rollout, simulate, callback graph, HJI policy; coupled N=31 / N=16, decoupled) meant to be
run under `compute-sanitizer --tool memcheck` (recorded in profiles/r1b_sanitizer.md)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pigeon.jl_b200 as p
trajs = p.synthetic.synthetic_trajectories(n_traj=2, n_nodes=300)
B = 6
tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
other = np.tile(np.array([1e4, 1e4, 0.0, 5.0]), (B, 1))
for cls, kw in ((p.BatchedCoupledTrajectoryTrackingMPC, {}), (p.BatchedCoupledTrajectoryTrackingMPC, dict(N_short=5, N_long=10)), (p.BatchedDecoupledTrajectoryTrackingMPC, {})):
    m = cls(p.X1(), trajs, B, trajectory_index=tid, **kw)
    m.set_state(state, control, other)
    m.set_guards(nan_fallback=True, pause_below_speed=1.0)
    u = m.step(t0)
    m.rollout(0.01)
    m.simulate_device(t0 + 0.01, 0.01, 2)
    q, uu = m.get_state()
    out = m.from_autobox(q, uu, 0.0, other_car=other)
    out = m.from_autobox(q, out[:, :3], 0.0)
    knots, V, gV = p.synthetic.analytic_hji_grid((5, 4, 5, 4, 3, 4, 3))
    m.set_HJI_cache(p.HJICache(knots, V, gV)); m.set_hji_policy(True)
    oth = other.copy(); oth[:, 0] = q[:, 0] + 1.0; oth[:, 1] = q[:, 1]; oth[:, 2] = q[:, 2]
    out = m.from_autobox(q, out[:, :3], 0.0, other_car=oth)
    print(cls.__name__, kw, np.isfinite(out).all(), m.stats()["iters"])
    # pipeline parts: vehicle ranges on their own streams (ragged: 4 parts over 6 vehicles), joined per call and free-running in simulate
    m.set_pipeline_parts(4)
    m.set_state(state, control, other)
    u = m.step(t0)
    m.rollout(0.01)
    m.simulate_device(t0 + 0.01, 0.01, 3)
    import torch
    d = torch.tensor(t0 + 0.04, dtype=torch.float64, device="cuda"); o3 = torch.zeros(3 * B, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    m.step_rollout_device(d.data_ptr(), o3.data_ptr(), 0.01)
    m.simulate_device_async(d.data_ptr(), 0.01, 2, k0=1)
    m.synchronize()
    print("  parts", m.pipeline_parts, np.isfinite(m.get_state()[0]).all(), m.stats()["iters"])
    m.close()
print("done")
