#!/usr/bin/env python
"""GPU box: per-step stage times and ADMM iteration statistics of the bench workload's closed loop (transient after the cold start)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pigeon.jl_b200 as p
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
trajs, tid, state, control, t0, other = bench.make_workload(B)
g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
g.set_state(state, control, other)
g.set_profiling(1)
for k in range(steps):
    g.stage_ms(reset=True)
    g.step(t0 + 0.01 * k); g.rollout(0.01)
    st = g.stage_ms(reset=True); s = g.stats()
    it = s["iters"]
    print("step %3d admm %7.3f ms nodes %.3f lin %.3f roll %.3f | iters mean %6.1f p99 %5d max %5d  rho_upd mean %.2f max %d  not solved %d" % (
        k, st["admm"], st["nodes"], st["linearize"], st["rollout"], it.mean(), np.percentile(it, 99), it.max(), s["rho_updates"].mean(), s["rho_updates"].max(), int((s["status"] != 1).sum())))
