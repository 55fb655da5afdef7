#!/usr/bin/env python
"""GPU box: cycle shares inside the Ruiz equilibration of the ADMM kernel (profiling build, counters of thread 0 summed over CTAs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pigeon.jl_b200 as p
trajs, tid, state, control, t0, other = bench.make_workload(1, 1024, 0)
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
m.set_state(state, control, other)
for k in range(35):
    m.step(t0 + 0.01 * k); m.rollout(0.01)
m.set_profiling(2); m.admm_cycles(reset=True)
m.step(t0 + 0.35)
out = np.zeros(512); m._lib.pgn_get_admm_cycles(m._h, out.ctypes.data_as(__import__("ctypes").c_void_p), 1)
tot = out[:8].sum()
print("phases (gather, ruiz, factor, solve, update, check, store, ticket):", np.round(out[:8] / tot, 3), "cycles per QP:", tot / 1024)
names = ["norms + sqrt/rcp", "barrier after norms", "scale A", "vector updates", "block reduce", "cost scaling"]
rz = out[8:14]
print("inside Ruiz (share of the kernel):", {n: round(v / tot, 4) for n, v in zip(names, rz)}, "sum", round(rz.sum() / tot, 4), "cycles per pass and QP:", np.round(rz / 1024 / 10))
m.close()
