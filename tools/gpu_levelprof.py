#!/usr/bin/env python
"""GPU box: per-level cycle profile of the ADMM triangular solves (CTA 0), after a few warm steps."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigeon.jl_b200 as p
B = 148
trajs = p.synthetic.synthetic_trajectories(n_traj=8, n_nodes=400)
tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
g.set_state(state, control, np.tile([1e4, 1e4, 0, 5.0], (B, 1)))
for k in range(6):
    g.step(t0 + 0.01 * k); g.rollout(0.01)
g.set_profiling(1)
g.step(t0 + 0.06)
it = g.stats()["iters"]
c = g.admm_cycles()
lv = g.level_cycles
n_it = it[0] if True else 0
print("phases (all CTAs, cycles):", {k: int(v) for k, v in c.items()}, "iters CTA0 vehicle:", it[:4], "mean", it.mean())
# which vehicle did CTA 0 process? unknown with the ticket; normalise per solve using total solves = sum of per-level counts is not available -> print raw and per-iteration using mean iters
print("forward levels (cycles per solve, assuming %d iterations):" % it.mean())
nit = it.mean()
for l in range(1, 60):
    if lv[l] > 0: print("  fwd l=%2d %8.0f" % (l, lv[l] / nit))
print("  tail stages", [round(lv[100 + i] / nit) for i in range(3)])
for l in range(59, -1, -1):
    if lv[128 + l] > 0: print("  bwd l=%2d %8.0f" % (l, lv[128 + l] / nit))
print("sum per solve", (lv.sum()) / nit)
