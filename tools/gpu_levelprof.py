#!/usr/bin/env python
"""GPU box: per-phase cycle profile of the ADMM kernel (phase counters summed over all CTAs, solve-phase counters from CTA 0)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pigeon.jl_b200 as p
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
trajs = p.synthetic.synthetic_trajectories(n_traj=8, n_nodes=400)
tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
g.set_state(state, control, np.tile([1e4, 1e4, 0, 5.0], (B, 1)))
for k in range(6):
    g.step(t0 + 0.01 * k); g.rollout(0.01)
g.set_profiling(2)
g.admm_cycles(reset=True)
g.step(t0 + 0.06)
it = g.stats()["iters"]
c = g.admm_cycles()
lv = g.level_cycles
print("program:", g.qp_program)
print("iters: mean %.1f min %d max %d; rho updates mean %.2f" % (it.mean(), it.min(), it.max(), g.stats()["rho_updates"].mean()))
tot = sum(c.values())
print("phases (cycles per QP, all CTAs):", {k: int(v / B) for k, v in c.items()}, "total per QP", int(tot / B))
n_solves = it[:].mean()   # CTA 0 handles ~B/148 vehicles; normalise by their mean iteration count
per = B / min(B, 148)
print("solve phases of CTA 0 (cycles per KKT solve, assuming %.1f vehicles x %.1f iterations):" % (per, n_solves))
for l in range(0, 110):
    if lv[l] > 0:
        print("  phase %3d %8.0f" % (l, lv[l] / (per * n_solves)))
print("  sum per solve %.0f" % (lv[:110].sum() / (per * n_solves)))
print("factor of CTA 0 (cycles per factorisation, %.1f vehicles):" % per)
print("  init %.0f  levels %.0f  range inverses %.0f  tail %.0f" % (lv[110] / per, lv[120:220].sum() / per, lv[111] / per, lv[112] / per))
print("  per level:", [int(x / per) for x in lv[120:120 + min(g.n_levels, 100)]])
print("probe of forward phase 1 on warp 0 (cycles per solve): task load %.0f  gather loop %.0f  reduce %.0f  epilogue %.0f  loop exit %.0f  barrier %.0f" % tuple(lv[200 + i] / (per * n_solves) for i in range(6)))
