#!/usr/bin/env python
"""GPU box: find the vehicles of the Monte-Carlo workload whose closed loop goes non-finite, when, and whether the CPU oracle does the same."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pigeon.jl_b200 as p
from pigeon.jl_b200 import synthetic
import oracle_py as o
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
trajs = synthetic.synthetic_trajectories(seed=synthetic.SEED, n_traj=64, n_nodes=1000, ds=0.25)
tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 2)
far = np.tile([1e4, 1e4, 0, 5.0], (B, 1))
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
m.set_state(state, control, far)
first_bad = np.full(B, -1)
hist = {}
for k in range(200):
    u = m.step(t0 + 0.01 * k)
    st = m.stats()
    m.rollout(0.01)
    q, _ = m.get_state()
    bad = ~np.isfinite(q).all(axis=1) | ~np.isfinite(u).all(axis=1)
    new = bad & (first_bad < 0)
    for i in np.nonzero(new)[0]:
        first_bad[i] = k
        hist[i] = (k, int(st["iters"][i]), int(st["status"][i]), u[i].copy(), q[i].copy())
print("B", B, "non-finite vehicles", int((first_bad >= 0).sum()), "first steps histogram", np.bincount(first_bad[first_bad >= 0], minlength=1)[:40])
ids = list(hist)[:6]
for i in ids:
    k, it, stt, u, q = hist[i]
    print("vehicle", i, "traj", tid[i], "first bad step", k, "iters", it, "status", stt, "u", u, "q", q, "init state", state[i], "init control", control[i], "t0", t0[i])
    om = o.Mpc(o.MPC_COUPLED)
    om.set_trajectory(o.Trajectory(**{kk: trajs[kk][int(tid[i])] for kk in o.TRAJ_FIELDS}))
    om.set_state(state[i], control[i], other4=far[i])
    for kk in range(k + 1):
        om.simulate_step(t0[i] + 0.01 * kk, 0.01)
        s = om.stats()
        if kk >= k - 2:
            qq, uu = om.get_state()
            print("   oracle step", kk, "iters", s["iter"], "status", s.get("status"), "u", uu, "q", qq)
