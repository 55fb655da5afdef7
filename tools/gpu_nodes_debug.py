#!/usr/bin/env python
"""GPU box: linearisation nodes of fast vehicles (Ux = 14.9, close to V_max and above the power-limit speed) against the oracle, node by node."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as o
import pigeon.jl_b200 as p
np.set_printoptions(linewidth=250, precision=6, suppress=False)
B = 64
trajs = p.synthetic.synthetic_trajectories(n_traj=4, n_nodes=300)
tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
other = np.tile([1e4, 1e4, 0.0, 5.0], (B, 1))
state = state.copy(); state[::7, 3] = 14.9; state[3::11, 4] += 1.5
g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
g.set_state(state, control, other)
g.compute_time_steps(t0); g.compute_linearization_nodes()
qs, us, ps = g.nodes()
for i in (0, 7, 14, 21, 63, 3):
    m = o.Mpc(0)
    m.set_trajectory(o.Trajectory(**{k: trajs[k][int(tid[i])] for k in o.TRAJ_FIELDS}))
    m.set_state(state[i], control[i], other4=other[i])
    m.compute_time_steps(t0[i]); m.compute_linearization_nodes()
    qo, uo, po = m.nodes()
    dq, du, dp = np.abs(qs[i] - qo), np.abs(us[i] - uo), np.abs(ps[i] - po)
    print(f"v{i}: state {state[i]} s_end {trajs['s'][tid[i]][-1]:.1f}")
    bad = np.where((dq.max(axis=1) > 1e-9) | (du.max(axis=1) > 1e-6) | (dp.max(axis=1) > 1e-9))[0]
    print("   first bad nodes:", bad[:6])
    for k in bad[:3]:
        print(f"   node {k}: q gpu {qs[i, k]} \n            q orc {qo[k]}\n            u gpu {us[i, k]} orc {uo[k]}  p gpu {ps[i, k]} orc {po[k]}")
g.close()
