#!/usr/bin/env python
"""GPU box: ADMM cut at max_iter — per-vehicle deviation of the GPU solution from the oracle's, with statuses / iteration counts / rho,
on the perturbed batch of tests/test_gpu_boundary.py::test_max_iter_runs_the_approximate_tests, stage by stage."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as o
import pigeon.jl_b200 as p
B = 64
trajs = p.synthetic.synthetic_trajectories(n_traj=4, n_nodes=300)
tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
other = np.tile([1e4, 1e4, 0.0, 5.0], (B, 1))
state = state.copy(); state[::7, 3] = 14.9; state[3::11, 4] += 1.5
for kw in (dict(max_iter=40), dict(max_iter=4000), dict(max_iter=50)):
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid, **kw)
    g.set_state(state, control, other)
    g.compute_time_steps(t0); g.compute_linearization_nodes()
    qs, us, ps = g.nodes()
    g.update_QP(); qd = g.qp_data(); g.solve(); ug = g.get_next_control()
    st = g.stats(); xg, yg = g.solution()
    print(kw)
    for i in range(B):
        m = o.Mpc(0, settings=o.osqp_settings_default(**kw))
        m.set_trajectory(o.Trajectory(**{k: trajs[k][int(tid[i])] for k in o.TRAJ_FIELDS}))
        m.set_state(state[i], control[i], other4=other[i])
        m.compute_time_steps(t0[i]); m.compute_linearization_nodes()
        qo, uo_, po = m.nodes()
        m.update_qp(); pc = m.qp_pieces(); m.solve(); uo = m.get_next_control()
        so = m.stats(); xo, yo = m.solution()
        dn = max(np.max(np.abs(qs[i] - qo)), np.max(np.abs(us[i] - uo_) / np.array([0.3, 1e4])), np.max(np.abs(ps[i] - po)))
        dq = max(np.max(np.abs(qd[k][i] - pc[k])) / max(1.0, np.max(np.abs(pc[k]))) for k in ("A", "B0", "Bf", "c", "H", "G", "dmin", "dmax", "fxmax"))
        du = np.nanmax(np.abs(ug[i] - uo) / np.array([0.314, 16793.7, 16793.7])) if np.isfinite(uo).all() else float("nan")
        flag = " <<<" if (du > 1e-4 or st["status"][i] != so["status"] or st["iters"][i] != so["iter"]) else ""
        if flag or i < 4:
            print(f"  v{i}: status {st['status'][i]}/{so['status']} iters {st['iters'][i]}/{so['iter']} rho_upd {st['rho_updates'][i]}/{so['rho_updates']} rho {st['rho'][i]:.6g}/{so['rho']:.6g}"
                  f" nodes {dn:.2e} qpdata {dq:.2e} max|dx| {np.nanmax(np.abs(xg[i]-xo)):.3e} du {du:.3e}{flag}")
    g.close()
