#!/usr/bin/env python
"""Development check (GPU box): runs every stage of the hot path on a small batch and prints max deviations from the CPU oracle."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as o  # noqa: E402
import pigeon.jl_b200 as p  # noqa: E402
from pigeon.jl_b200 import synthetic  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / (1e-12 + np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0


def abserr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) if np.asarray(a).size else 0.0


def make_oracles(kind, trajs, tid, state, control, other=None, hji=None, **kw):
    ms = []
    tr_cache = {}
    for i in range(len(tid)):
        j = int(tid[i])
        if j not in tr_cache:
            tr_cache[j] = o.Trajectory(**{k: trajs[k][j] for k in o.TRAJ_FIELDS})
        m = o.Mpc(kind, **kw)
        m.set_trajectory(tr_cache[j])
        if hji is not None:
            m.set_hji(hji)
        m.set_state(state[i], control[i], other4=None if other is None else other[i])
        m._tr = tr_cache[j]
        ms.append(m)
    return ms


def section(name):
    print(f"\n=== {name} ===", flush=True)


def run(kind, B=64, steps=4, Ns=10, Nl=20):
    kname = "coupled" if kind == 0 else "decoupled"
    section(f"{kname} B={B} N_short={Ns} N_long={Nl}")
    trajs = synthetic.synthetic_trajectories(n_traj=8, n_nodes=400)
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B)
    other = np.tile(np.array([1e4, 1e4, 0.0, 5.0]), (B, 1))
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if kind == 0 else p.BatchedDecoupledTrajectoryTrackingMPC
    g = ctor(p.X1(), trajs, B, N_short=Ns, N_long=Nl, trajectory_index=tid)
    print("dims", dict(N=g.N, n=g.n, m=g.m, nnzA=g.nnzA, nnzL=g.nnzL, levels=g.n_levels))
    g.set_state(state, control, other)
    ms = make_oracles(kind, trajs, tid, state, control, other, N_short=Ns, N_long=Nl)
    for k in range(steps):
        tk = t0 + 0.01 * k
        tt = time.time()
        g.compute_time_steps(tk); g.compute_linearization_nodes(); g.update_QP(); g.solve(); ug = g.get_next_control()
        g.synchronize()
        tg = time.time() - tt
        for i, m in enumerate(ms):
            m.compute_time_steps(tk[i]); m.compute_linearization_nodes(); m.update_qp(); m.solve()
        uo = np.array([m.get_next_control() for m in ms])
        ts_g, dt_g, _ = g.time_steps()
        ts_o = np.array([m.time_steps()[0] for m in ms])
        qs_g, us_g, ps_g = g.nodes()
        no = [m.nodes() for m in ms]
        qs_o, us_o, ps_o = (np.array([x[j] for x in no]) for j in range(3))
        d = g.qp_data()
        po = [m.qp_pieces() for m in ms]
        un = g.u_normalization if kind == 0 else np.array([1.0, 1.0])
        A_o = np.array([x["A"] for x in po]); c_o = np.array([x["c"] for x in po])
        B0_o = np.array([x["B0"] for x in po]) * un[None, None, None, :g.nu]; Bf_o = np.array([x["Bf"] for x in po]) * un[None, None, None, :g.nu]
        H_o = np.array([x["H"] for x in po]); G_o = np.array([x["G"] for x in po])
        xg, yg = g.solution()
        so = [m.solution() for m in ms]
        xo = np.array([s[0] for s in so]); yo = np.array([s[1] for s in so])
        st = g.stats()
        it_o = np.array([m.stats()["iter"] for m in ms]); st_o = np.array([m.stats()["status"] for m in ms]); rho_o = np.array([m.stats()["rho"] for m in ms])
        print(f"step {k}: gpu wall {tg*1e3:.1f} ms | ts {abserr(ts_g, ts_o):.1e} | nodes q {abserr(qs_g, qs_o):.1e} u {relerr(us_g, us_o):.1e} p {abserr(ps_g, ps_o):.1e}"
              f" | A {abserr(d['A'], A_o):.1e} B0 {abserr(d['B0'], B0_o):.1e} Bf {abserr(d['Bf'], Bf_o):.1e} c {abserr(d['c'], c_o):.1e} H {relerr(d['H'], H_o):.1e} G {relerr(d['G'], G_o):.1e}"
              f" | x {abserr(xg, xo):.1e} y {abserr(yg, yo):.1e} | u {abserr(ug[:, 0], uo[:, 0]):.1e} rad, {abserr(ug[:, 1:], uo[:, 1:]):.1e} N"
              f" | iters gpu {st['iters'].mean():.1f} (max {st['iters'].max()}) oracle {it_o.mean():.1f} mismatches {(st['iters'] != it_o).sum()}"
              f" status!=: {(st['status'] != st_o).sum()} rho {relerr(st['rho'], rho_o):.1e}", flush=True)
        # closed loop: propagate both with their own controls
        g.rollout(0.01)
        for i, m in enumerate(ms):
            q, u = m.get_state()
            xn = o.flow(o.MODEL_BICYCLE, m.vp, q, 0.01, [u[0], u[1] + u[2], 0, 0, 0, 0])
            m.set_state(xn, uo[i], other4=other[i])
        qg, ucur = g.get_state()
        qo = np.array([m.get_state()[0] for m in ms])
        print(f"        after rollout: state {abserr(qg, qo):.1e} control {abserr(ucur, uo):.1e}")
    g.close()


def hji_check():
    section("HJI lookup + constraint")
    dims = (7, 6, 5, 5, 4, 5, 4)
    knots, V, gV = synthetic.analytic_hji_grid(dims)
    cache_o = o.HjiCache(knots, V, gV)
    B = 256
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), p.straight_trajectory(30., 5.), B)
    g.set_HJI_cache(p.HJICache(knots, V, gV))
    rng = np.random.default_rng(0)
    lo = np.array([k[0] for k in knots], float); hi = np.array([k[-1] for k in knots], float)
    x = lo + (hi - lo) * rng.random((B, 7))
    x[:20] += (hi - lo) * 0.6 * (rng.random((20, 7)) < 0.3)    # some out of grid
    Vg, gg = g.hji_lookup(x)
    Vo, go = cache_o.lookup(x)
    fin = np.isfinite(Vo)
    print("lookup: inf mismatch", int((np.isfinite(Vg) != fin).sum()), "V", abserr(Vg[fin], Vo[fin]), "gradV", abserr(gg, go))
    # constraint through update_QP
    state = np.zeros((B, 6)); state[:, 2] = rng.uniform(-3, 3, B); state[:, 3] = rng.uniform(2, 12, B); state[:, 4] = rng.normal(0, 0.3, B); state[:, 5] = rng.normal(0, 0.2, B)
    other = np.zeros((B, 4)); other[:, 0] = rng.uniform(-12, 12, B); other[:, 1] = rng.uniform(-12, 12, B); other[:, 2] = rng.uniform(-3, 3, B); other[:, 3] = rng.uniform(1.5, 12, B)
    control = np.zeros((B, 3)); control[:, 0] = rng.normal(0, 0.05, B); control[:, 2] = rng.uniform(-500, 1500, B)
    g.set_state(state, control, other)
    g.compute_time_steps(0.0); g.compute_linearization_nodes(); g.update_QP()
    hg = g.qp_data()["hji"]
    vp = o.x1(); un = g.u_normalization
    ho = np.zeros((B, 3)); act = 0
    for i in range(B):
        x7 = o.hji_relative_state(state[i], other[i])
        M, b = o.reachability_constraint(vp, cache_o, x7, 0.05, [control[i, 0], control[i, 1] + control[i, 2]])
        ho[i] = [M[0] * un[0], M[1] * un[1], b]
        act += int(not (M[0] == 0 and M[1] == 0 and b == 1.0))
    print("constraint: active", act, "of", B, "| M,b abs err", abserr(hg, ho), "rel", relerr(hg, ho))
    g.close()


if __name__ == "__main__":
    for fn, args in [(run, (0, 64, 4)), (run, (1, 64, 4)), (run, (0, 32, 3, 5, 10)), (hji_check, ())]:
        try:
            fn(*args)
        except Exception:
            traceback.print_exc()
