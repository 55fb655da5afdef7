"""CPU tests of the oracle's MPC step (time steps -> nodes -> QP -> OSQP -> control) on the reference's smoke scenario
(src/Pigeon.jl:34-57) and its path fixtures (test/path/*.world)."""
import numpy as np
import pytest

import oracle_py as o
from helpers import world_trajectory

FAR = [1e4, 1e4, 0.0, 5.0]    # other car far outside the HJI grid => constraint inactive


def test_qp_dimensions_match_reference_construction():
    # counted from construct_coupled_tracking_QP / construct_lateral_tracking_QP
    m = o.Mpc(o.MPC_COUPLED)
    assert (m.N, m.n, m.m, m.nnzA) == (31, 378, 691, 2751)
    m = o.Mpc(o.MPC_COUPLED, N_short=5, N_long=10)
    assert (m.N, m.n, m.m, m.nnzA) == (16, 193, 351, 1381)
    m = o.Mpc(o.MPC_DECOUPLED)
    assert (m.N, m.n, m.m, m.nnzA) == (31, 245, 455, 1435)


def test_smoke_scenario_straight_line():
    # Pigeon.jl:34-39: straight 30 m @ 5 m/s, state (0,0,0,5,0,0), zero control
    for kind in (o.MPC_COUPLED, o.MPC_DECOUPLED):
        m = o.Mpc(kind, N_short=5, N_long=10) if kind == o.MPC_COUPLED else o.Mpc(kind)
        m.set_state([0, 0, 0, 5, 0, 0], [0, 0, 0], other4=FAR)
        u = m.step(0.0)
        st = m.stats()
        assert st["status"] == 1 and st["iter"] % 25 == 0
        assert abs(u[0]) < 1e-6                       # delta ~ 0
        assert u[1] == 0 and u[2] > 0                 # drive force goes to the rear axle (fwd_frac = 0)
        if kind == o.MPC_DECOUPLED:
            assert u[2] == pytest.approx(241 + 25.1 * 5, rel=0.05)    # drag equilibrium 366.5 N (feed-forward)
        qs, us, ps = m.nodes()
        assert np.allclose(ps[:, 0] if kind == o.MPC_COUPLED else ps[:, 0], 5.0, atol=0.05)
        assert np.all(ps[:, 1] == 0)


def test_coupled_drag_equilibrium_in_closed_loop():
    m = o.Mpc(o.MPC_COUPLED, N_short=5, N_long=10)
    tr = o.Trajectory(t=[0, 60.0], s=[0, 300.0], V=[5, 5], A=[0, 0], E=[0, 0], N=[0, 300.0], psi=[0, 0], kappa=[0, 0])
    m.set_trajectory(tr)
    m.set_state([0, 0, 0, 5, 0, 0], [0, 0, 0], other4=FAR)
    for k in range(400):
        m.simulate_step(0.01 * k)
    q, u = m.get_state()
    assert u[1] + u[2] == pytest.approx(366.5, rel=0.08)
    assert q[3] == pytest.approx(5.0, abs=0.05) and abs(q[0]) < 1e-6


def test_qp_solution_is_near_the_true_optimum():
    tr = world_trajectory("skidpadoval")
    m = o.Mpc(o.MPC_COUPLED)
    m.set_trajectory(tr)
    f = tr.fields
    k = 420   # inside the first curve
    m.set_state([f["E"][k] + 0.2, f["N"][k] - 0.1, f["psi"][k] + 0.03, 6.2, 0.05, 0.3], [0.1, 0, 300.0], other4=FAR)
    m.compute_time_steps(f["t"][k])
    m.compute_linearization_nodes()
    m.update_qp()
    m.solve()
    x, y = m.solution()
    qp = m.qp()
    import scipy.sparse as sp
    hi = o.Osqp(sp.diags(qp["Pdiag"]).tocsc(), qp["q"], qp["A"], qp["l"], qp["u"], o.osqp_settings_default(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000))
    xs, ys, info = hi.solve()
    assert info["status"] == 1
    un = [m.vp[20], max(-m.vp[18], m.vp[17])]
    iu = 6 * 31 + 2 * 1
    # at eps = 1e-3 the ADMM iterate is within ~1e-2 of the optimum in the normalised controls
    assert abs(x[iu] - xs[iu]) < 2e-2 and abs(x[iu + 1] - xs[iu + 1]) < 2e-2
    # equality rows hold: q1 = q_curr, dynamics residual small at the high-accuracy solution
    Ax = qp["A"] @ xs
    eq = qp["l"] == qp["u"]
    assert eq.sum() == 248
    assert np.max(np.abs(Ax[eq] - qp["l"][eq])) < 1e-6


@pytest.mark.parametrize("kind,corrected,tol", [(o.MPC_COUPLED, 0, 0.1), (o.MPC_DECOUPLED, 1, 0.3), (o.MPC_COUPLED, 1, 0.1)])
def test_closed_loop_tracks_skidpad_curve(kind, corrected, tol):
    tr = world_trajectory("skidpadoval")
    f = tr.fields
    vp = o.x1()
    vp[22] = corrected
    m = o.Mpc(kind, vp=vp)
    m.set_trajectory(tr)
    k0 = 180   # on the straight, 10 m before the first curve (kappa ramps up from s = 53 m)
    m.set_state([f["E"][k0], f["N"][k0], f["psi"][k0], 6, 0, 0], [0, 0, 0], other4=FAR)
    t0 = f["t"][k0]
    emax, rmax, iters = 0.0, 0.0, []
    for k in range(800):
        m.simulate_step(t0 + 0.01 * k)
        st = m.stats()
        assert st["status"] in (1, 2)
        iters.append(st["iter"])
        q, u = m.get_state()
        s, e, _ = tr.path_coordinates(q[0], q[1])
        emax = max(emax, abs(e)); rmax = max(rmax, q[5])
    assert s > f["s"][k0] + 40.0           # made progress through the curve (6 m/s * 8 s)
    assert emax < tol
    assert rmax > 0.35                     # yawed with the curve (kappa*V ~ 0.42 rad/s)
    assert np.mean(iters) < 120


def test_decoupled_literal_inverse_fiala_quirk_is_reproduced():
    """vehicle_dynamics.jl:56-62 returns the slip ratio instead of tan(alpha) in the unsaturated branch; restated literally
    (inv_fiala_corrected = 0, the default) the decoupled controller's steady-state nodes sit near tire saturation and
    it tracks the skidpad curve poorly.  This pins the literal behaviour."""
    vp = o.x1()
    est_lit = o.steady_state(vp, 6.0, 0.0, 0.0693)
    vp2 = vp.copy(); vp2[22] = 1
    est_fix = o.steady_state(vp2, 6.0, 0.0, 0.0693)
    L = vp[0]
    assert est_fix["delta"] == pytest.approx(np.arctan(L * 0.0693), abs=0.02)      # ~ Ackermann + small understeer
    assert abs(est_fix["beta"]) < 0.1
    ar_lit = np.arctan2(est_lit["Uy"] - vp[2] * est_lit["r"], est_lit["Ux"])
    ar_fix = np.arctan2(est_fix["Uy"] - vp[2] * est_fix["r"], est_fix["Ux"])
    assert abs(ar_lit) > 5 * abs(ar_fix)                                            # literal nodes: rear slip angle ~ slip ratio


def test_warm_nodes_interpolate_previous_solution():
    tr = world_trajectory("skidpadoval")
    f = tr.fields
    m = o.Mpc(o.MPC_COUPLED)
    m.set_trajectory(tr)
    m.set_state([f["E"][400], f["N"][400], f["psi"][400], 6, 0, 0.3], [0.08, 0, 390.0], other4=FAR)
    m.step(f["t"][400])
    x, _ = m.solution()
    ts_prev, _, _ = m.time_steps()
    m.compute_time_steps(f["t"][400] + 0.01)
    m.compute_linearization_nodes()
    qs, us, ps = m.nodes()
    ts, _, prev = m.time_steps()
    assert np.array_equal(prev, ts_prev)
    un = np.array([m.vp[20], max(-m.vp[18], m.vp[17])])
    X = x[:186].reshape(31, 6); U = x[186:248].reshape(31, 2)
    for i in (1, 5, 12, 30):
        tq = min(ts[i], prev[-1])
        for c in range(6):
            assert qs[i, c] == pytest.approx(np.interp(tq, prev, X[:, c]), rel=1e-12, abs=1e-14)
        for c in range(2):
            assert us[i, c] == pytest.approx(np.interp(tq, prev, U[:, c]) * un[c], rel=1e-12, abs=1e-12)


def test_from_autobox_callback_semantics():
    """from_autobox_callback (ros_integration.jl:48-151): time selection, early returns, reply fields."""
    tr = world_trajectory("skidpadoval")
    f = tr.fields
    q = [f["E"][400], f["N"][400] + 0.2, f["psi"][400], 6, 0, 0.3]
    u = [0.08, 0, 390.0]
    s_e_t = tr.path_coordinates(q[0], q[1])
    # path-tracking mode (time_offset NaN): the MPC time is the path_coordinates time; reply carries (s, e)
    m = o.Mpc(o.MPC_COUPLED); m.set_trajectory(tr); m.set_state(q, u, other4=FAR)
    pub, out = m.from_autobox(q, u, stamp=123.0, pause_speed=1.0)
    assert pub and out[3] == s_e_t[0] and out[4] == s_e_t[1]
    ref = o.Mpc(o.MPC_COUPLED); ref.set_trajectory(tr); ref.set_state(q, u, other4=FAR)
    assert np.array_equal(out[:3], ref.step(s_e_t[2]))
    # trajectory mode: t = stamp - time_offset; outside [0, t_end] the callback returns early and the control stays
    m2 = o.Mpc(o.MPC_COUPLED); m2.set_trajectory(tr); m2.set_state(q, u, other4=FAR, time_offset=100.0)
    pub, out = m2.from_autobox(q, u, stamp=99.0)
    assert not pub and np.array_equal(out[:3], u) and m2.stats()["iter"] == 0
    pub, out = m2.from_autobox(q, u, stamp=100.0 + f["t"][-1] + 1.0)
    assert not pub
    pub, out = m2.from_autobox(q, u, stamp=100.0 + f["t"][400])
    ref2 = o.Mpc(o.MPC_COUPLED); ref2.set_trajectory(tr); ref2.set_state(q, u, other4=FAR, time_offset=100.0)
    assert pub and np.array_equal(out[:3], ref2.step(f["t"][400]))
    # paused below 1 m/s
    qs = list(q); qs[3] = 0.4
    pub, out = m2.from_autobox(qs, u, stamp=100.0 + f["t"][400], pause_speed=1.0)
    assert not pub and np.array_equal(out[:3], u)
