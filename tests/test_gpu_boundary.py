"""GPU tests of the round-2 boundary work (run with `-m gpu`), all through the C ABI:
   pause -> resume parity of the callback guards, the on-device history recorder of `simulate`, masked resets on the handle's stream,
   transactional pgn_set_control_params, the approximate (x10) termination tests at max_iter, multi-device handles driven by one host
   thread and the NCCL gather (pgn_comm_init_all / pgn_gather_all)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle_py as o  # noqa: E402

pytestmark = pytest.mark.gpu

FAR = np.array([1e4, 1e4, 0.0, 5.0])
U_RANGE = np.array([0.3141592653589793, 16793.73299576057, 16793.73299576057])


@pytest.fixture(scope="module")
def p():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pigeon.jl_b200 as pkg
    pkg.load()
    return pkg


def batch(p, B, n_traj=2):
    trajs = p.synthetic.synthetic_trajectories(n_traj=n_traj, n_nodes=300)
    tid, state, control, t0 = p.synthetic.synthetic_batch(trajs, B)
    return trajs, tid, state, control, t0, np.tile(FAR, (B, 1))


def oracle_for(trajs, tid, i, kind=0, **kw):
    m = o.Mpc(kind, **kw)
    m.set_trajectory(o.Trajectory(**{k: trajs[k][int(tid[i])] for k in o.TRAJ_FIELDS}))
    return m


def test_pause_then_resume_matches_the_callback(p):
    """The reference's callback returns BEFORE compute_time_steps! when Ux < 1 (src/ros_integration.jl:84-87), so prev_ts keeps the knots of
    the last SOLVED QP and a resuming vehicle interpolates its warm nodes on them.  Sequence per vehicle: solve, solve, pause, pause,
    resume (warm), solve — against the oracle's restatement of the callback, with identical iteration counts."""
    B = 12
    trajs, tid, state, control, t0, other = batch(p, B)
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    g.set_guards(nan_fallback=True, pause_below_speed=1.0)
    g.set_state(state, control, other)           # time_offset stays NaN: path-tracking mode, the stamp is not used
    ms = []
    for i in range(B):
        m = oracle_for(trajs, tid, i)
        m.set_state(state[i], control[i], other4=other[i])
        ms.append(m)
    paused_at = {2, 3}                           # callbacks in which vehicles 1, 4, 7 report a speed below 1 m/s
    slow = [1, 4, 7]
    q, u = state.copy(), control.copy()
    for k in range(6):
        qk = q.copy()
        if k in paused_at:
            qk[slow, 3] = 0.4
        out = g.from_autobox(qk, u, 0.0, other_car=other if k == 0 else None)
        it = g.stats()["iters"]
        ts_g, _, pts_g = g.time_steps()
        for i, m in enumerate(ms):
            pub, ref = m.from_autobox(qk[i], u[i], 0.0, pause_speed=1.0, nan_fallback=True)
            assert pub == (not (k in paused_at and i in slow)), (k, i)
            assert np.max(np.abs(out[i, :3] - ref[:3]) / U_RANGE) < 1e-4, (k, i)
            ts_o, _, pts_o = m.time_steps()
            assert np.allclose(ts_g[i], ts_o, rtol=0, atol=1e-12) and np.allclose(pts_g[i], pts_o, rtol=0, atol=1e-12), (k, i)
            if pub:
                assert it[i] == m.stats()["iter"], (k, i, it[i], m.stats()["iter"])
        u = out[:, :3].copy()
        # the vehicles move a little along their path between callbacks
        q = q.copy(); q[:, 0] += -np.sin(q[:, 2]) * q[:, 3] * 0.01; q[:, 1] += np.cos(q[:, 2]) * q[:, 3] * 0.01
    g.close()


@pytest.mark.parametrize("kind", ["coupled", "decoupled"])
def test_history_recorder_matches_the_host_loop(p, kind):
    """simulate returns (qs, xs, us, ps) per step (model_predictive_control.jl:84-99).  The on-device recorder of pgn_simulate must give
    exactly what a host loop over the five calls sees: state / control before the step, mpc.qs[1], mpc.ps[1] of the step."""
    B = 37
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if kind == "coupled" else p.BatchedDecoupledTrajectoryTrackingMPC
    a, b = ctor(p.X1(), trajs, B, trajectory_index=tid), ctor(p.X1(), trajs, B, trajectory_index=tid)
    a.set_state(None, None, other); b.set_state(None, None, other)
    n = 9
    qs, xs, us, ps = p.simulate(a, state, control, 0.01, t0=t0, n_steps=n, record=True, stride=2)
    assert qs.shape == (5, B, 6) and xs.shape == (5, B, a.nx) and us.shape == (5, B, 3) and ps.shape == (5, B, 4)
    b.set_state(state, control)
    for k in range(n):
        q, u = b.get_state()
        b.compute_time_steps(t0 + k * 0.01); b.compute_linearization_nodes()
        nq, _, npar = b.nodes()
        b.update_QP(); b.solve(); b.get_next_control(); b.rollout(0.01)
        if k % 2 == 0:
            r = k // 2
            assert np.array_equal(qs[r], q) and np.array_equal(us[r], u), k
            assert np.array_equal(xs[r], nq[:, 0]) and np.array_equal(ps[r], npar[:, 0]), k
    qa, ua = a.get_state(); qb, ub = b.get_state()
    assert np.array_equal(qa, qb) and np.array_equal(ua, ub)
    # default horizon of simulate: 0:dt:trajectory.t[end]
    assert abs(a.trajectory_end_time - trajs["t"][0, -1]) == 0
    a.close(); b.close()


def test_masked_resets_touch_only_the_masked_vehicles(p):
    B = 20
    trajs, tid, state, control, t0, other = batch(p, B)
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    ref = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    for m in (g, ref):
        m.set_state(state, control, other)
        m.step(t0); m.rollout(0.01)
    mask = np.zeros(B, np.uint8); mask[[3, 4, 11]] = 1
    g.reset_solver(mask); g.reset_solved(mask)
    u_g, u_r = g.step(t0 + 0.01), ref.step(t0 + 0.01)
    keep = mask == 0
    assert np.array_equal(u_g[keep], u_r[keep])                       # untouched vehicles: bit-identical warm step
    assert np.array_equal(g.stats()["iters"][keep], ref.stats()["iters"][keep])
    # the masked vehicles behave like freshly constructed controllers at the same state
    fresh = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    q0, c0 = g.get_state()
    fresh.set_state(q0, c0, other)
    u_f = fresh.step(t0 + 0.01)
    assert np.array_equal(u_g[mask == 1], u_f[mask == 1])
    assert np.array_equal(g.stats()["iters"][mask == 1], fresh.stats()["iters"][mask == 1])
    g.close(); ref.close(); fresh.close()


def test_set_control_params_is_transactional(p):
    B = 6
    trajs, tid, state, control, t0, other = batch(p, B)
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    h = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    bad = p.CoupledControlParams(N_HJI=99, k_V=123.0, V_min=3.0)
    with pytest.raises(p.PigeonError):
        g.set_control_params(bad)
    for m in (g, h):
        m.set_state(state, control, other)
    g.control_params = dict(h.control_params)
    assert np.array_equal(g.step(t0), h.step(t0))                     # nothing of the rejected parameter set was applied
    g.close(); h.close()


def test_max_iter_runs_the_approximate_tests(p):
    """osqp_solve ends with check_termination(work, approximate = 1): the last iteration is always checked, rho is adapted if the interval
    falls on it (the value is the next step's warm-start rho), then the optimality test and the infeasibility certificates run with every
    tolerance x 10.  Cutting a cold start short gives a mix of solved_inaccurate / max_iter_reached: statuses, iteration counts, rho, its
    update count and the (unconverged) controls must equal the oracle's for every vehicle.
    (States are left inside the tire-friction limit on purpose: at the limit the reference's steady_state_estimates puts the front force
    exactly ON the friction circle and `_invfialatiremodel` branches on |Fy| >= Fy_max, a comparison decided by the last bit of sincos —
    tools/gpu_nodes_debug.py; no two libm implementations agree there.)"""
    B = 48
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)
    seen = set()
    for kw in (dict(max_iter=24), dict(max_iter=26), dict(max_iter=40), dict(max_iter=50), dict(max_iter=37, check_termination=0, adaptive_rho_interval=10)):
        g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid, **kw)
        g.set_state(state, control, other)
        ug = g.step(t0)
        st = g.stats()
        for i in range(B):
            m = oracle_for(trajs, tid, i, settings=o.osqp_settings_default(**kw))
            m.set_state(state[i], control[i], other4=other[i])
            uo = m.step(t0[i])
            so = m.stats()
            assert st["status"][i] == so["status"] and st["iters"][i] == so["iter"], (kw, i, st["status"][i], so["status"], st["iters"][i], so["iter"])
            assert st["rho_updates"][i] == so["rho_updates"] and abs(st["rho"][i] - so["rho"]) <= 1e-6 * so["rho"], (kw, i, st["rho"][i], so["rho"])
            seen.add(int(so["status"]))
            assert np.max(np.abs(ug[i] - uo) / U_RANGE) < 1e-4, (kw, i)
        g.close()
    assert {2, -2} <= seen, seen              # the cuts really produced solved_inaccurate and max_iter_reached


def test_one_rank_gather_goes_through_nccl(p):
    """pgn_comm_init_all + pgn_gather_all with a single handle: exercises the dlopen'ed NCCL path on one GPU."""
    B = 10
    trajs, tid, state, control, t0, other = batch(p, B)
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    g.set_state(state, control, other)
    u = g.step(t0)
    p.comm_init_all([g])
    c, it, st = p.gather_all([g])
    assert np.array_equal(c, u) and np.array_equal(it, g.stats()["iters"]) and np.array_equal(st, g.stats()["status"])
    g.close()


def test_two_devices_one_host_thread_and_gather(p):
    """The Julia deployment: ONE host thread, one handle per GPU (every entry point switches to its handle's device), batch sharded in
    contiguous ranges, final gather over NCCL.  Must equal one handle holding the whole batch."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    B = 48
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)
    whole = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid, device=0)
    whole.set_state(state, control, other)
    halves = []
    for r in range(2):
        sl = slice(r * B // 2, (r + 1) * B // 2)
        m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B // 2, trajectory_index=tid[sl], device=r)
        m.set_state(state[sl], control[sl], other[sl])
        halves.append((m, sl))
    torch.cuda.set_device(0)                      # the caller's current device is NOT the second handle's
    p.comm_init_all([m for m, _ in halves])
    for k in range(3):
        uw = whole.step(t0 + 0.01 * k); whole.rollout(0.01)
        for m, sl in halves:
            us = m.step(t0[sl] + 0.01 * k); m.rollout(0.01)
            assert np.array_equal(us, uw[sl]), (k, sl)
        assert torch.cuda.current_device() == 0
    c, it, st = p.gather_all([m for m, _ in halves])
    assert np.array_equal(c, uw) and np.array_equal(it, whole.stats()["iters"]) and np.array_equal(st, whole.stats()["status"])
    whole.close()
    for m, _ in halves:
        m.close()


@pytest.mark.parametrize("kind", ["coupled", "decoupled"])
def test_pipelined_submit_collect_is_bit_identical(p, kind):
    """pgn_step_submit / pgn_step_collect keep up to 4 steps in flight with the pipeline parts never joined; per vehicle the operations are
    those of pgn_set_state + pgn_step, so every step's controls, iteration counts and the final solver state must be bit-identical — for a
    ragged batch, any part count, and with measured states that do NOT depend on the previous output (the callback's situation)."""
    B = 203
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if kind == "coupled" else p.BatchedDecoupledTrajectoryTrackingMPC
    rng = np.random.default_rng(3)
    K = 9
    states = [state + rng.normal(0, 0.02, state.shape) * np.array([1, 1, 0.1, 1, 0.2, 0.05]) for _ in range(K)]
    ctrls = [control * (1 + 0.05 * rng.normal(size=control.shape)) for _ in range(K)]
    ref = ctor(p.X1(), trajs, B, trajectory_index=tid)
    ref.set_state(state, control, other)
    want = []
    for k in range(K):
        ref.set_state(states[k], ctrls[k])
        want.append((ref.step(t0 + 0.01 * k), ref.stats()["iters"].copy()))
    for parts in (1, 4, 7):
        g = ctor(p.X1(), trajs, B, trajectory_index=tid)
        g.set_pipeline_parts(parts)
        g.set_state(state, control, other)
        got = []
        for k in range(K):
            g.step_submit(t0 + 0.01 * k, states[k], ctrls[k], other if k == 2 else None)
            if k >= 2:
                got.append(g.step_collect())
        while len(got) < K:
            got.append(g.step_collect())
        with pytest.raises(p.PigeonError):
            g.step_collect()                                           # nothing in flight
        for k in range(K):
            assert np.array_equal(got[k], want[k][0]), (parts, k)
        assert np.array_equal(g.stats()["iters"], want[-1][1])
        xs_g, xs_r = g.solution()[0], ref.solution()[0]
        assert np.array_equal(xs_g, xs_r)
        # a fifth submit without a collect is refused; other entry points wait for the steps in flight
        for k in range(4):
            g.step_submit(t0 + 0.01 * (K + k), states[k], ctrls[k])
        with pytest.raises(p.PigeonError):
            g.step_submit(t0, states[0], ctrls[0])
        it = g.stats()["iters"]                                        # drains the ring, results stay collectable
        last = [g.step_collect() for _ in range(4)]
        assert np.all(np.isfinite(last[-1])) and it.shape == (B,)
        g.close()
    ref.close()


@pytest.mark.parametrize("kind,parts", [("coupled", 1), ("coupled", 4), ("decoupled", 3)])
def test_deferred_solves_are_bit_identical(p, kind, parts):
    """pgn_set_solve_cap: inside the simulate loops a QP that has not terminated after `cap` iterations of one launch continues in the next
    round while only its vehicle waits.  With a cap of 25 or 50 almost every cold-start solve is cut several times; final states, controls,
    solver statistics (total iteration counts included), warm-start state (checked through one more step) and the recorded histories must
    be bit-identical to the uncapped loop — also when the loop is issued in pieces on one time axis, and with the NaN guard on."""
    import torch
    B = 101
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)
    state = state.copy(); state[5, 4] = np.nan                      # one vehicle whose QP returns NaN: guard path (cold re-initialisation)
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if kind == "coupled" else p.BatchedDecoupledTrajectoryTrackingMPC
    n = 12

    def run(cap, pieces):
        g = ctor(p.X1(), trajs, B, trajectory_index=tid)
        g.set_pipeline_parts(parts)
        g.set_guards(nan_fallback=True, pause_below_speed=0.0)
        g.set_solve_cap(cap)
        g.set_state(state, control, other)
        g.set_history(n, 1)
        d = torch.tensor(t0, dtype=torch.float64, device="cuda")
        k = 0
        for m in pieces:
            g.simulate_device_async(d.data_ptr(), 0.01, m, k0=k); k += m
        q, u = g.get_state()                                            # a getter: waits for the vehicles that are behind
        st = g.stats()
        hist = g.history()
        nxt = g.step(t0 + 0.01 * n)                                     # warm-start state carried over correctly
        out = (q, u, st["iters"], st["status"], st["rho"], st["rho_updates"], nxt) + hist
        g.close()
        return out
    ref = run(0, [n])
    assert np.isfinite(ref[0][np.arange(B) != 5]).all() and ref[2].max() > 50          # real closed loop, some long solves
    for cap, pieces in ((25, [n]), (50, [5, 7]), (200, [1] * n)):
        got = run(cap, pieces)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b, equal_nan=True), (cap, pieces)
    with pytest.raises(p.PigeonError):
        g = ctor(p.X1(), trajs, 4, trajectory_index=tid[:4])
        try:
            g.set_solve_cap(30)                                         # not a multiple of check_termination = 25
        finally:
            g.close()


def test_cell_ordered_hji_lookup_is_bit_identical(p):
    """pgn_set_hji_lookup_order: large stand-alone lookups visit the queries in grid-cell order (counting sort by cell, then the same
    interpolation) so that neighbouring queries share their corners through L2.  Values and gradients must equal the input-order kernel bit
    for bit — in-grid, on faces / knots, and outside the grid — and agree with the oracle."""
    dims = (7, 6, 5, 5, 4, 5, 4)
    knots, V, gV = p.synthetic.analytic_hji_grid(dims)
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), [p.straight_trajectory(30.0, 5.0)], 1)
    g.set_HJI_cache(p.HJICache(knots, V, gV))
    rng = np.random.default_rng(11)
    M = 70001                                   # not a multiple of anything (the automatic switch to cell order is at 2^19 queries)
    lo = np.array([r[0] for r in p.synthetic.HJI_RANGES]); hi = np.array([r[1] for r in p.synthetic.HJI_RANGES])
    x = lo + (hi - lo) * rng.uniform(-0.05, 1.05, (M, 7))         # ~30 % of the queries leave the grid in some dimension
    x[:500, 0] = knots[0][rng.integers(0, dims[0], 500)]          # exactly on knots and faces
    x[500:900, 3] = hi[3]; x[900:1200, 6] = lo[6]
    g.set_hji_lookup_order(0)
    V0, g0 = g.hji_lookup(x)
    g.set_hji_lookup_order(1)
    V1, g1 = g.hji_lookup(x)
    g.set_hji_lookup_order(-1)
    V2, g2 = g.hji_lookup(x)
    g.set_hji_lookup_order(2)                   # cell order + one TMA-staged tile of corners per block of cells (cp.async.bulk.tensor.5d)
    V3, g3 = g.hji_lookup(x)
    assert np.array_equal(V0, V1) and np.array_equal(g0, g1) and np.array_equal(V0, V2) and np.array_equal(g0, g2)
    assert np.array_equal(V0, V3) and np.array_equal(g0, g3)
    inside = np.isfinite(V0)
    assert 0.3 < inside.mean() < 0.9 and np.all(g0[~inside] == 0)
    cache = o.HjiCache(knots, V, gV)
    for j in rng.integers(0, M, 300):
        v, gr = cache.lookup(x[j])
        if np.isinf(v):
            assert np.isinf(V1[j])
        else:
            assert abs(v - V1[j]) < 1e-6 and np.max(np.abs(gr - g1[j])) < 1e-6
    for mode in (1, 2):
        g.set_hji_lookup_order(mode)
        Vs, gs = g.hji_lookup(x[:33])           # tiny sets work in cell order too
        assert np.array_equal(Vs, V0[:33]) and np.array_equal(gs, g0[:33])
    g.close()


@pytest.mark.gpu
def test_free_running_profile_records_every_admm_cta_and_changes_nothing(p):
    """pgn_set_profiling(3): the counting build of the ADMM kernel inside the ordinary free-running loop (pipeline parts and graphs on).  One
    record per ADMM CTA (start <= end on %globaltimer, its pipeline part, the QPs it solved) and seven time stamps per round and part; the
    closed loop itself must come out bit-identical to the unprofiled one."""
    import torch
    B, n, parts = 96, 6, 3
    trajs, tid, state, control, t0, other = batch(p, B, n_traj=4)

    def run(prof):
        g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
        g.set_pipeline_parts(parts)
        g.set_profiling(prof)
        g.set_state(state, control, other)
        d = torch.tensor(t0, dtype=torch.float64, device="cuda")
        g.simulate_device_async(d.data_ptr(), 0.01, n)
        q, u = g.get_state()
        st = g.stats()
        tr = g.admm_trace() if prof == 3 else None
        cyc = g.admm_cycles() if prof == 3 else None
        g.close()
        return q, u, st["iters"], st["status"], tr, cyc
    ref = run(0)
    got = run(3)
    for a, b in zip(ref[:4], got[:4]):
        assert np.array_equal(a, b, equal_nan=True)
    tr, cyc = got[4], got[5]
    stamp = (tr[:, 2] >> np.uint64(63)) == 1
    cta, st = tr[~stamp], tr[stamp]
    assert len(cta) > 0 and (cta[:, 0] <= cta[:, 1]).all()
    part = (cta[:, 2] & np.uint64(0xff)).astype(int)
    nqp = ((cta[:, 2] >> np.uint64(8)) & np.uint64(0xffffff)).astype(int)
    assert set(part) == set(range(parts))
    assert nqp.sum() >= B * n                                     # every vehicle solved n QPs (deferred solves may add continuation launches)
    assert len(st) % (7 * parts) == 0 and len(st) >= 7 * parts * n
    stage = ((st[:, 2] >> np.uint64(8)) & np.uint64(0xff)).astype(int)
    assert set(stage) == set(range(7))
    assert sum(cyc.values()) > 0
