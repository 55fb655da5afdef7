"""CPU tests pinning the oracle's OSQP restatement: optimality (independent KKT check), OSQP-specific behaviours
(Ruiz scaling equilibrates, rho per constraint type, 25-iteration check cadence, adaptive rho, warm start, infeasibility
certificates)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle_py as o


def random_qp(rng, n=20, m=30, n_eq=5):
    M = rng.normal(size=(n, n))
    P = M @ M.T * 0.1 + np.diag(rng.random(n) * (rng.random(n) < 0.5))
    P = sp.csc_matrix(np.triu(P))
    q = rng.normal(size=n)
    A = sp.random(m, n, density=0.3, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csc") + sp.csc_matrix(
        (np.ones(min(m, n)), (np.arange(min(m, n)), np.arange(min(m, n)))), shape=(m, n))
    A = sp.csc_matrix(A)
    x0 = rng.normal(size=n)
    Ax0 = A @ x0
    l = Ax0 - rng.random(m)
    u = Ax0 + rng.random(m)
    l[:n_eq] = u[:n_eq] = Ax0[:n_eq]
    l[n_eq:n_eq + 5] = -np.inf
    u[n_eq + 5:n_eq + 10] = np.inf
    return P, q, A, l, u


def kkt_violation(P, q, A, l, u, x, y):
    Pf = sp.csc_matrix(P)
    Pf = Pf + Pf.T - sp.diags(Pf.diagonal())
    stat = np.abs(Pf @ x + q + A.T @ y).max()
    Ax = A @ x
    prim = max(np.max(np.maximum(l - Ax, 0)), np.max(np.maximum(Ax - u, 0)))
    # complementary slackness: y_i > 0 only at the upper bound, y_i < 0 only at the lower bound
    comp = max(np.max(np.maximum(y, 0) * np.where(np.isfinite(u), u - Ax, 0)), np.max(np.maximum(-y, 0) * np.where(np.isfinite(l), Ax - l, 0)))
    return stat, prim, comp


def test_high_accuracy_solution_satisfies_kkt():
    rng = np.random.default_rng(0)
    for trial in range(5):
        P, q, A, l, u = random_qp(rng)
        s = o.Osqp(P, q, A, l, u, o.osqp_settings_default(eps_abs=1e-10, eps_rel=1e-10, max_iter=20000))
        x, y, info = s.solve()
        assert info["status"] == 1
        stat, prim, comp = kkt_violation(P, q, A, l, u, x, y)
        assert stat < 1e-7 and prim < 1e-7 and comp < 1e-6


def test_default_tolerance_solution_close_to_optimum_and_check_cadence():
    rng = np.random.default_rng(1)
    P, q, A, l, u = random_qp(rng)
    xs, _, _ = o.Osqp(P, q, A, l, u, o.osqp_settings_default(eps_abs=1e-10, eps_rel=1e-10, max_iter=20000)).solve()
    s = o.Osqp(P, q, A, l, u)
    x, y, info = s.solve()
    assert info["status"] == 1 and info["iter"] % 25 == 0 and info["iter"] >= 25
    assert np.max(np.abs(x - xs)) < 5e-2
    # warm-started re-solve of the same problem terminates at the first check
    x2, y2, info2 = s.solve()
    assert info2["iter"] == 25
    # cold start reproduces the first solve exactly (deterministic)
    s.cold_start()
    s2 = o.Osqp(P, q, A, l, u)
    xa, _, ia = s2.solve()
    assert np.array_equal(xa, x) and ia["iter"] == info["iter"]


def test_ruiz_scaling_equilibrates_kkt():
    rng = np.random.default_rng(2)
    P, q, A, l, u = random_qp(rng)
    A = sp.csc_matrix(sp.diags(10.0 ** rng.uniform(-2, 2, A.shape[0])) @ A @ sp.diags(10.0 ** rng.uniform(-2, 2, A.shape[1])))
    s = o.Osqp(P, q, A, l, u)
    D = np.zeros(s.n); E = np.zeros(s.m); c = o.C.c_double(0)
    o.lib().orc_osqp_get_scaling(s.h, D.ctypes.data_as(o.dp), E.ctypes.data_as(o.dp), o.C.byref(c))
    Pf = sp.csc_matrix(P); Pf = (Pf + Pf.T - sp.diags(Pf.diagonal())).toarray()
    Ps = c.value * D[:, None] * Pf * D[None, :]
    As = E[:, None] * A.toarray() * D[None, :]
    K = np.block([[Ps, As.T], [As, np.zeros((s.m, s.m))]])
    norms = np.abs(K).max(axis=0)
    # after 10 Ruiz passes the A-rows are equilibrated to O(1) (cost scaling perturbs the P-columns only)
    assert 0.5 < norms[s.n:].min() and norms[s.n:].max() < 2.0
    assert 0.3 < norms[:s.n].min() and norms[:s.n].max() < 3.0
    assert np.all(D >= 1e-4) and np.all(E >= 1e-4)


def test_primal_infeasible_detected():
    # x <= 0 and x >= 1
    P = sp.csc_matrix([[1.0]]); q = np.array([0.0])
    A = sp.csc_matrix([[1.0], [1.0]]); l = np.array([-np.inf, 1.0]); u = np.array([0.0, np.inf])
    x, y, info = o.Osqp(P, q, A, l, u).solve()
    assert info["status"] == -3 and np.all(np.isnan(x))


def test_dual_infeasible_detected():
    # min -x s.t. x >= 0  (unbounded)
    P = sp.csc_matrix((1, 1)); q = np.array([-1.0])
    A = sp.csc_matrix([[1.0]]); l = np.array([0.0]); u = np.array([np.inf])
    x, y, info = o.Osqp(P, q, A, l, u).solve()
    assert info["status"] == -4


def test_adaptive_rho_changes_iteration_count_and_refactors():
    rng = np.random.default_rng(3)
    P, q, A, l, u = random_qp(rng, n=30, m=50, n_eq=10)
    triggered = 0
    for scale in (1e-3, 1e-2, 1e2, 1e3):
        Ps, qs = P * scale, q * scale
        a = o.Osqp(Ps, qs, A, l, u, o.osqp_settings_default(adaptive_rho=1, adaptive_rho_interval=25, scaling=0))
        b = o.Osqp(Ps, qs, A, l, u, o.osqp_settings_default(adaptive_rho=0, scaling=0))
        xa, _, ia = a.solve()
        xb, _, ib = b.solve()
        assert ia["status"] == 1
        assert ia["n_factor"] == 1 + ia["rho_updates"] and ib["n_factor"] == 1 and ib["rho"] == 0.1
        if ia["rho_updates"] >= 1:
            triggered += 1
            assert ia["rho"] != 0.1
            assert ia["iter"] <= ib["iter"]
    assert triggered >= 2


def test_update_values_equals_fresh_setup_with_warm_iterates():
    rng = np.random.default_rng(4)
    P, q, A, l, u = random_qp(rng)
    s = o.Osqp(P, q, A, l, u)
    s.solve()
    A2 = sp.csc_matrix(A); A2.data = A2.data * (1 + 0.01 * rng.normal(size=A2.nnz))
    q2 = q + 0.01 * rng.normal(size=q.size)
    s.update(Ax=A2.data, q=q2)
    x1, y1, i1 = s.solve()
    assert i1["status"] == 1
    xs, _, _ = o.Osqp(P, q2, A2, l, u, o.osqp_settings_default(eps_abs=1e-10, eps_rel=1e-10, max_iter=20000)).solve()
    assert np.max(np.abs(x1 - xs)) < 5e-2
