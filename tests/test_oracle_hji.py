"""CPU tests pinning the oracle's HJI lookup / constraint (HJI_computation.jl) against scipy's RegularGridInterpolator
and the reference's placeholder-cache known answers."""
import importlib.util
import os

import numpy as np
import pytest
from scipy.interpolate import RegularGridInterpolator

import oracle_py as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("synthetic", os.path.join(ROOT, "pigeon.jl_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synthetic)

DIMS = (5, 4, 5, 4, 3, 4, 3)


@pytest.fixture(scope="module")
def grid():
    knots, V, gV = synthetic.analytic_hji_grid(DIMS)
    return knots, V, gV, o.HjiCache(knots, V, gV)


def test_lookup_matches_scipy_regular_grid(grid):
    knots, V, gV, cache = grid
    rng = np.random.default_rng(0)
    lo = np.array([k[0] for k in knots], dtype=np.float64)
    hi = np.array([k[-1] for k in knots], dtype=np.float64)
    x = lo + (hi - lo) * rng.random((500, 7))
    x[:7] = np.where(np.eye(7, dtype=bool), hi, x[:7])     # exactly on the upper faces
    x[7:14] = np.where(np.eye(7, dtype=bool), lo, x[7:14])
    k64 = [k.astype(np.float64) for k in knots]
    Vi = RegularGridInterpolator(k64, V.astype(np.float64))(x)
    Vo, go = cache.lookup(x)
    assert np.allclose(Vo, Vi, rtol=0, atol=1e-12)
    for c in range(7):
        gi = RegularGridInterpolator(k64, gV[c].astype(np.float64))(x)
        assert np.allclose(go[:, c], gi, rtol=0, atol=1e-12)
    # against the analytic function the multilinear interpolant is only O(h^2)-accurate, sanity only
    Va, _ = synthetic.analytic_hji_value(x)
    assert np.max(np.abs(Vo - Va)) < 5.0   # coarse test grid


def test_lookup_at_nodes_returns_table_values(grid):
    knots, V, gV, cache = grid
    idx = (2, 1, 3, 0, 2, 3, 1)
    x = np.array([knots[d][idx[d]] for d in range(7)], dtype=np.float64)
    Vo, go = cache.lookup(x)
    assert Vo[0] == np.float64(V[idx])
    assert np.array_equal(go[0], gV[(slice(None),) + idx].astype(np.float64))


def test_out_of_grid_returns_inf_and_zero(grid):
    knots, V, gV, cache = grid
    x = np.array([0.0, 0.0, 0.0, 5.0, 0.0, 5.0, 0.0])
    for d in range(7):
        y = x.copy()
        y[d] = float(knots[d][-1]) + 1e-6
        Vo, go = cache.lookup(y)
        assert np.isinf(Vo[0]) and Vo[0] > 0 and np.all(go == 0)


def test_placeholder_cache_known_answers(vp):
    # HJI_computation.jl:32-37,66-72: V = 0, gradV = 0 inside +-1000; (Inf, 0) outside; out of grid => M=(0,0), b=1 (:163-164)
    cache = o.HjiCache()
    Vo, go = cache.lookup(np.array([[1.0, -2.0, 0.3, 5.0, 0.1, 4.0, 0.0], [1001.0, 0, 0, 5, 0, 4, 0]]))
    assert Vo[0] == 0 and np.all(go[0] == 0) and np.isinf(Vo[1])
    M, b = o.reachability_constraint(vp, cache, [1001.0, 0, 0, 5, 0, 4, 0], 0.05, [0.0, 100.0])
    assert np.all(M == 0) and b == 1.0
    M, b = o.reachability_constraint(vp, cache, [1.0, 0, 0, 5, 0, 4, 0], 0.05, [0.0, 100.0])
    assert np.all(M == 0) and b == 0.0      # active with a zero gradient


def test_relative_state_convention():
    # psi measured from North: a car straight ahead (North of us when psi=0) has dE_rel = +d in the body x axis
    x = o.hji_relative_state([0, 0, 0.0, 5, 0.1, 0.2], [0.0, 10.0, 0.3, 4.0])
    assert np.allclose(x, [10.0, 0.0, 0.3, 5, 0.1, 4.0, 0.2])
    x = o.hji_relative_state([1, 2, np.pi / 2, 5, 0, 0], [1.0 - 10.0, 2.0, np.pi / 2, 4.0])   # heading West, other car 10 m West
    assert np.allclose(x[:3], [10.0, 0.0, 0.0], atol=1e-12)


def test_reachability_constraint_matches_finite_difference(vp, grid):
    knots, V, gV, cache = grid
    x7 = np.array([3.0, 0.5, 0.2, 8.0, 0.1, 6.0, 0.05])
    uR = np.array([0.02, 400.0])
    Vv, g = cache.lookup(x7)
    EPS = 10.0     # coarse test grid: force the constraint active
    uH = o.optimal_disturbance(vp, x7, g[0])
    M, b = o.reachability_constraint(vp, cache, x7, EPS, uR)

    def H(u):
        bd = o.vehicle_model(o.MODEL_BICYCLE, vp, [x7[0], x7[1], x7[2], x7[3], x7[4], x7[6]], u, [0, 0, 0, 0])
        f = np.array([x7[5] * np.cos(x7[2]) - x7[3] + x7[1] * x7[6], x7[5] * np.sin(x7[2]) - x7[4] - x7[0] * x7[6], uH[0] - x7[6],
                      bd[3], bd[4], uH[1], bd[5]])
        return g[0] @ f
    Mfd = np.array([(H(uR + [1e-6, 0]) - H(uR - [1e-6, 0])) / 2e-6, (H(uR + [0, 1e-2]) - H(uR - [0, 1e-2])) / 2e-2])
    assert np.allclose(M, Mfd, rtol=1e-5, atol=1e-9)
    assert b == pytest.approx(H(uR) - M @ uR, rel=1e-12)


def test_optimal_disturbance_branches(vp):
    x7 = np.array([3.0, 0.5, 0.2, 8.0, 0.1, 6.0, 0.05])
    g = np.zeros(7)
    assert np.all(o.optimal_disturbance(vp, x7, g) == 0)                     # lam_norm < 1e-3
    g[5] = -1.0                                                               # wants max acceleration (dMode=:min => sgn=-1)
    uH = o.optimal_disturbance(vp, x7, g)
    assert uH[1] == pytest.approx(min(5600 / 1964, 75e3 / 1964 / 6.0))
    # desAy = sgn*0.0*... = -0.0 and copysign(maxAy, -0.0) = -maxAy: restated literally (HJI_computation.jl:113-117)
    assert uH[0] == pytest.approx(-min(vp[21] * 36.0, np.sqrt((0.9 * 0.92 * 9.80665) ** 2 - uH[1] ** 2)) / 6.0)
    g[5] = 1.0                                                                # wants max braking; desAy = 0 => (0, maxAx)  [reference quirk]
    uH = o.optimal_disturbance(vp, x7, g)
    assert uH[1] == pytest.approx(min(5600 / 1964, 75e3 / 1964 / 6.0))
    x0 = x7.copy(); x0[5] = 0.0
    assert np.all(o.optimal_disturbance(vp, x0, g) == 0)                     # documented deviation: V_other = 0


def _optimal_control_numpy(vp, x7, g, N=50):
    """optimal_control (HJI_computation.jl:133-158, uMode = :max) written out with numpy over the whole Fx grid; the tire forces come
    from the oracle's lateral_tire_forces(q, u) (pinned in test_oracle_primitives.py)."""
    P = dict(zip(["L", "a", "b", "h", "G", "m", "Izz", "mu", "Caf", "Car", "Cd0", "Cd1", "Cd2", "fwd", "rwd", "fwb", "rwb", "Fx_max", "Fx_min", "Px_max",
                  "delta_max", "kappa_max", "corr"], vp))
    A = g[3] / P["m"]; B = g[4] / P["m"] + P["a"] * g[6] / P["Izz"]; Cc = g[4] / P["m"] - P["b"] * g[6] / P["Izz"]
    d_opt = P["delta_max"] if B >= 0 else -P["delta_max"]
    frac = np.arange(N) / (N - 1)
    Fx = frac * P["Fx_max"] + (1 - frac) * P["Fx_min"]
    q = np.array([0, 0, 0, x7[3], x7[4], x7[6]], float)
    val = np.empty(N)
    for n in range(N):
        u3 = [d_opt, Fx[n] * (P["fwd"] if Fx[n] > 0 else P["fwb"]), Fx[n] * (P["rwd"] if Fx[n] > 0 else P["rwb"])]
        Fyf, Fyr = o.lateral_tire_forces(vp, q, u3)
        val[n] = A * Fx[n] + B * Fyf + Cc * Fyr
    return np.array([d_opt, Fx[int(np.argmax(val))]]), val     # argmax = first maximum, as the strict > of the reference loop


def test_optimal_control_hammer_policy():
    vp = o.x1()
    rng = np.random.default_rng(3)
    for k in range(200):
        x7 = np.array([rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-3, 3), rng.uniform(1, 15), rng.uniform(-1.5, 1.5), rng.uniform(1, 12), rng.uniform(-0.8, 0.8)])
        g = rng.normal(0, 1, 7) * np.array([1, 1, 1, 0.5, 1, 1, 2.0])
        u = o.optimal_control(vp, x7, g)
        ref, val = _optimal_control_numpy(vp, x7, g)
        assert abs(u[0]) == vp[20] and u[0] == ref[0]                      # steering at the limit, sign of B
        assert vp[18] <= u[1] <= vp[17]
        assert u[1] == ref[1], (k, u, ref)
    # known answers: zero gradient => B = 0 >= 0 => +delta_max, all grid values tie at 0 => the first one (Fx_min) wins
    u = o.optimal_control(vp, np.array([0, 0, 0, 8.0, 0, 5.0, 0]), np.zeros(7))
    assert u[0] == vp[20] and u[1] == vp[18]
    # only dV/dUx > 0 => maximise A * Fx => Fx_max; < 0 => Fx_min
    assert o.optimal_control(vp, np.array([0, 0, 0, 8.0, 0, 5.0, 0]), np.array([0, 0, 0, 1.0, 0, 0, 0]))[1] == vp[17]
    assert o.optimal_control(vp, np.array([0, 0, 0, 8.0, 0, 5.0, 0]), np.array([0, 0, 0, -1.0, 0, 0, 0]))[1] == vp[18]


def test_mpc_hji_policy_override(grid):
    """use_HJI_policy (ros_integration.jl:115-118): V <= eps => get_next_control returns BicycleControl(LP, optimal_control(...))."""
    knots, V, gV = synthetic.analytic_hji_grid((13, 13, 7, 7, 5, 7, 5))
    cache = o.HjiCache(knots, V, gV)
    trajs = synthetic.synthetic_trajectories(n_traj=1, n_nodes=300)
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, 2)
    tr = o.Trajectory(**{k: trajs[k][0] for k in o.TRAJ_FIELDS})
    outs = []
    for i, other in enumerate([np.array([state[0, 0] + 1.0, state[0, 1] + 0.5, state[0, 2], 6.0]), np.array([1e4, 1e4, 0.0, 5.0])]):
        res = []
        for pol in (False, True):
            m = o.Mpc(o.MPC_COUPLED)
            m.set_trajectory(tr); m.set_hji(cache); m.set_hji_policy(pol)
            m.set_state(state[0], control[0], other4=other)
            res.append((m.step(t0[0]), m.hji_values()))
        outs.append(res)
    (u_qp, (V0, g0)), (u_pol, _) = outs[0]
    assert V0 <= 0.05                                    # other car 1 m away: inside the unsafe set
    vp = o.x1()
    x7 = o.hji_relative_state(state[0], np.array([state[0, 0] + 1.0, state[0, 1] + 0.5, state[0, 2], 6.0]))
    d, Fx = o.optimal_control(vp, x7, g0)
    exp = np.array([d, Fx * (vp[13] if Fx > 0 else vp[15]), Fx * (vp[14] if Fx > 0 else vp[16])])
    assert np.array_equal(u_pol, exp) and not np.allclose(u_pol, u_qp)
    (u_qp_far, (Vf, _)), (u_pol_far, _) = outs[1]
    assert np.isinf(Vf) and np.array_equal(u_qp_far, u_pol_far)        # out of grid: policy inactive
