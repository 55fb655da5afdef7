"""CPU tests pinning the oracle's primitives against independent computations (scipy, finite differences) and the
known answers derivable from the reference."""
import numpy as np
import pytest
import scipy.linalg

import oracle_py as o
from helpers import fd_jacobian, world_trajectory


def test_x1_known_constants(vp):
    d = dict(zip(o.VP_NAMES, vp))
    # /root/reference/src/vehicles.jl:1-59
    assert d["m"] == 1964
    assert d["a"] == pytest.approx(1.49783604887984, rel=1e-14)
    assert d["b"] == pytest.approx(1.37216395112016, rel=1e-14)
    assert d["h"] == pytest.approx(0.47, rel=1e-14)
    assert d["Fx_min"] == pytest.approx(-16793.7329957606, rel=1e-13)
    assert d["delta_max"] == pytest.approx(0.314159265358979, rel=1e-14)
    assert d["kappa_max"] == pytest.approx(0.113212437711814, rel=1e-13)


def test_time_steps_correction_step():
    # /root/reference/src/model_predictive_control.jl:17-30 ; SURVEY §8a A1 example t0=0.123 => first long step 0.177
    m = o.Mpc(o.MPC_COUPLED)
    m.compute_time_steps(0.123)
    ts, dt, prev = m.time_steps()
    assert np.allclose(prev, np.arange(1, 32))
    assert np.allclose(ts[:11], 0.123 + 0.01 * np.arange(11))
    assert dt[10] == pytest.approx(0.177, abs=1e-12)
    assert np.allclose(dt[11:], 0.2)
    assert np.allclose(ts[11:], 0.2 * np.ceil((0.123 + 0.1 + 0.01) / 0.2 - 1) + 0.2 * np.arange(1, 21))
    m.compute_time_steps(0.133)
    _, _, prev2 = m.time_steps()
    assert np.allclose(prev2, ts)
    m2 = o.Mpc(o.MPC_COUPLED, use_correction_step=False)
    m2.compute_time_steps(0.123)
    ts2, dt2, _ = m2.time_steps()
    assert dt2[10] == pytest.approx(0.2)


def test_adiff_wraps():
    # /root/reference/src/PigeonViz.jl:24-28
    assert o.lib().orc_adiff(0.1, 0.0) == pytest.approx(0.1)
    assert o.lib().orc_adiff(0.0, 0.1) == pytest.approx(-0.1)
    assert o.lib().orc_adiff(3.0, -3.0) == pytest.approx(6.0 - 2 * np.pi)
    assert o.lib().orc_adiff(np.pi, 0.0) == pytest.approx(np.pi)
    assert o.lib().orc_adiff(-np.pi + 1e-9, 0.0) == pytest.approx(-np.pi + 1e-9)


def test_fiala_saturation_and_linear_region(vp):
    Ca, mu, Fz = 150e3, 0.92, 9000.0
    # small slip: Fy ~ -Ca*tan(alpha)
    a = 1e-4
    assert o.lib().orc_fiala(a, Ca, mu, 0.0, Fz) == pytest.approx(-Ca * np.tan(a), rel=1e-3)
    # full slide: |Fy| = mu Fz
    assert o.lib().orc_fiala(0.5, Ca, mu, 0.0, Fz) == pytest.approx(-mu * Fz)
    assert o.lib().orc_fiala(-0.5, Ca, mu, 0.0, Fz) == pytest.approx(mu * Fz)
    # friction circle derating and |Fx| >= mu Fz => 0
    assert o.lib().orc_fiala(0.5, Ca, mu, 3000.0, Fz) == pytest.approx(-np.sqrt((mu * Fz) ** 2 - 3000.0 ** 2))
    assert o.lib().orc_fiala(0.5, Ca, mu, mu * Fz, Fz) == 0.0
    # the reference's inverse returns the slip *ratio* in the unsaturated branch (vehicle_dynamics.jl:56-62)
    Fy_max = mu * Fz
    r = 0.3
    Fy = -Fy_max * (1 - (1 - r) ** 3)
    assert o.lib().orc_invfiala(Fy, Ca, Fy_max) == pytest.approx(r, rel=1e-12)
    assert o.lib().orc_invfiala(2 * Fy_max, Ca, Fy_max) == pytest.approx(-3 * Fy_max / Ca)


@pytest.mark.parametrize("kind,nx", [(o.MODEL_BICYCLE, 6), (o.MODEL_TRACKING, 6), (o.MODEL_LATERAL, 4)])
def test_continuous_jacobian_matches_finite_differences(vp, kind, nx):
    rng = np.random.default_rng(0)
    for trial in range(5):
        if kind == o.MODEL_LATERAL:
            x = np.array([0.2, 0.1, 0.02, 0.3]) * rng.normal(size=4)
            up = np.array([0.03 * rng.normal(), 800 * rng.normal(), 6 + rng.random() * 4, 0.02 * rng.normal(), 0, 0])
        elif kind == o.MODEL_TRACKING:
            x = np.array([0.5 * rng.normal(), 6 + 4 * rng.random(), 0.2 * rng.normal(), 0.1 * rng.normal(), 0.02 * rng.normal(), 0.3 * rng.normal()])
            up = np.array([0.03 * rng.normal(), 800 * rng.normal(), 7.0, 0.02 * rng.normal(), 0, 0])
        else:
            x = np.array([3 * rng.normal(), 3 * rng.normal(), rng.normal(), 6 + 4 * rng.random(), 0.2 * rng.normal(), 0.1 * rng.normal()])
            up = np.array([0.03 * rng.normal(), 800 * rng.normal(), 0, 0, 0, 0])
        A, B, f = o.linearize_continuous(kind, vp, x, up)
        f0 = o.vehicle_model(kind, vp, x, up[:2], up[2:])
        assert np.allclose(f, f0, rtol=0, atol=1e-13)
        Afd = fd_jacobian(lambda z: o.vehicle_model(kind, vp, z, up[:2], up[2:]), x)
        Bfd = fd_jacobian(lambda w: o.vehicle_model(kind, vp, x, w[:2], w[2:]), up)
        if kind == o.MODEL_TRACKING:
            # apply_control_limits strips the dual of Ux (vehicle_dynamics.jl:295): identical here because Px_max/Ux is inactive
            pass
        assert np.allclose(A, Afd, rtol=2e-5, atol=2e-5 * max(1, np.abs(Afd).max()))
        assert np.allclose(B[:, :4], Bfd[:, :4], rtol=2e-5, atol=2e-5 * max(1, np.abs(Bfd).max()))


def test_ux_dual_is_stripped_in_power_limit(vp):
    # Fx above Px_max/Ux: the limit is active and d/dUx of (Px_max/Ux) must NOT appear (vehicle_dynamics.jl:295-297)
    x = np.array([0.0, 14.0, 0.0, 0.0, 0.0, 0.0])
    up = np.array([0.0, 5500.0, 14.0, 0.0, 0.0, 0.0])  # Px_max/Ux = 5357 < 5500
    A, B, f = o.linearize_continuous(o.MODEL_TRACKING, vp, x, up)
    m, Cd1 = vp[5], vp[11]
    assert A[1, 1] == pytest.approx(-Cd1 / m, rel=1e-9)       # only drag; no -Px_max/Ux^2/m term
    assert B[1, 1] == 0.0                                       # saturated input => zero gain
    Afd = fd_jacobian(lambda z: o.vehicle_model(o.MODEL_TRACKING, vp, z, up[:2], up[2:]), x)
    assert abs(Afd[1, 1] - A[1, 1]) > 1e-3                      # a true derivative would differ


def test_flow_jacobians_match_finite_differences(vp):
    rng = np.random.default_rng(1)
    x = np.array([0.3, 8.0, 0.1, 0.05, 0.01, 0.2])
    up0 = np.array([0.02, 500.0, 8.0, 0.01, 0, 0])
    upf = np.array([0.03, 300.0, 8.2, 0.015, 0, 0])
    for ramp, dt in [(False, 0.01), (True, 0.2), (True, 0.177)]:
        A, B0, Bf, c = o.linearize_flow(o.MODEL_TRACKING, vp, x, dt, up0, upf, ramp=ramp, nk=2)
        uf = upf if ramp else up0
        xp = o.flow(o.MODEL_TRACKING, vp, x, dt, up0, uf)
        Afd = fd_jacobian(lambda z: o.flow(o.MODEL_TRACKING, vp, z, dt, up0, uf), x)
        assert np.allclose(A, Afd, rtol=1e-6, atol=1e-7)
        if ramp:
            B0fd = fd_jacobian(lambda w: o.flow(o.MODEL_TRACKING, vp, x, dt, np.r_[w, up0[2:]], upf), up0[:2])
            Bffd = fd_jacobian(lambda w: o.flow(o.MODEL_TRACKING, vp, x, dt, up0, np.r_[w, upf[2:]]), upf[:2])
            assert np.allclose(B0, B0fd, rtol=1e-5, atol=1e-9)
            assert np.allclose(Bf, Bffd, rtol=1e-5, atol=1e-9)
            assert np.allclose(c, xp - A @ x - B0 @ up0[:2] - Bf @ upf[:2], atol=1e-12)
        else:
            Bfd = fd_jacobian(lambda w: o.flow(o.MODEL_TRACKING, vp, x, dt, np.r_[w, up0[2:]], np.r_[w, up0[2:]]), up0[:2])
            assert np.allclose(B0, Bfd, rtol=1e-5, atol=1e-9)
            assert np.all(Bf == 0)
            assert np.allclose(c, xp - A @ x - B0 @ up0[:2], atol=1e-12)


def test_rk4_flow_converges_to_fine_integration(vp):
    x = np.array([0.3, 8.0, 0.1, 0.05, 0.01, 0.2])
    up0 = np.array([0.02, 500.0, 8.0, 0.01, 0, 0])
    x10 = o.flow(o.MODEL_TRACKING, vp, x, 0.2, up0, nsub=10)
    x400 = o.flow(o.MODEL_TRACKING, vp, x, 0.2, up0, nsub=400)
    assert np.allclose(x10, x400, atol=5e-5)   # the yaw/sideslip modes are stiff at 8 m/s: RK4 with h = 0.02 is only ~1e-5 accurate


def test_expm_matches_scipy():
    rng = np.random.default_rng(2)
    for n in (2, 4, 12):
        for scale in (0.01, 1.0, 30.0):
            A = rng.normal(size=(n, n)) * scale / np.sqrt(n)
            E = o.expm(A)
            Es = scipy.linalg.expm(A)
            assert np.allclose(E, Es, rtol=1e-10, atol=1e-10 * np.abs(Es).max())


def test_exact_discretisation_matches_augmented_expm(vp):
    x = np.array([0.1, 0.05, 0.01, 0.2])
    up0 = np.array([0.02, 500.0, 8.0, 0.01, 0, 0])
    upf = np.array([0.03, 300.0, 8.3, 0.015, 0, 0])
    Ac, Bc, f = o.linearize_continuous(o.MODEL_LATERAL, vp, x, up0)
    c0 = f - Ac @ x - Bc @ up0
    for ramp, dt in [(False, 0.01), (True, 0.2), (True, 0.05)]:
        A, B0, Bf, c = o.linearize_exact(o.MODEL_LATERAL, vp, x, dt, up0, upf, ramp=ramp, nk=1)
        # independent: integrate the affine system [x; u; 1] with u ramping, via one big expm on [x, u, du, 1]
        nxx = 4
        M = np.zeros((nxx + 6 + 6 + 1, nxx + 6 + 6 + 1))
        M[:nxx, :nxx] = Ac
        M[:nxx, nxx:nxx + 6] = Bc
        M[:nxx, -1] = c0
        M[nxx:nxx + 6, nxx + 6:nxx + 12] = np.eye(6)   # u' = du
        E = scipy.linalg.expm(M * dt)
        du = (upf - up0) / dt if ramp else np.zeros(6)
        z0 = np.r_[x, up0, du, 1.0]
        xp = (E @ z0)[:nxx]
        assert np.allclose(A, scipy.linalg.expm(Ac * dt), rtol=1e-11, atol=1e-13)
        pred = A @ x + B0[:, 0] * up0[0] + (Bf[:, 0] * upf[0] if ramp else 0) + c
        assert np.allclose(pred, xp, rtol=1e-10, atol=1e-12)
        # sensitivity to the kept control: finite differences of the exact flow
        def flow_u0(d0):
            u0 = up0.copy(); u0[0] = d0
            duu = (upf - u0) / dt if ramp else np.zeros(6)
            return (E @ np.r_[x, u0, duu, 1.0])[:nxx]
        J0 = fd_jacobian(lambda w: flow_u0(w[0]), np.array([up0[0]]))
        assert np.allclose(B0[:, 0], J0[:, 0], rtol=1e-6, atol=1e-9)


def test_stable_limits_geometry(vp):
    d = dict(zip(o.VP_NAMES, vp))
    S = o.stable_limits(vp, 8.0, 0.0, 300.0)
    # vertices C, D lie on the r_max line; E, F on the r_min line (vehicle_dynamics.jl:244-261)
    assert S["delta_max"] > 0 > S["delta_min"]
    assert S["delta_max"] == pytest.approx(-S["delta_min"], rel=1e-12)
    assert S["G"][0] == S["G"][1] > 0
    rC = d["mu"] * d["G"] / 8.0
    Fzr = (d["m"] * d["G"] * d["a"] + d["h"] * 300.0) / d["L"]
    tan_ar = 3 * np.sqrt((d["mu"] * Fzr) ** 2 - 300.0 ** 2) / d["Car"]
    UyC = -8.0 * tan_ar + d["b"] * rC
    assert S["H"][2] @ np.array([UyC, rC]) == pytest.approx(S["G"][2], rel=1e-12)
    assert np.allclose(S["H"][0], [1 / 8.0, -d["b"] / 8.0])


def test_steady_state_straight_line_is_drag_equilibrium(vp):
    d = dict(zip(o.VP_NAMES, vp))
    est = o.steady_state(vp, 5.0, 0.0, 0.0)
    assert est["delta"] == pytest.approx(0.0, abs=1e-15)
    assert est["Fxf"] + est["Fxr"] == pytest.approx(d["Cd0"] + d["Cd1"] * 5.0)     # 366.5 N
    assert est["A"] == pytest.approx(0.0, abs=1e-12)
    est = o.steady_state(vp, 6.0, 0.0, 0.05)
    assert est["r"] == pytest.approx(0.3)
    assert est["delta"] > 0


def test_trajectory_lookups_on_fixture():
    tr = world_trajectory("skidpadoval")
    f = tr.fields
    # at a knot: searchsortedfirst-1 picks the previous interval, value is continuous
    k = 400
    a = tr.at_s(f["s"][k])
    assert a["E"] == pytest.approx(f["E"][k], abs=1e-9) and a["kappa"] == pytest.approx(f["kappa"][k], abs=1e-12)
    assert a["V"] == pytest.approx(6.0)
    b = tr.at_time(f["t"][k] + 1e-3)
    assert b["s"] == pytest.approx(f["s"][k] + 6e-3, rel=1e-9)
    # Line() extrapolation beyond the end (trajectories.jl:32-35)
    e = tr.at_s(f["s"][-1] + 1.0)
    slope = (f["E"][-1] - f["E"][-2]) / (f["s"][-1] - f["s"][-2])
    assert e["E"] == pytest.approx(f["E"][-1] + slope * 1.0, rel=1e-9)
    # path_coordinates: a point offset to the left of segment k by 0.5 m
    psi = f["psi"][k]
    mid = 0.5 * np.array([f["E"][k] + f["E"][k + 1], f["N"][k] + f["N"][k + 1]])
    v = np.array([f["E"][k + 1] - f["E"][k], f["N"][k + 1] - f["N"][k]])
    nrm = np.array([-v[1], v[0]]) / np.linalg.norm(v)
    s, e_lat, t = tr.path_coordinates(*(mid + 0.5 * nrm))
    assert e_lat == pytest.approx(0.5, abs=1e-6)
    assert s == pytest.approx(0.5 * (f["s"][k] + f["s"][k + 1]), abs=2e-3)
    assert t == pytest.approx(s / 6.0, abs=1e-3)


def test_path_coordinates_first_minimum_wins():
    tr = o.Trajectory(t=[0, 1, 2], s=[0, 1, 2], V=[1, 1, 1], A=[0, 0, 0], E=[0, 1, 2], N=[0, 0, 0], psi=[0, 0, 0], kappa=[0, 0, 0])
    s, e, t = tr.path_coordinates(1.0, 0.3)        # equidistant from both segments -> the first one
    assert s == pytest.approx(1.0) and e == pytest.approx(0.3) and t == pytest.approx(1.0)
    s, e, t = tr.path_coordinates(1.5, -0.2)
    assert s == pytest.approx(1.5) and e == pytest.approx(-0.2)
