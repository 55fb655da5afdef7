"""Shared by tests/test_ref_golden.py (CPU) and tests/test_gpu_ref_golden.py (GPU): loads the reference-generated golden vectors
(tests/golden/ref_*.npz, produced by julia/make_golden.jl + tests/golden/import_ref_golden.py) and compares one MPC step of an
implementation (the CPU oracle or the CUDA engine, behind a tiny adapter) with the reference's own numbers.

Tolerances (BASELINE.json north_star): time steps 1e-12; linearisation nodes 1e-7; QP data 1e-7 relative to the block's largest entry
(the reference integrates the flow with its own ODE scheme: this is where a different sub-step count of `propagate` would show up); QP
solution / controls 1e-4 of the actuator range at eps 1e-3, with IDENTICAL OSQP iteration counts and statuses in the runs with the
adaptive-rho interval pinned to 25 (`dry25`, `sim25_*`); with OSQP's own wall-clock dependent interval (`dry`, `sim_*`) the counts are
reported, not asserted."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RECIPE = ("parity UNPINNED: tests/golden/ref_*.npz are absent.  They can only be produced where Julia 1.0.x and the reference's environment are installed: "
          "`julia --project=<Pigeon.jl>/env julia/make_golden.jl <Pigeon.jl> tests/golden/ref_golden.bin && python tests/golden/import_ref_golden.py tests/golden/ref_golden.bin`")
U_RANGE = np.array([0.3141592653589793, 16793.73299576057, 16793.73299576057])
CONTROLLERS = {"C31": dict(kind=0, N_short=10, N_long=20), "X1CMPC": dict(kind=0, N_short=5, N_long=10), "X1DMPC": dict(kind=1, N_short=10, N_long=20)}


def available(group):
    return os.path.exists(os.path.join(HERE, f"ref_{group}.npz"))


def any_available():
    return bool(glob.glob(os.path.join(HERE, "ref_*.npz")))


def load(group):
    return dict(np.load(os.path.join(HERE, f"ref_{group}.npz")))


def rel_close(a, b, tol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    return float(np.max(np.abs(a - b))) / scale <= tol if a.size else True


def compare_step(impl, G, pre, ctl, pinned, report):
    """impl: dict of arrays of ONE vehicle's step from the implementation under test:
         ts, dt, qs [N][nx], us [N][2], ps [N][4], A [T][nx][nx], B0, Bf [T][nx][nu], c [T][nx], H [T][4][2], G [T][4], dmin, dmax, fxmax [T], hji [3],
         x [n], iters, status, control [3]
       G: golden dict; pre: name prefix of the step ("dry25/C31", "sim25/C31/step007")."""
    cfg = CONTROLLERS[ctl]
    Ns, N = cfg["N_short"], 1 + cfg["N_short"] + cfg["N_long"]
    coupled = cfg["kind"] == 0
    g = lambda k: G[f"{pre}/{k}"]
    assert np.allclose(impl["ts"], g("ts"), rtol=0, atol=1e-12) and np.allclose(impl["dt"], g("dt"), rtol=0, atol=1e-12), pre
    assert np.allclose(impl["qs"], g("qs"), rtol=1e-7, atol=1e-7), pre
    assert np.allclose(impl["us"], g("us"), rtol=1e-7, atol=1e-5), pre
    assert np.allclose(impl["ps"], g("ps"), rtol=1e-7, atol=1e-7), pre
    tolq = 1e-7
    assert rel_close(impl["A"], g("A"), tolq), pre
    assert rel_close(impl["B0"][:Ns], g("B"), tolq) and rel_close(impl["B0"][Ns:], g("B0"), tolq) and rel_close(impl["Bf"][Ns:], g("Bf"), tolq), pre
    assert rel_close(impl["c"], g("c"), tolq) and rel_close(impl["H"], g("H"), tolq) and rel_close(impl["G"], g("G"), tolq), pre
    assert rel_close(impl["dmin"], np.ravel(g("δ_min")), tolq) and rel_close(impl["dmax"], np.ravel(g("δ_max")), tolq), pre
    if coupled:
        assert rel_close(impl["fxmax"], np.ravel(g("Fx_max")), tolq), pre
        M, b = np.ravel(g("M_HJI")), np.ravel(g("b_HJI"))
        assert rel_close(impl["hji"][:2], M, tolq) and abs(impl["hji"][2] - b[0]) <= tolq * max(1.0, abs(b[0])), pre
    it, st, ru, interval = [int(v) for v in g("osqp")]
    report.append((pre, int(impl["iters"]), it, int(impl["status"]), st, interval))
    if pinned:
        assert (int(impl["iters"]), int(impl["status"])) == (it, st), (pre, impl["iters"], it, impl["status"], st)
    nx = 6 if coupled else 4
    xq = np.ravel(g("x_q"))
    assert np.max(np.abs(impl["x"][:nx * N] - xq)) <= (1e-4 if pinned else 5e-3) * max(1.0, np.max(np.abs(xq))), pre
    assert np.max(np.abs(impl["control"] - np.ravel(g("next_control"))) / U_RANGE) <= (1e-4 if pinned else 5e-3), pre
