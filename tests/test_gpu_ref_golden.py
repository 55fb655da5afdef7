"""-m gpu: the CUDA engine (through the C ABI) against the reference-generated golden vectors of julia/make_golden.jl.
Skips, with the recipe in the reason, until tests/golden/ref_*.npz exist (they need Julia 1.0.x + the reference's environment to produce);
tests/test_ref_golden.py holds the CPU twin (oracle vs golden) and the always-on checks of the container / comparison machinery."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import ref_golden as rg  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rg.any_available(), reason=rg.RECIPE)]
FAR = np.array([[1e4, 1e4, 0.0, 5.0]])


@pytest.fixture(scope="module")
def p():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pigeon.jl_b200 as pkg
    pkg.load()
    return pkg


def gpu_step_arrays(g, t):
    g.compute_time_steps(t); g.compute_linearization_nodes()
    ts, dt, _ = g.time_steps()
    qs, us, ps = g.nodes()
    g.update_QP()
    d = g.qp_data()
    g.solve()
    x, _ = g.solution()
    st = g.stats()
    out = dict(ts=ts[0], dt=dt[0], qs=qs[0], us=us[0], ps=ps[0], x=x[0], iters=st["iters"][0], status=st["status"][0], control=g.get_next_control()[0])
    out.update({k: d[k][0] for k in ("A", "B0", "Bf", "c", "H", "G", "dmin", "dmax", "fxmax", "hji")})
    return out


def make(p, ctl, traj):
    cfg = rg.CONTROLLERS[ctl]
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if cfg["kind"] == 0 else p.BatchedDecoupledTrajectoryTrackingMPC
    return ctor(p.X1(), [traj], 1, N_short=cfg["N_short"], N_long=cfg["N_long"])


@pytest.mark.parametrize("ctl", ["C31", "X1CMPC", "X1DMPC"])
def test_gpu_matches_reference_dry_run(p, ctl):
    for group, pinned in (("dry25", True), ("dry", False)):
        G = rg.load(group)
        g = make(p, ctl, p.straight_trajectory(30.0, 5.0))
        g.set_state(np.array([[0.0, 0.0, 0.0, 5.0, 0.0, 0.0]]), np.zeros((1, 3)), FAR)
        rg.compare_step(gpu_step_arrays(g, 0.0), G, f"{group}/{ctl}", ctl, pinned, [])
        g.close()


@pytest.mark.parametrize("ctl", ["C31", "X1CMPC", "X1DMPC"])
def test_gpu_matches_reference_simulate(p, ctl):
    w = np.load(os.path.join(ROOT, "tests", "golden", "world_skidpadoval.npz"))
    traj = p.TrajectoryTube.from_path(w)
    for group, pinned in ((f"sim25_{ctl}", True), (f"sim_{ctl}", False)):
        if not rg.available(group):
            pytest.skip(rg.RECIPE)
        G = rg.load(group)
        g = make(p, ctl, traj)
        g.set_state(None, None, FAR)
        tag = group.split("_")[0]
        rep = []
        for k in range(200):
            pre = f"{tag}/{ctl}/step{k:03d}"
            g.set_state(G[f"{pre}/state"][None, :], G[f"{pre}/control"][None, :])
            rg.compare_step(gpu_step_arrays(g, float(G[f"{pre}/t"][0])), G, pre, ctl, pinned, rep)
        g.close()


def test_gpu_matches_reference_hji(p):
    G = rg.load("hji")
    g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), [p.straight_trajectory(30.0, 5.0)], 1)
    for name in ("placeholder", "analytic"):
        if name == "analytic":
            dims = tuple(int(d) for d in G["hji/analytic/dims"])
            kn = G["hji/analytic/knots"]; off = np.concatenate([[0], np.cumsum(dims)])
            g.set_HJI_cache(p.HJICache([kn[off[i]:off[i + 1]] for i in range(7)], np.transpose(G["hji/analytic/V"]), np.transpose(G["hji/analytic/gradV"])))
        X, Vr, Gr = G[f"hji/{name}/x"], G[f"hji/{name}/V"], G[f"hji/{name}/grad"]
        V, gr = g.hji_lookup(X)
        inf = np.isinf(Vr)
        assert np.array_equal(np.isinf(V), inf) and np.all(gr[inf] == 0)
        assert np.max(np.abs(V[~inf] - Vr[~inf])) <= 1e-6 and np.max(np.abs(gr[~inf] - Gr[~inf])) <= 1e-6
    g.close()
