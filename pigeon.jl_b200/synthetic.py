"""Synthetic workloads for tests and bench.py: clothoid-ish reference trajectories, perturbed initial
states, and an analytic 7-D HJI grid with a closed-form gradient.  numpy only; no GPU, no oracle."""
import numpy as np

SEED = 0x5049474E
MU_G = 0.92 * 9.80665


def synthetic_trajectories(seed=SEED, n_traj=64, n_nodes=1000, ds=0.25):
    """Returns dict of (n_traj, n_nodes) float64 arrays with the 12 TrajectoryTube fields
    (t,s,V,A,E,N,psi,kappa,theta,phi,edge_L,edge_R); piecewise-linear curvature |kappa| <= 0.07, V in [4,12],
    |A| <= 1.5, V^2 |kappa| <= 0.5 mu g."""
    rng = np.random.default_rng(seed)
    s = np.arange(n_nodes) * ds
    out = {k: np.zeros((n_traj, n_nodes)) for k in ("t", "s", "V", "A", "E", "N", "psi", "kappa", "theta", "phi", "edge_L", "edge_R")}
    for j in range(n_traj):
        # curvature: piecewise-linear (clothoid segments) between random knots every 10..40 m
        knots_s = [0.0]
        while knots_s[-1] < s[-1]:
            knots_s.append(knots_s[-1] + rng.uniform(10.0, 40.0))
        knots_k = rng.uniform(-0.07, 0.07, size=len(knots_s)) * (rng.random(len(knots_s)) < 0.7)
        kappa = np.interp(s, knots_s, knots_k)
        # speed: smooth profile from random knots, acceleration-limited, lateral-acceleration-limited
        knots_v = rng.uniform(4.0, 12.0, size=len(knots_s))
        V = np.interp(s, knots_s, knots_v)
        V = np.minimum(V, np.sqrt(0.5 * MU_G / np.maximum(np.abs(kappa), 1e-9)))
        V = np.clip(V, 4.0, 12.0)
        for _ in range(2):   # enforce |A| <= 1.5 via forward/backward passes on V^2
            for i in range(1, n_nodes):
                V[i] = min(V[i], np.sqrt(V[i - 1] ** 2 + 2 * 1.5 * ds))
            for i in range(n_nodes - 2, -1, -1):
                V[i] = min(V[i], np.sqrt(V[i + 1] ** 2 + 2 * 1.5 * ds))
        A = np.zeros(n_nodes)
        A[:-1] = (V[1:] ** 2 - V[:-1] ** 2) / (2 * ds)
        A[-1] = A[-2]
        psi0 = rng.uniform(-np.pi, np.pi)
        psi = psi0 + np.concatenate([[0.0], np.cumsum(0.5 * (kappa[1:] + kappa[:-1]) * ds)])
        # psi measured from North (vehicle_dynamics.jl:127-128): dE/ds = -sin(psi), dN/ds = cos(psi)
        E = rng.uniform(-100, 100) + np.concatenate([[0.0], np.cumsum(-np.sin(0.5 * (psi[1:] + psi[:-1])) * ds)])
        N = rng.uniform(-100, 100) + np.concatenate([[0.0], np.cumsum(np.cos(0.5 * (psi[1:] + psi[:-1])) * ds)])
        t = np.concatenate([[0.0], np.cumsum(2 * ds / (V[1:] + V[:-1]))])   # invcumtrapz (math.jl:2)
        out["t"][j], out["s"][j], out["V"][j], out["A"][j] = t, s, V, A
        out["E"][j], out["N"][j], out["psi"][j], out["kappa"][j] = E, N, psi, kappa
        out["edge_L"][j], out["edge_R"][j] = 4.0, -4.0
    return out


def synthetic_batch(trajs, B, seed=SEED + 17, s_frac=0.5, L=2.87, Cd0=241.0, Cd1=25.1):
    """Perturbed initial states for B vehicles; vehicle i follows trajectory i mod n_traj.
    Returns traj_id (B,) int32, state (B,6) [E,N,psi,Ux,Uy,r], control (B,3) [delta,Fxf,Fxr], t0 (B,).
    Vectorised over the vehicles of one trajectory (a million-vehicle batch in seconds); draws the random numbers in the same order as the
    per-vehicle loop it replaced, so every seed gives bit-identical batches (tests/test_host_cpu.py)."""
    rng = np.random.default_rng(seed)
    n_traj, n_nodes = trajs["s"].shape
    tid = (np.arange(B) % n_traj).astype(np.int32)
    s_end = trajs["s"][tid, -1]
    s0 = rng.uniform(0.0, s_frac, size=B) * s_end
    e = rng.normal(0, 0.3, B)
    dpsi = rng.normal(0, 0.05, B)
    z = rng.standard_normal((B, 3))          # per vehicle: Ux, Uy, r perturbations, in this order
    f = {k: np.zeros(B) for k in ("psi", "kappa", "V", "E", "N", "t")}
    for j in range(min(n_traj, B)):
        sel = np.arange(j, B, n_traj)
        for k in f:
            f[k][sel] = np.interp(s0[sel], trajs["s"][j], trajs[k][j])
    psi, kap, V = f["psi"], f["kappa"], f["V"]
    state = np.zeros((B, 6))
    control = np.zeros((B, 3))
    state[:, 0] = f["E"] + e * (-np.cos(psi))      # left normal of heading-from-North: (-cos psi, -sin psi)
    state[:, 1] = f["N"] + e * (-np.sin(psi))
    state[:, 2] = psi + dpsi
    state[:, 3] = np.maximum(V + (0.0 + 0.5 * z[:, 0]), 1.5)
    state[:, 4] = 0.0 + 0.1 * z[:, 1]
    state[:, 5] = kap * V + (0.0 + 0.02 * z[:, 2])
    control[:, 0] = np.arctan(L * kap)
    control[:, 2] = Cd0 + Cd1 * state[:, 3]        # drag equilibrium on the rear (driven) axle
    return tid, state, control, f["t"].copy()


def _synthetic_batch_loop(trajs, B, seed=SEED + 17, s_frac=0.5, L=2.87, Cd0=241.0, Cd1=25.1):
    """The original per-vehicle formulation of synthetic_batch (kept as the reference of the vectorised one)."""
    rng = np.random.default_rng(seed)
    n_traj, n_nodes = trajs["s"].shape
    tid = (np.arange(B) % n_traj).astype(np.int32)
    s_end = trajs["s"][tid, -1]
    s0 = rng.uniform(0.0, s_frac, size=B) * s_end
    e = rng.normal(0, 0.3, B)
    dpsi = rng.normal(0, 0.05, B)
    state = np.zeros((B, 6))
    control = np.zeros((B, 3))
    t0 = np.zeros(B)
    for i in range(B):
        j = tid[i]
        f = lambda k: np.interp(s0[i], trajs["s"][j], trajs[k][j])
        psi, kap, V = f("psi"), f("kappa"), f("V")
        state[i, 0] = f("E") + e[i] * (-np.cos(psi))
        state[i, 1] = f("N") + e[i] * (-np.sin(psi))
        state[i, 2] = psi + dpsi[i]
        state[i, 3] = max(V + rng.normal(0, 0.5), 1.5)
        state[i, 4] = rng.normal(0, 0.1)
        state[i, 5] = kap * V + rng.normal(0, 0.02)
        control[i, 0] = np.arctan(L * kap)
        control[i, 2] = Cd0 + Cd1 * state[i, 3]
        t0[i] = f("t")
    return tid, state, control, t0


HJI_DIMS = (13, 13, 9, 9, 9, 9, 9)
HJI_RANGES = ((-15.0, 15.0), (-15.0, 15.0), (-np.pi, np.pi), (1.0, 15.0), (-2.0, 2.0), (1.0, 15.0), (-1.0, 1.0))


def analytic_hji_value(x):
    """x: (...,7) = (dE, dN, dpsi, Ux, Uy, V, r)  ->  V(x), gradV(x) (...,7)  (float64, exact)."""
    dE, dN, dpsi, Ux, Uy, V, r = [x[..., i] for i in range(7)]
    R = np.sqrt((dE / 4) ** 2 + (dN / 2) ** 2 + 0.01)
    val = R - 1 + 0.05 * (Ux - V) * np.cos(dpsi) + 0.02 * Uy * r
    g = np.stack([dE / 16 / R, dN / 4 / R, -0.05 * (Ux - V) * np.sin(dpsi), 0.05 * np.cos(dpsi), 0.02 * r, -0.05 * np.cos(dpsi), 0.02 * Uy], axis=-1)
    return val, g


def analytic_hji_grid(dims=HJI_DIMS, ranges=HJI_RANGES):
    """Returns knots (list of 7 float32 arrays), V float32 array of shape dims, gradV float32 of shape (7,)+dims."""
    knots = [np.linspace(lo, hi, n).astype(np.float32) for (lo, hi), n in zip(ranges, dims)]
    mesh = np.meshgrid(*[k.astype(np.float64) for k in knots], indexing="ij")
    x = np.stack(mesh, axis=-1)
    val, g = analytic_hji_value(x)
    return knots, val.astype(np.float32), np.moveaxis(g, -1, 0).astype(np.float32)
