"""Shared helpers for the test-suite (oracle side): fixture trajectories, analytic HJI grid."""
import os

import numpy as np

import oracle_py as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def world_fields(name):
    """TrajectoryTube(p::path) of the reference (src/ros_integration.jl:13-16) applied to a .world fixture."""
    w = np.load(os.path.join(GOLDEN, f"world_{name}.npz"))
    t = o.invcumtrapz(w["UxDes_mps"], w["s_m"])
    return dict(t=t, s=w["s_m"], V=w["UxDes_mps"], A=w["AxDes_mps2"], E=w["posE_m"], N=w["posN_m"], psi=w["psi_rad"],
                kappa=w["k_1pm"], theta=w["grade_rad"], phi=0 * w["grade_rad"], edge_L=w["edgeL_m"], edge_R=w["edgeR_m"])


def world_trajectory(name):
    return o.Trajectory(**world_fields(name))


def fd_jacobian(f, x, h=1e-6):
    x = np.asarray(x, dtype=float)
    f0 = np.asarray(f(x))
    J = np.zeros((f0.size, x.size))
    for j in range(x.size):
        e = np.zeros_like(x)
        e[j] = h * max(1.0, abs(x[j]))
        J[:, j] = (np.asarray(f(x + e)) - np.asarray(f(x - e))) / (2 * e[j])
    return J
