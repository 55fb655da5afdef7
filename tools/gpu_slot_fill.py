#!/usr/bin/env python
"""GPU box: how full are the 296 ADMM slots in the free-running closed loop?  Sum of the ADMM CTA lifetimes (profiling build, clock64 of thread 0)
over one pgn_simulate_device call, against slots x wall time of the call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
import torch
import bench
import pigeon.jl_b200 as p
K = int(os.environ.get("K", "40"))
trajs, tid, state, control, t0, other = bench.make_workload(1, 1024, 0)
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
m.set_state(state, control, other)
m.simulate_device(t0, 0.01, 35)
for prof in (0, 3):
    m.set_profiling(prof)
    m.simulate_device(t0 + 0.35, 0.01, K)          # graphs of this mode captured and instantiated outside the timed call
    if prof: m.admm_cycles(reset=True)
    torch.cuda.synchronize()
    t = time.perf_counter(); m.simulate_device(t0 + 0.35 + 0.01 * K, 0.01, K); torch.cuda.synchronize(); t = time.perf_counter() - t
    print("profiling %d: %.3f ms / step (%.0f steps/s)" % (prof, t / K * 1e3, 1024 * K / t))
    if prof:
        out = np.zeros(512); m._lib.pgn_get_admm_cycles(m._h, out.ctypes.data_as(ctypes.c_void_p), 1)
        tot = out[:8].sum()
        print("cycles per QP %.0f (%.0f us); CTA lifetimes fill %.3f of 296 slots x wall time at 1.965 GHz" % (tot / 1024 / K, tot / 1024 / K / 1965.0, out[15] / (296 * t * 1.965e9)))
m.close()
