#!/usr/bin/env python
"""GPU box: cycle shares inside the Ruiz equilibration of the ADMM kernel (profiling build, counters of thread 0 summed over CTAs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pigeon.jl_b200 as p
trajs, tid, state, control, t0, other = bench.make_workload(1, 1024, 0)
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
m.set_state(state, control, other)
for k in range(35):
    m.step(t0 + 0.01 * k); m.rollout(0.01)
m.set_profiling(2); m.admm_cycles(reset=True)
import time
t_w = time.perf_counter(); m.step(t0 + 0.35); t_w = time.perf_counter() - t_w
out = np.zeros(512); m._lib.pgn_get_admm_cycles(m._h, out.ctypes.data_as(__import__("ctypes").c_void_p), 1)
tot = out[:8].sum()
print("step (profiling build, joined): %.3f ms; ADMM CTA lifetimes fill %.2f of 296 slots x step time at 1.965 GHz" % (t_w * 1e3, out[15] / (296 * t_w * 1.965e9)))
print("phases (gather, ruiz, factor, solve, update, check, store, ticket):", np.round(out[:8] / tot, 3), "cycles per QP:", tot / 1024)
print("CTA lifetimes / sum of phases: %.3f  (prologue: tensor-memory allocation, static tables)" % (out[15] / tot))
names = ["norms + sqrt/rcp", "barrier after norms", "scale A", "vector updates", "block reduce", "cost scaling"]
rz = out[8:14]
print("norm gathers (cycles per pass and QP): %.0f  sqrt / reciprocal / publish: %.0f" % (out[14] / 1024 / 10, out[8] / 1024 / 10))
print("inside Ruiz (share of the kernel):", {n: round(v / tot, 4) for n, v in zip(names, rz)}, "sum", round(rz.sum() / tot, 4), "cycles per pass and QP:", np.round(rz / 1024 / 10))
m.close()
# factorisation breakdown (CTA 0 only: cycles of its thread 0 over the QPs it solved in this launch)
m2 = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1024, trajectory_index=tid)
m2.set_state(state, control, other)
for k in range(35):
    m2.step(t0 + 0.01 * k); m2.rollout(0.01)
m2.set_profiling(2); m2.admm_cycles(reset=True)
m2.step(t0 + 0.35)
out = np.zeros(512); m2._lib.pgn_get_admm_cycles(m2._h, out.ctypes.data_as(__import__("ctypes").c_void_p), 1)
init, inv, tail = out[126], out[127], out[128]
lv = out[136:186]
tot_f = init + inv + tail + lv.sum()
print("factor (CTA 0): init+scatter %.0f  levels %.0f  range inverses %.0f  dense tail sweep %.0f  (shares %.2f %.2f %.2f %.2f)" % (init, lv.sum(), inv, tail, init / tot_f, lv.sum() / tot_f, inv / tot_f, tail / tot_f))
print("levels:", np.round(lv / max(1.0, lv.sum()), 3))
sol = out[16:16 + 8]; tl = out[116]
print("solve phases (CTA 0):", np.round(sol / (sol.sum() + tl), 3), "dense tail matvec", round(tl / (sol.sum() + tl), 3))
m2.close()
