#!/usr/bin/env python
"""GPU box: timeline of the ADMM CTAs in the free-running closed loop (profiling 3: one %globaltimer record per CTA).
Prints, per pipeline part, the length of its ADMM phases and of the gaps between them (time steps + nodes + linearisation + HJI + controls +
propagation of that part), and how many ADMM CTAs are resident over time."""
import os, sys, time, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import pigeon.jl_b200 as p
K = int(os.environ.get("K", "40"))
B = int(os.environ.get("B", "1024"))
trajs, tid, state, control, t0, other = bench.make_workload(1, B, 0)
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
if os.environ.get("PARTS"): m.set_pipeline_parts(int(os.environ["PARTS"]))
m.set_state(state, control, other)
m.simulate_device(t0, 0.01, 35)
m.set_profiling(3)
m.simulate_device(t0 + 0.35, 0.01, K)
torch.cuda.synchronize()
lib = m._lib
cap = 1 << 19
buf = np.zeros(3 * cap, dtype=np.uint64); n = ctypes.c_int32(0)
lib.pgn_get_admm_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]
lib.pgn_get_admm_trace(m._h, buf.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n), 1)       # drop the records of the capture call
t = time.perf_counter(); m.simulate_device(t0 + 0.35 + 0.01 * K, 0.01, K); torch.cuda.synchronize(); t = time.perf_counter() - t
lib.pgn_get_admm_trace(m._h, buf.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n), 1)
n = n.value
r = buf[:3 * n].reshape(n, 3)
stamp = (r[:, 2] >> np.uint64(63)) == 1
sr = r[stamp]; r = r[~stamp]; n = len(r)
if len(sr):
    names = ["time steps", "nodes", "linearise + HJI", "ADMM (order + solve)", "controls", "join propagate + commit", "to the next round"]
    tt = sr[:, 0].astype(np.int64); sp = (sr[:, 2] & np.uint64(0xff)).astype(int); sg = ((sr[:, 2] >> np.uint64(8)) & np.uint64(0xff)).astype(int)
    for pt in sorted(set(sp)):
        o = np.argsort(tt[sp == pt], kind="stable"); t_ = tt[sp == pt][o]; g_ = sg[sp == pt][o]
        d = np.diff(t_) * 1e-3; frm = g_[:-1]
        print("part %d stage latencies us (mean / p90): " % pt + "; ".join("%s %.0f / %.0f" % (names[k], d[frm == k].mean(), np.percentile(d[frm == k], 90)) for k in range(7) if (frm == k).any()))
st, en, meta = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64), r[:, 2]
part = (meta & np.uint64(0xff)).astype(int); nqp = ((meta >> np.uint64(8)) & np.uint64(0xffffff)).astype(int); sm = (meta >> np.uint64(32)).astype(int)
T0, T1 = st.min(), en.max()
print("%d CTA records, %d QPs, wall %.3f ms / step (profiling build), trace span %.3f ms / step" % (n, nqp.sum(), t / K * 1e3, (T1 - T0) / K * 1e-6))
life = (en - st).astype(float)
print("CTA lifetime us: mean %.0f  p50 %.0f  p90 %.0f  max %.0f ; QPs per CTA mean %.2f" % (life.mean() / 1e3, np.median(life) / 1e3, np.percentile(life, 90) / 1e3, life.max() / 1e3, nqp.mean()))
# resident CTAs over time
ev = np.concatenate([np.stack([st, np.ones(n, dtype=np.int64)], 1), np.stack([en, -np.ones(n, dtype=np.int64)], 1)])
ev = ev[np.argsort(ev[:, 0], kind="stable")]
res = np.cumsum(ev[:, 1]); dtv = np.diff(ev[:, 0]); lvl = res[:-1]
tot = dtv.sum()
print("resident ADMM CTAs: time-average %.1f of 296; share of time with >= 290: %.2f, < 200: %.2f, < 100: %.2f, 0: %.2f" % (
    (lvl * dtv).sum() / tot, dtv[lvl >= 290].sum() / tot, dtv[lvl < 200].sum() / tot, dtv[lvl < 100].sum() / tot, dtv[lvl == 0].sum() / tot))
# per part: launches = clusters of records separated in start time
for pt in sorted(set(part)):
    s_, e_ = st[part == pt], en[part == pt]
    o = np.argsort(s_); s_, e_ = s_[o], e_[o]
    nv = len(s_) // K if K else len(s_)
    if nv == 0: continue
    ph_s = s_[: nv * K].reshape(K, nv).min(1); ph_e = np.sort(e_)[: nv * K].reshape(K, nv).max(1)
    length = (ph_e - ph_s) * 1e-3; gap = (ph_s[1:] - ph_e[:-1]) * 1e-3
    print("part %d: %d CTAs per launch; ADMM phase us mean %.0f (min %.0f max %.0f); gap to the next launch us mean %.0f (min %.0f max %.0f); first starts at %.0f us" % (
        pt, nv, length.mean(), length.min(), length.max(), gap.mean(), gap.min(), gap.max(), (ph_s[0] - T0) * 1e-3))
m.close()
