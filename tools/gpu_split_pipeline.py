#!/usr/bin/env python
"""GPU box experiment: does splitting the B = 1024 batch of configs[1] over H handles on H streams (the thread-per-vehicle stages of one
part running while the ADMM kernel of another part drains) beat one handle?  Same closed loop and timing method as bench.py
(device-resident, settled region); the total time is bracketed by device synchronisation and an event pair on every stream.
Usage: python tools/gpu_split_pipeline.py [H ...]   (default 1 2 4; a negative H joins all streams after every step; B from $SPLIT_B)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pigeon.jl_b200 as p
from pigeon.jl_b200 import synthetic

SETTLE, W, K, B = 30, 5, 50, int(os.environ.get('SPLIT_B', '1024'))
dev = torch.device("cuda", 0)
trajs = synthetic.synthetic_trajectories(seed=synthetic.SEED, n_traj=64, n_nodes=1000, ds=0.25)
tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 17)
other = np.tile(np.array([1e4, 1e4, 0.0, 5.0]), (B, 1))

for H in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    joined = H < 0
    H = abs(H)
    parts = [(k * B // H, (k + 1) * B // H) for k in range(H)]
    hs = []
    for lo, hi in parts:
        m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, hi - lo, trajectory_index=tid[lo:hi])
        s = torch.cuda.Stream(device=dev)
        m.set_stream(s.cuda_stream)
        m.set_state(state[lo:hi], control[lo:hi], other[lo:hi])
        with torch.cuda.stream(s):
            d_t0 = torch.tensor(t0[lo:hi], dtype=torch.float64, device=dev)
            d_out = torch.zeros(3 * (hi - lo), dtype=torch.float64, device=dev)
        hs.append((m, s, d_t0, d_out))
    torch.cuda.synchronize()

    def step():
        for m, s, d_t0, d_out in hs:
            m.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), 0.01)
            with torch.cuda.stream(s):
                d_t0.add_(0.01)
        if joined:
            evs = []
            for m, s, _, _ in hs:
                e = torch.cuda.Event(); e.record(s); evs.append(e)
            for m, s, _, _ in hs:
                for e in evs:
                    s.wait_event(e)
    for _ in range(SETTLE + W):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in hs]
    tw = time.perf_counter()
    for (m, s, _, _), (a, b) in zip(hs, ev):
        a.record(s)
    for _ in range(K):
        step()
    for (m, s, _, _), (a, b) in zip(hs, ev):
        b.record(s)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - tw) * 1e3
    per = [a.elapsed_time(b) for a, b in ev]
    ms = max(max(per), wall) if H > 1 else per[0]
    its = np.concatenate([m.stats()["iters"] for m, *_ in hs])
    print(json.dumps({"handles": H, "joined_every_step": joined, "batch": B, "ms_per_step": ms / K, "steps_per_s": B * K / (ms * 1e-3), "wall_ms": wall, "stream_ms": per,
                      "mean_iters": float(its.mean())}), flush=True)
    for m, *_ in hs:
        m.close()

# ---- the same through the library's own pipeline parts (pgn_set_pipeline_parts): per-step calls (parts joined every call) and the on-device
# simulate loop (parts free-running over all K steps)
for P in [int(a) for a in os.environ.get("SPLIT_LIB", "").split()]:
    m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    s = torch.cuda.Stream(device=dev)
    m.set_stream(s.cuda_stream)
    m.set_pipeline_parts(P)
    m.set_state(state, control, other)
    with torch.cuda.stream(s):
        d_t0 = torch.tensor(t0, dtype=torch.float64, device=dev)
        d_out = torch.zeros(3 * B, dtype=torch.float64, device=dev)
        for _ in range(SETTLE + W):
            m.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), 0.01); d_t0.add_(0.01)
        torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(s)
        for _ in range(K):
            m.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), 0.01); d_t0.add_(0.01)
        b.record(s)
        m.simulate_device_async(d_t0.data_ptr(), 0.01, K, 0)
        c.record(s)
        torch.cuda.synchronize()
    print(json.dumps({"library_parts": m.pipeline_parts, "batch": B, "per_step_calls_ms": a.elapsed_time(b) / K, "per_step_calls_steps_per_s": B * K / (a.elapsed_time(b) * 1e-3),
                      "simulate_ms": b.elapsed_time(c) / K, "simulate_steps_per_s": B * K / (b.elapsed_time(c) * 1e-3), "mean_iters": float(m.stats()["iters"].mean())}), flush=True)
    m.close()
