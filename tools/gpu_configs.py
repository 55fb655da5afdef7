#!/usr/bin/env python
"""GPU box: the secondary workloads of BASELINE.json (configs[2..4]) at the size ONE GPU carries in the 8-GPU run, measured the same way
as bench.py (closed loop, device-resident, CUDA events on the launch stream, settled region after SETTLE steps).  One JSON line each:
  configs[2]  decoupled lat-long MPC, 8,192 vehicles per GPU (65,536 on 8 GPUs)
  configs[3]  coupled MPC with the HJI constraint active, 16,384 scenarios, 13x13x9^5 grid (319 MB) resident in HBM
  configs[4]  closed-loop Monte-Carlo, 131,072 perturbed initial states per GPU (1,048,576 on 8) x 200 steps, fully on the device (pgn_simulate)
  configs[0'] the deployed horizon N = 16 (N_short = 5, N_long = 10) at B = 1024, as the secondary point of SURVEY.md 8d
Usage: python tools/gpu_configs.py [2] [3] [4] [n16]   (default: all)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pigeon.jl_b200 as p
from pigeon.jl_b200 import synthetic

SETTLE, K = 30, 30
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
which = sys.argv[1:] or ["n16", "2", "3", "4"]
trajs = synthetic.synthetic_trajectories(seed=synthetic.SEED, n_traj=64, n_nodes=1000, ds=0.25)


def closed_loop(mpc, B, t0, label, extra):
    mpc.set_stream(stream.cuda_stream)
    parts = mpc.set_pipeline_parts(int(os.environ.get("PGN_PARTS", "0")))
    d_base = torch.tensor(t0, dtype=torch.float64, device=dev)
    d_t0 = d_base.clone()
    d_out = torch.zeros(3 * B, dtype=torch.float64, device=dev)
    extra = dict(extra, pipeline_parts=parts, loop="pgn_simulate_device")

    def step():
        mpc.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), 0.01); d_t0.add_(0.01)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(stream)
    mpc.simulate_device_async(d_base.data_ptr(), 0.01, SETTLE, k0=0)
    e[1].record(stream)
    mpc.simulate_device_async(d_base.data_ptr(), 0.01, K, k0=SETTLE)
    e[2].record(stream)
    torch.cuda.synchronize()
    d_t0.copy_(d_base + 0.01 * (SETTLE + K))
    cold, ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    st = mpc.stats()
    mpc.set_profiling(1); mpc.stage_ms(reset=True)
    for _ in range(3):
        step()
    stage = mpc.stage_ms(reset=True); mpc.set_profiling(0)
    line = {"config": label, "batch": B, "steps_per_s_settled": B * K / (ms * 1e-3), "ms_per_step": ms / K, "steps_per_s_first_%d" % SETTLE: B * SETTLE / (cold * 1e-3),
            "admm_iters": {"mean": float(st["iters"].mean()), "p50": float(np.median(st["iters"])), "p99": float(np.percentile(st["iters"], 99)), "max": int(st["iters"].max())},
            "pct_not_solved": float((st["status"] != 1).mean() * 100), "qp": {"n": mpc.n, "m": mpc.m},
            "stage_ms_per_step": {k: stage[k] / 3 for k in ("nodes", "linearize", "hji", "admm", "controls", "rollout")}}
    line.update(extra)
    print(json.dumps(line), flush=True)


if "n16" in which:
    B = 1024
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 17)
    m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid, N_short=5, N_long=10)
    m.set_state(state, control, np.tile([1e4, 1e4, 0, 5.0], (B, 1)))
    closed_loop(m, B, t0, "coupled, deployed horizon N=16 (N_short=5, N_long=10), B=1024", {})
    m.close()
if "2" in which:
    B = 8192
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 1)
    m = p.BatchedDecoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    m.set_state(state, control, np.tile([1e4, 1e4, 0, 5.0], (B, 1)))
    closed_loop(m, B, t0, "configs[2]: decoupled lat-long MPC, 8,192 vehicles per GPU (65,536 on 8 GPUs)", {})
    m.close()
if "3" in which:
    B = 16384
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 3)
    rng = np.random.default_rng(synthetic.SEED + 4)
    other = np.zeros((B, 4))
    rad = rng.uniform(0.5, 12.0, B); ang = rng.uniform(-np.pi, np.pi, B)
    other[:, 0] = state[:, 0] + rad * np.cos(ang); other[:, 1] = state[:, 1] + rad * np.sin(ang)
    other[:, 2] = state[:, 2] + rng.normal(0, 0.5, B); other[:, 3] = rng.uniform(1.5, 12, B)
    other[::10, 0] += 500.0                                  # ~10 % out of the grid
    knots, V, gV = synthetic.analytic_hji_grid()
    m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    m.set_HJI_cache(p.HJICache(knots, V, gV))
    del V, gV
    m.set_state(state, control, other)
    m.set_stream(stream.cuda_stream)
    m.step(t0)
    Vv, _ = m.hji_values()
    m.reset_solver(); m.reset_solved(); m.set_state(state, control, other)
    closed_loop(m, B, t0, "configs[3]: coupled MPC + HJI constraint, 16,384 scenarios, 13x13x9^5 grid resident in HBM (other car fixed in the world frame)",
                {"hji_first_step": {"pct_active": float((Vv <= 0.05).mean() * 100), "pct_out_of_grid": float(np.isinf(Vv).mean() * 100)}})
    m.close()
if "4" in which:
    B, NSTEP = 131072, 200
    t_gen = time.perf_counter()
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 2)
    t_gen = time.perf_counter() - t_gen
    m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid)
    m.set_stream(stream.cuda_stream)
    m.set_state(state, control, np.tile([1e4, 1e4, 0, 5.0], (B, 1)))
    # the callback's NaN guard (ros_integration.jl:134-147): ~0.7 % of these perturbed states give a primal-infeasible QP on the second
    # step (the CPU oracle reports the same status and iteration count, tools/gpu_nan_debug.py); OSQP then returns NaN and, unguarded
    # (`simulate` has no guard), the NaN control poisons the state for good
    m.set_guards(nan_fallback=True, pause_below_speed=0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    m.simulate_device(t0, 0.01, NSTEP)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    q, u = m.get_state()
    st = m.stats()
    # final tracking error of every vehicle: lateral offset from its own trajectory (host-side nearest-segment scan on a subsample)
    sub = np.arange(0, B, B // 1024)
    err = []
    for i in sub:
        E, N = trajs["E"][tid[i]], trajs["N"][tid[i]]
        err.append(np.sqrt(np.min((E - q[i, 0]) ** 2 + (N - q[i, 1]) ** 2)))
    err = np.array(err)
    print(json.dumps({"config": "configs[4]: closed-loop Monte-Carlo, 131,072 initial states per GPU x 200 steps, on-device linearize -> QP -> rollout (pgn_simulate), NaN guard on",
                      "batch": B, "steps": NSTEP, "steps_per_s": B * NSTEP / (ms * 1e-3), "seconds": ms * 1e-3, "host_workload_generation_s": t_gen,
                      "pct_finite": float(np.isfinite(q).all(axis=1).mean() * 100), "pct_not_solved_last_step": float((st["status"] != 1).mean() * 100),
                      "final_distance_to_path_m_subsample_1024": {"p50": float(np.median(err)), "p99": float(np.percentile(err, 99)), "max": float(err.max())},
                      "admm_iters_last_step": {"mean": float(st["iters"].mean()), "max": int(st["iters"].max())}}), flush=True)
    m.close()
