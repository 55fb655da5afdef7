#!/usr/bin/env python
"""bench.py — batched MPC steps/s on B200 for the workloads of BASELINE.json (metric: "MPC QP steps/sec, batched, whole box").

A "step" is one pass of the hot path over the whole batch: compute_time_steps -> compute_linearization_nodes -> update_QP
(linearisation + envelope + HJI constraint) -> solve (ADMM) -> get_next_control, followed by the plant rollout of `simulate`
(reference src/model_predictive_control.jl:87-98) so that every step solves a new, warm-started QP.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4] [--batch B] [--impl reference] [--other-configs all|none|2,3,4]

--config selects the entry of BASELINE.json `configs` that the JSON line reports (default 1, the configuration the metric is quoted on):
  1  1024 X1 vehicles per GPU, coupled lat-long MPC (N = 31)                      coupled_lat_long.jl:62-142,315-368
  2  decoupled lat-long MPC, 8,192 vehicles per GPU (65,536 on 8 GPUs)            decoupled_lat_long.jl:52-104,228-273
  3  coupled MPC with the HJI constraint ACTIVE for about half of 16,384 scenarios, 13x13x9^5 grid (319 MB) in HBM   coupled_lat_long.jl:341-346
  4  closed-loop Monte-Carlo: 131,072 perturbed initial states per GPU (1,048,576 on 8) x 200 steps on the device    model_predictive_control.jl:80-100
The default run (config 1) also measures configs 2-4 at their per-GPU size and appends them as `other_configs` to its ONE JSON line.

N > 1 is launched by torchrun (one process per GPU); the vehicle batch is sharded with no data-path collective (weak scaling, the per-GPU
batch above on every rank, a different seed per rank) and the final controls / statistics are gathered once over NCCL (pgn_gather) after the
timed region.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_qp_steps_per_sec"
UNIT = "steps/s"
SETTLE = 30     # closed-loop steps run before the warm-up: the perturbed cold start (a handful of QPs need thousands of ADMM iterations) is reported separately
DT = 0.01
FAR = np.array([1e4, 1e4, 0.0, 5.0])      # other car far outside the HJI grid: constraint evaluated, inactive

CONFIGS = {
    1: dict(kind="coupled", batch=1024, seed=17, hji=False, guards=False,
            label="configs[1]: batch of 1024 X1 vehicles per GPU, coupled lat-long MPC (N_short=10, N_long=20, N=31), 64 synthetic 1000-node trajectories, "
                  "closed loop dt=0.01, timed after %d settling steps from the perturbed cold start" % SETTLE),
    2: dict(kind="decoupled", batch=8192, seed=1, hji=False, guards=False,
            label="configs[2]: decoupled lat-long MPC (lateral QP + longitudinal PD), 8,192 vehicles per GPU (65,536 on 8 GPUs), closed loop dt=0.01, "
                  "timed after %d settling steps" % SETTLE),
    3: dict(kind="coupled", batch=16384, seed=3, hji=True, guards=False,
            label="configs[3]: coupled MPC with the HJI value/gradient safety constraint, 16,384 scenarios per GPU, analytic 13x13x9^5 grid (319 MB) resident in HBM, "
                  "other car placed so that ~50 %% of the scenarios start with an active constraint and ~10 %% outside the grid; timed after %d settling steps" % SETTLE),
    4: dict(kind="coupled", batch=131072, seed=2, hji=False, guards=True, mc_steps=200,
            label="configs[4]: closed-loop Monte-Carlo, 131,072 perturbed initial states per GPU (1,048,576 on 8) x 200 timesteps, linearize -> QP -> rollout fully "
                  "on the device (pgn_simulate), NaN guard of the callback on; timed from the cold start over all 200 steps"),
}


def trajectories():
    from pigeon.jl_b200 import synthetic
    return synthetic.synthetic_trajectories(seed=synthetic.SEED, n_traj=64, n_nodes=1000, ds=0.25)


def make_workload(cfg_idx, B, seed_shift=0, trajs=None):
    """Synthetic batch of BASELINE config `cfg_idx` (SURVEY.md 8d generator): trajectories, trajectory ids, states, controls, t0, other cars."""
    from pigeon.jl_b200 import synthetic
    c = CONFIGS[cfg_idx]
    trajs = trajectories() if trajs is None else trajs
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + c["seed"] + seed_shift)
    other = np.tile(FAR, (B, 1))
    if c["hji"]:
        # other car in the EGO BODY frame of the HJI relative state (HJI_computation.jl:20-24: x_rel = M (p_other - p_ego), M = [-sin psi, cos psi; -cos psi, -sin psi]):
        # half of the scenarios inside the V <= eps ellipse of the analytic grid ((dE/4)^2 + (dN/2)^2 <= ~1), 40 % outside it but in the grid, 10 % out of the grid
        rng = np.random.default_rng(synthetic.SEED + 4 + seed_shift)
        cls = rng.random(B)
        R = np.where(cls < 0.5, rng.uniform(0.25, 0.95, B), rng.uniform(1.3, 3.0, B))
        ang = rng.uniform(-np.pi, np.pi, B)
        r0, r1 = 4.0 * R * np.cos(ang), 2.0 * R * np.sin(ang)
        psi = state[:, 2]
        other[:, 0] = state[:, 0] + (-np.sin(psi) * r0 - np.cos(psi) * r1)
        other[:, 1] = state[:, 1] + (np.cos(psi) * r0 - np.sin(psi) * r1)
        other[:, 2] = psi + rng.normal(0, 0.3, B)
        other[:, 3] = np.clip(state[:, 3] + rng.normal(0, 1.0, B), 1.5, 14.0)
        far = cls >= 0.9
        other[far, 0] += 500.0
    return trajs, tid, state, control, t0, other


def make_mpc(p, cfg_idx, trajs, tid, B, device):
    c = CONFIGS[cfg_idx]
    ctor = p.BatchedCoupledTrajectoryTrackingMPC if c["kind"] == "coupled" else p.BatchedDecoupledTrajectoryTrackingMPC
    mpc = ctor(p.X1(), trajs, B, trajectory_index=tid, device=device)
    if c["guards"]:
        # the callback's NaN guard (ros_integration.jl:134-147): ~0.7 % of the perturbed states of config 4 give a primal-infeasible QP on the second step
        # (the CPU oracle reports the same status and iteration count); OSQP then returns NaN and, unguarded, the NaN control poisons the state for good
        mpc.set_guards(nan_fallback=True, pause_below_speed=0.0)
    return mpc


_HJI = {}


def hji_cache(p):
    if "c" not in _HJI:
        from pigeon.jl_b200 import synthetic
        knots, V, gV = synthetic.analytic_hji_grid()
        _HJI["c"] = p.HJICache(knots, V, gV)
    return _HJI["c"]


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.stamps, self.proc, self.gpu = [], [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())
            self.stamps.append(time.perf_counter())

    def stop(self, t_from=None):
        """Samples taken at or after `t_from` (perf_counter); if none landed there, every sample of the loaded region."""
        if t_from is not None:
            keep = [r for r, t in zip(self.rows, self.stamps) if t >= t_from]
            if keep:
                self.rows = keep
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def config_dict(cfg_idx, B, world, steps, warmup):
    """The `config` object of the JSON line — the SAME for the B200 arm and the reference (CPU) arm of one command line."""
    c = CONFIGS[cfg_idx]
    return {"workload": c["label"], "baseline_config_index": cfg_idx, "controller": c["kind"], "batch_per_gpu": B, "global_batch": B * world,
            "horizon_nodes": 31, "dt": DT, "settle_steps": 0 if cfg_idx == 4 else SETTLE, "timed_steps": c.get("mc_steps", steps), "warmup_steps": 0 if cfg_idx == 4 else warmup,
            "seed": "0x5049474E + %d (+1000 per rank)" % c["seed"],
            "l2": "inputs larger than L2 are not needed: the per-step working set (QP records + ADMM iterates, ~60 KB per vehicle) is rewritten by every step, nothing is re-read across timed iterations",
            "parallelism": f"batch sharded over {world} GPU(s), no hot-path collective"}


# ------------------------------------------------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------------------------------------------------
def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as o
    return o


def cpu_closed_loop(cfg_idx, B_sample, steps, warmup, nthreads=0, native=True, settle=None, return_state=False, sel=None):
    """Times the CPU oracle (oracle/, a restatement of the reference algorithm: kind "port") on the workload of config `cfg_idx`.
    sel: indices of the vehicles of the full-size batch to run (default: the first B_sample of a B_sample batch)."""
    o = oracle_module()
    flags = o.use_native_build() if native else o.build_flags
    c = CONFIGS[cfg_idx]
    settle = (0 if cfg_idx == 4 else SETTLE) if settle is None else settle
    if sel is None:
        trajs, tid, state, control, t0, other = make_workload(cfg_idx, B_sample)
    else:
        trajs, tid, state, control, t0, other = sel
        B_sample = len(tid)
    cache, ms = {}, []
    hc = None
    if c["hji"]:
        from pigeon.jl_b200 import synthetic
        knots, V, gV = synthetic.analytic_hji_grid()
        hc = o.HjiCache(knots, V, gV)
    kind = o.MPC_COUPLED if c["kind"] == "coupled" else o.MPC_DECOUPLED
    for i in range(B_sample):
        j = int(tid[i])
        if j not in cache:
            cache[j] = o.Trajectory(**{k: trajs[k][j] for k in o.TRAJ_FIELDS})
        m = o.Mpc(kind)
        m.set_trajectory(cache[j])
        if hc is not None:
            m.set_hji(hc)
        m.set_state(state[i], control[i], other4=other[i])
        ms.append(m)
    cores = o.max_threads() if nthreads <= 0 else nthreads
    for k in range(settle + warmup):
        o.batch_step(ms, t0 + DT * k, rollout=True, nthreads=cores)
    t = time.perf_counter()
    for k in range(steps):
        o.batch_step(ms, t0 + DT * (settle + warmup + k), rollout=True, nthreads=cores)
    el = time.perf_counter() - t
    iters = float(np.mean([m.stats()["iter"] for m in ms]))
    r = {"value": B_sample * steps / el, "unit": UNIT, "cores": cores, "kind": "port", "build_flags": flags,
         "sample": f"{B_sample} vehicles x {steps} closed-loop steps after {settle} settling + {warmup} warm-up steps, oracle (C++ restatement, std::thread over vehicles, {flags}), mean ADMM iters {iters:.1f}",
         "ms_per_step": el / steps * 1e3}
    if return_state:
        r["final_state"] = np.array([m.get_state()[0] for m in ms])
    return r


def cpu_single_vehicle(steps=400):
    """BASELINE configs[0] — the reference's own operating point: ONE X1 vehicle, coupled MPC (N = 31) tracking test/path/skidpadoval.world
    (converted as TrajectoryTube(p::path) does), closed loop on ONE host thread; per-step wall time of the four calls + propagate."""
    o = oracle_module()
    w = np.load(os.path.join(ROOT, "tests", "golden", "world_skidpadoval.npz"))
    s_, V = w["s_m"], w["UxDes_mps"]
    t = np.concatenate([[0.0], np.cumsum(2 * np.diff(s_) / (V[:-1] + V[1:]))])
    traj = o.Trajectory(t=t, s=s_, V=V, A=w["AxDes_mps2"], E=w["posE_m"], N=w["posN_m"], psi=w["psi_rad"], kappa=w["k_1pm"], theta=w["grade_rad"],
                        phi=np.zeros(len(s_)), edge_L=w["edgeL_m"], edge_R=w["edgeR_m"])
    m = o.Mpc(o.MPC_COUPLED)
    m.set_trajectory(traj)
    m.set_state(np.array([w["posE_m"][0], w["posN_m"][0], w["psi_rad"][0], 6.0, 0.0, 0.0]), np.zeros(3), other4=FAR)
    ts_, its = [], []
    for k in range(steps):
        a = time.perf_counter()
        m.simulate_step(DT * k)
        ts_.append((time.perf_counter() - a) * 1e3)
        its.append(m.stats()["iter"])
    ts_ = np.array(ts_[5:])
    return {"workload": "configs[0]: single X1 vehicle, coupled lat-long MPC (N = 31) on test/path/skidpadoval.world, closed loop, 1 host thread", "steps": steps,
            "ms_per_step": {"mean": float(ts_.mean()), "p50": float(np.percentile(ts_, 50)), "p99": float(np.percentile(ts_, 99))}, "steps_per_s": float(1e3 / ts_.mean()),
            "mean_iters": float(np.mean(its)), "reference_budget_ms": 10.0}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg_idx = args.config
    B = args.batch or CONFIGS[cfg_idx]["batch"]
    # bounded sample of the workload: at most 1024 vehicles per step (the CPU throughput does not depend on the batch size beyond the thread count)
    Bs = min(B, 1024 if cfg_idx != 3 else 512)
    steps, warm = (args.steps, args.warmup) if cfg_idx != 4 else (min(CONFIGS[4]["mc_steps"], 40), 0)
    r = cpu_closed_loop(cfg_idx, Bs, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg_idx, B, world, args.steps, args.warmup),
            "note": "the reference's Julia cannot run here (no julia in the image, no network); CPU arm = the C++ port of its algorithm on all host threads, built on this box with " + r["build_flags"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "build_flags")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    try:
        line["single_vehicle"] = cpu_single_vehicle()
    except Exception as ex:
        line["single_vehicle"] = {"error": repr(ex)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def iter_stats(st):
    it = st["iters"]
    return {"mean_iters": float(it.mean()), "p50_iters": float(np.median(it)), "p99_iters": float(np.percentile(it, 99)), "max_iters": int(it.max()),
            "pct_not_solved": float((st["status"] != 1).mean() * 100)}


def run_settled(cx, p, cfg_idx, B, K, W, seed_shift, want_stage=True, trajs=None):
    """Closed loop of config 1 / 2 / 3 on the device: cold start, SETTLE steps (timed as `cold_start`), W warm-up steps, K timed steps as ONE
    pgn_simulate_device call (max over ranks).  Returns the result dict and the live controller."""
    import torch
    trajs, tid, state, control, t0, other = make_workload(cfg_idx, B, seed_shift, trajs)
    mpc = make_mpc(p, cfg_idx, trajs, tid, B, cx.local)
    mpc.set_stream(cx.stream.cuda_stream)
    parts = mpc.set_pipeline_parts(cx.args.parts)
    extra = {}
    if CONFIGS[cfg_idx]["hji"]:
        mpc.set_HJI_cache(hji_cache(p))
    mpc.set_state(state, control, other)
    d_base = torch.tensor(t0, dtype=torch.float64, device=cx.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # one-time costs stay outside the cold-start region (it is about cold SOLVER state, not a cold driver): CUDA loads a kernel's module at its
    # first launch (a throwaway 64-vehicle handle runs two steps) and this handle's round graphs are captured and instantiated by a simulate
    # call of zero steps.  Measured: the region read 203 k steps/s on one box and 300 k on another before this, for the same 30 steps.
    nw = min(64, B)
    warm = make_mpc(p, cfg_idx, trajs, tid[:nw], nw, cx.local)
    warm.set_state(state[:nw], control[:nw], other[:nw] if other is not None else None)
    warm.simulate_device(t0[:nw], DT, 2)
    warm.close()
    mpc.simulate_device_async(d_base.data_ptr(), DT, 0, k0=0)
    mpc.synchronize()
    cx.barrier()
    ev[0].record(cx.stream)
    mpc.simulate_device_async(d_base.data_ptr(), DT, 1, k0=0)
    if CONFIGS[cfg_idx]["hji"]:
        cx.barrier()
        Vv, _ = mpc.hji_values()
        extra["hji_first_step"] = {"pct_active": float((Vv <= 0.05).mean() * 100), "pct_out_of_grid": float(np.isinf(Vv).mean() * 100)}
    mpc.simulate_device_async(d_base.data_ptr(), DT, SETTLE - 1, k0=1)
    mpc.synchronize()              # deferred solves: the vehicles that fell behind are caught up INSIDE the region that caused them
    ev[1].record(cx.stream)
    cx.barrier()
    cold_ms = ev[0].elapsed_time(ev[1])
    mpc.simulate_device_async(d_base.data_ptr(), DT, W, k0=SETTLE)
    mpc.synchronize()
    cx.barrier()
    mpc.stage_ms(reset=True)
    ev[2].record(cx.stream)
    mpc.simulate_device_async(d_base.data_ptr(), DT, K, k0=SETTLE + W)
    mpc.synchronize()
    ev[3].record(cx.stream)
    cx.barrier()
    ms_own = ev[2].elapsed_time(ev[3])
    ms = cx.max_over_ranks(ms_own)
    sm_ = mpc.stage_ms(reset=True)
    launches = sm_["launches"]
    st = mpc.stats()
    extra["by_rank"] = cx.gather_ranks([ms_own / K, float(st["iters"].max()), float(st["iters"].mean()), float(sm_["catchup_rounds"])],
                                       ["ms_per_step", "max_iters_last_step", "mean_iters_last_step", "catchup_rounds"])
    if CONFIGS[cfg_idx]["hji"]:
        Vv, _ = mpc.hji_values()
        extra["hji_last_step"] = {"pct_active": float((Vv <= 0.05).mean() * 100), "pct_out_of_grid": float(np.isinf(Vv).mean() * 100)}
    res = {"workload": CONFIGS[cfg_idx]["label"], "batch_per_gpu": B, "n_gpus": cx.world, "value": cx.world * B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
           "steps": K, "warmup": W, "settle_steps": SETTLE, "pipeline_parts": parts, "gpu_launches": int(launches),
           "cold_start": {"steps": SETTLE, "value": B * SETTLE / (cold_ms * 1e-3), "unit": UNIT + " (this rank)", "ms_per_step": cold_ms / SETTLE},
           "admm": iter_stats(st), "qp": {"n": mpc.n, "m": mpc.m, "nnzA": mpc.nnzA, "admm_threads": mpc.qp_program["admm_threads"], "admm_smem_bytes": mpc.qp_program["admm_smem_bytes"]}}
    res.update(extra)
    if want_stage:
        d_t0 = d_base + DT * (SETTLE + W + K)
        mpc.set_profiling(1)
        for i in range(3):
            mpc.step_rollout_device(d_t0.data_ptr(), None, DT); d_t0 += DT
        stage = mpc.stage_ms(reset=True); mpc.set_profiling(0)
        res["stage_ms_per_step"] = {k: stage[k] / 3 for k in ("nodes", "linearize", "hji", "admm", "controls", "rollout")}
    return res, mpc, (trajs, tid, state, control, t0, other)


def run_monte_carlo(cx, p, B, seed_shift, trajs=None, oracle_check=True):
    """configs[4]: B perturbed initial states x 200 closed-loop steps as ONE pgn_simulate call from the cold start, history of the tracking error
    recorded on the device every 20 steps; final-state check of a 1,024-vehicle subsample against the CPU oracle (rank 0)."""
    import torch
    c = CONFIGS[4]
    NS = c["mc_steps"]
    t_gen = time.perf_counter()
    trajs, tid, state, control, t0, other = make_workload(4, B, seed_shift, trajs)
    t_gen = time.perf_counter() - t_gen
    mpc = make_mpc(p, 4, trajs, tid, B, cx.local)
    mpc.set_stream(cx.stream.cuda_stream)
    parts = mpc.set_pipeline_parts(cx.args.parts)
    mpc.set_history(NS // 20, 20)
    d_base = torch.tensor(t0, dtype=torch.float64, device=cx.dev)
    pin_q = torch.from_numpy(state).pin_memory(); pin_u = torch.from_numpy(control).pin_memory(); pin_o = torch.from_numpy(other).pin_memory()
    cx.barrier()
    tw = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(cx.stream)
    mpc.set_state(pin_q.numpy(), pin_u.numpy(), pin_o.numpy())      # e2e: the job's inputs go in from pinned host memory ...
    e[1].record(cx.stream)
    mpc.simulate_device_async(d_base.data_ptr(), DT, NS, k0=0)
    mpc.synchronize()              # includes the catch-up rounds of the vehicles whose solves were deferred
    e[2].record(cx.stream)
    cx.barrier()
    q, u = mpc.get_state()                                          # ... and its result (final states and controls) comes back
    wall_ms = (time.perf_counter() - tw) * 1e3
    ms = cx.max_over_ranks(e[1].elapsed_time(e[2]))
    e2e_ms = cx.max_over_ranks(max(wall_ms, e[0].elapsed_time(e[2])))
    st = mpc.stats()
    qs, xs, us, ps = mpc.history()
    e_hist = np.abs(xs[:, :, 5])                                    # |e| (lateral error of node 1) every 20 steps
    res = {"workload": c["label"], "batch_per_gpu": B, "n_gpus": cx.world, "value": cx.world * B * NS / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / NS, "steps": NS,
           "seconds": ms * 1e-3, "pipeline_parts": parts, "host_workload_generation_s": t_gen,
           "e2e": {"value": cx.world * B * NS / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 13 * 8 / NS, "d2h_bytes_per_step": B * 9 * 8 / NS,
                   "call": "pgn_set_state (pinned host arrays) + pgn_simulate_device (200 steps) + pgn_get_state, wall clock"},
           "pct_finite": float(np.isfinite(q).all(axis=1).mean() * 100), "admm_last_step": iter_stats(st),
           "lateral_error_m_by_step": {str(20 * i): {"p50": float(np.nanmedian(e_hist[i])), "p99": float(np.nanpercentile(e_hist[i], 99)), "max": float(np.nanmax(e_hist[i]))} for i in range(e_hist.shape[0])}}
    if oracle_check and cx.rank == 0:
        # final tracking error of a 1,024-vehicle subsample on the GPU and on the CPU oracle, and the deviation between the two final states
        sub = np.arange(0, B, max(1, B // 1024))[:1024]
        sel = (trajs, tid[sub], state[sub], control[sub], t0[sub], other[sub])
        t_or = time.perf_counter()
        r = cpu_closed_loop(4, len(sub), NS, 0, settle=0, return_state=True, sel=sel, native=False)
        t_or = time.perf_counter() - t_or
        qo = r["final_state"]

        def dist(qq):
            return np.array([np.sqrt(np.min((trajs["E"][tid[i]] - qq[k, 0]) ** 2 + (trajs["N"][tid[i]] - qq[k, 1]) ** 2)) for k, i in enumerate(sub)])
        ok = np.isfinite(qo).all(axis=1) & np.isfinite(q[sub]).all(axis=1)
        dg, do = dist(q[sub]), dist(qo)
        dev_pos = np.hypot(q[sub][ok, 0] - qo[ok, 0], q[sub][ok, 1] - qo[ok, 1])
        pct = lambda a: {"p50": float(np.median(a)), "p99": float(np.percentile(a, 99)), "max": float(np.max(a))}
        res["oracle_subsample"] = {"vehicles": int(len(sub)), "both_finite": int(ok.sum()), "oracle_seconds": t_or, "oracle_has_no_nan_guard": True,
                                   "final_distance_to_path_m": {"gpu": pct(dg[ok]), "oracle": pct(do[ok])},
                                   "final_position_deviation_gpu_vs_oracle_m": pct(dev_pos)}
    mpc.set_history(0)
    return res, mpc


def hji_roofline(cx, p, mpc, M=1 << 24):
    """Stand-alone HJI micro-benchmark: M uniformly random in-grid queries against the 319 MB grid (>> 126 MB L2)."""
    import torch
    from pigeon.jl_b200 import synthetic
    g = torch.Generator(device=cx.dev); g.manual_seed(7)
    lo = torch.tensor([r[0] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=cx.dev)[:, None]
    hi = torch.tensor([r[1] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=cx.dev)[:, None]
    x = (lo + (hi - lo) * (0.001 + 0.998 * torch.rand((7, M), dtype=torch.float64, device=cx.dev, generator=g))).contiguous()
    V = torch.empty(M, dtype=torch.float64, device=cx.dev); gV = torch.empty((7, M), dtype=torch.float64, device=cx.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(mode):
        mpc.set_hji_lookup_order(mode)
        for _ in range(2):
            mpc.hji_lookup_device(M, x.data_ptr(), V.data_ptr(), gV.data_ptr())
        best = 1e30
        for _ in range(3):
            e0.record(cx.stream)
            mpc.hji_lookup_device(M, x.data_ptr(), V.data_ptr(), gV.data_ptr())
            e1.record(cx.stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    ms_input = timed(0)          # the queries in the order they were given: a random 4 KB gather per query
    Vi = V.clone()
    best = timed(-1)             # default: counting sort by grid cell + the same gather in cell order (sort time included)
    same = bool(torch.equal(Vi, V))
    mpc.set_hji_lookup_order(-1)
    peak, src = cx.hbm_peak
    ach = M * 4096.0 / (best * 1e-3) / 1e9
    out = {"kernel": "k_hji_keys + scan + k_hji_scatter + k_hji_lookup_perm (cell-ordered lookup)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src, "queries": M, "launch_ms": best,
           "queries_per_s": M / (best * 1e-3), "algorithmic_bytes_per_query": 4096, "traffic": None, "bit_identical_to_input_order": same,
           "input_order": {"kernel": "k_hji_lookup", "launch_ms": ms_input, "achieved": M * 4096.0 / (ms_input * 1e-3) / 1e9, "frac": M * 4096.0 / (ms_input * 1e-3) / 1e9 / peak,
                           "note": "random gather: every 64-byte corner pair costs ~100 bytes of DRAM traffic (tools/ubench/hji_fetch.cu), 6.6 KB per query, DRAM pipe at 96 % of the measured copy bandwidth"},
           "note": "128 corners x one 32-byte record {gradV[7], V} per query; the 319 MB grid is larger than the 126 MB L2; frac > 1 is possible in cell order because neighbouring queries share corners through L2 (algorithmic bytes count every query's 4 KB)"}
    tr = ncu_metric("r2_hji_lookup_ncu.md")
    if tr:
        out["traffic"] = tr["dram_bytes"] * (M / tr["queries"] if tr.get("queries") else 1.0)
        out["traffic_source"] = tr["source"] + " (DRAM bytes read + written by the kernels of one cell-ordered lookup, scaled to this launch's query count)"
    tr0 = ncu_metric("r1_hji_lookup_ncu_full.md")
    if tr0:
        out["input_order"]["traffic"] = tr0["dram_bytes"] * M / (1 << 22)
        out["input_order"]["traffic_source"] = tr0["source"] + " (one launch of 2^22 queries, scaled)"
    return out


def ncu_metric(*names):
    """dram__bytes_read.sum + dram__bytes_write.sum (and the shared-memory wavefront share) of the first committed `ncu --set full` summary that exists."""
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for nm in names:
        path = os.path.join(ROOT, "profiles", nm)
        if not os.path.exists(path):
            continue
        txt = open(path).read()
        rd = re.search(r"dram__bytes_read.sum`\) \| ([0-9.]+) \| (\w+)", txt); wr = re.search(r"dram__bytes_write.sum`\) \| ([0-9.]+) \| (\w+)", txt)
        if not rd or not wr:
            continue
        out = {"dram_bytes": float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)], "source": "profiles/" + nm}
        wf = re.search(r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed`\) \| ([0-9.]+)", txt)
        if wf:
            out["smem_pct"] = float(wf.group(1))
        qn = re.search(r"queries per launch: (\d+)", txt)
        if qn:
            out["queries"] = int(qn.group(1))
        return out
    return None


def measured_peaks():
    """FP64-FMA and shared-memory peaks measured on this GPU by tools/ubench/peaks (run now when the binary is there, else the committed numbers)."""
    exe = os.path.join(ROOT, "tools", "ubench", "peaks")
    try:
        if os.path.exists(exe):
            o = json.loads(subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1])
            o["source"] = "tools/ubench/peaks run on this GPU beside the benchmark"
            return o
    except Exception:
        pass
    try:
        o = json.load(open(os.path.join(ROOT, "profiles", "peaks_r2.json")))
        o["source"] = "profiles/peaks_r2.json (tools/ubench/peaks on a B200 of this pool)"
        return o
    except Exception:
        return {"fp64_fma_tflops": 37.0, "source": "nominal (64 DFMA/clk/SM x 148 SMs x 1.965 GHz)"}


def run_gpu(args):
    import torch
    import pigeon.jl_b200 as p
    cx = Ctx()
    cx.args = args
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if cx.world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(cx.local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local))
    torch.cuda.set_device(cx.local)
    cx.dev = torch.device("cuda", cx.local)
    cx.stream = torch.cuda.Stream(device=cx.dev)          # a real (non-NULL) stream: the library launches on it, the events are recorded on it
    torch.cuda.set_stream(cx.stream)
    world, rank, dev, stream = cx.world, cx.rank, cx.dev, cx.stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    def gather_ranks(vals, names):
        """Per-rank values (the throughput is the max over ranks of the timed region: this shows which rank's batch set it)."""
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is None:
            return {n: [float(v)] for n, v in zip(names, vals)}
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return {n: [float(o[i].item()) for o in out] for i, n in enumerate(names)}
    cx.barrier, cx.max_over_ranks, cx.gather_ranks = barrier, max_over_ranks, gather_ranks
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cx.hbm_peak = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    cfg_idx = args.config
    B, K, W = args.batch or CONFIGS[cfg_idx]["batch"], args.steps, args.warmup
    seed_shift = 1000 * rank + args.seed_shift
    trajs = trajectories()
    sampler = ClockSampler(cx.local)
    if rank == 0:
        sampler.start()           # nvidia-smi needs ~0.2 s to deliver its first sample
        time.sleep(0.3)
    t_timed = time.perf_counter()

    if cfg_idx == 4:
        res, mpc = run_monte_carlo(cx, p, B, seed_shift, trajs)
        clocks = sampler.stop(t_timed) if rank == 0 else None
        gathered = final_gather(cx, p, mpc, dist)
        if rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"], "warmup": 0, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(4, B, world, K, W), "e2e": res["e2e"],
                    "gpu_launches": None, "clocks": clocks, "details": res, "gather": gathered}
            print(json.dumps(line), flush=True)
        mpc.close()
        finish(dist)
        return

    res, mpc, wl = run_settled(cx, p, cfg_idx, B, K, W, seed_shift, trajs=trajs)
    trajs, tid, state, control, t0, other = wl
    ms_max = res["ms_per_step"] * K
    value = res["value"]
    st = mpc.stats()
    d_base = torch.tensor(t0, dtype=torch.float64, device=dev)
    d_t0 = d_base.clone()
    d_out = torch.zeros(3 * B, dtype=torch.float64, device=dev)
    kstep = [SETTLE + W + K + 3]

    def dev_step(record=None):
        if record is not None:
            q, u = mpc.get_state()
            record[0].append(q); record[1].append(u)
        torch.add(d_base, kstep[0] * DT, out=d_t0)
        mpc.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), DT)      # step + plant rollout (launched beside the QP solve)
        kstep[0] += 1

    # the same K steps as K per-step calls (pgn_step_rollout_device: the parts are joined at the end of every call), reported beside `value`
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record(stream)
    for _ in range(K):
        dev_step()
    ev3.record(stream)
    barrier()
    ms_calls = ev2.elapsed_time(ev3)
    clocks = sampler.stop(t_timed) if rank == 0 else None

    # in-kernel cycle shares of the ADMM phases (one extra step, not timed)
    mpc.set_profiling(2)
    mpc.admm_cycles(reset=True)
    dev_step()
    cyc = mpc.admm_cycles(reset=True)
    mpc.stage_ms(reset=True)
    mpc.set_profiling(0)
    stage = res["stage_ms_per_step"]
    admm_ms = stage["admm"]
    mean_iters = res["admm"]["mean_iters"]

    # ---------------- SURVEY.md 8(d) literal timing of config 2 (= BASELINE configs[1]): "1 cold step + 100 warm steps timed" ----------------
    mpc.reset_solver(); mpc.reset_solved()
    mpc.set_state(state, control, other)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    mpc.simulate_device_async(d_base.data_ptr(), DT, 101, k0=0)
    mpc.synchronize()
    c1.record(stream)
    barrier()
    ms101 = max_over_ranks(c0.elapsed_time(c1))
    cold_literal = {"steps": 101, "value": world * B * 101 / (ms101 * 1e-3), "unit": UNIT, "ms_per_step": ms101 / 101,
                    "note": "SURVEY.md 8(d) wording: 1 cold step + 100 warm steps from the perturbed cold start, one pgn_simulate_device call, nothing excluded; "
                            "the settled `value` starts after %d of these steps" % SETTLE}

    # ---------------- end-to-end arm (`e2e`): host buffers through the C ABI, copies inside the timed region ----------------
    # replay of the closed loop: pass 1 records the state before every step (device loop), pass 2 feeds them from pinned host memory
    mpc.reset_solver(); mpc.reset_solved()
    mpc.set_state(state, control, other)
    rec = ([], [])
    kstep[0] = 0
    for _ in range(SETTLE + W + K):
        dev_step(rec)
    nrec = len(rec[0])
    pin_q = torch.empty((nrec, B, 6), dtype=torch.float64).pin_memory()
    pin_u = torch.empty((nrec, B, 3), dtype=torch.float64).pin_memory()
    pin_t = torch.empty((nrec, B), dtype=torch.float64).pin_memory()
    pin_o = torch.empty((nrec, B, 3), dtype=torch.float64).pin_memory()
    pin_q.numpy()[:] = np.stack(rec[0]); pin_u.numpy()[:] = np.stack(rec[1])
    pin_t.numpy()[:] = t0[None, :] + DT * np.arange(nrec)[:, None]
    mpc.reset_solver(); mpc.reset_solved()
    qn, un_, tn, on = pin_q.numpy(), pin_u.numpy(), pin_t.numpy(), pin_o.numpy()
    lib, h = mpc._lib, mpc._h

    def e2e_step(k):
        lib.pgn_set_state(h, C.c_void_p(qn[k].ctypes.data), C.c_void_p(un_[k].ctypes.data), None, None)
        lib.pgn_step(h, C.c_void_p(tn[k].ctypes.data), C.c_void_p(on[k].ctypes.data))

    for k in range(SETTLE + W):
        e2e_step(k)
    barrier()
    tw = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(SETTLE + W, SETTLE + W + K):
        e2e_step(k)
    e1.record(stream)
    barrier()
    e2e_joined_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - tw) * 1e3))
    # the same replay through the pipelined calls: every step's inputs go in from pinned host memory and every step's controls come back,
    # with up to DEPTH steps in flight (the callback is fed MEASURED states, so step k+1 does not wait for step k's output)
    DEPTH = 3
    mpc.reset_solver(); mpc.reset_solved()

    def submit(k):
        lib.pgn_step_submit(h, C.c_void_p(qn[k].ctypes.data), C.c_void_p(un_[k].ctypes.data), None, C.c_void_p(tn[k].ctypes.data))

    def run_pipelined(k0, k1):
        for k in range(k0, k1):
            submit(k)
            if k - k0 >= DEPTH - 1:
                lib.pgn_step_collect(h, C.c_void_p(on[k - DEPTH + 1].ctypes.data))
        for k in range(max(k0, k1 - DEPTH + 1), k1):
            lib.pgn_step_collect(h, C.c_void_p(on[k].ctypes.data))
    run_pipelined(0, SETTLE + W)
    barrier()
    tw = time.perf_counter()
    run_pipelined(SETTLE + W, SETTLE + W + K)          # ends with every result of the K steps on the host
    e2e_wall = (time.perf_counter() - tw) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_wall)
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    e2e_joined = world * B * K / (e2e_joined_ms * 1e-3)

    # ---------------- per-call latency (BASELINE.json: "p50 per-step latency"): host wall clock around one C-ABI call, host buffers ----------------
    latency = None
    if rank == 0 and not args.no_latency:
        def pct(fn, n=200, warm=20):
            for i in range(warm):
                fn(i)
            ts_ = []
            for i in range(n):
                a_ = time.perf_counter(); fn(warm + i); ts_.append((time.perf_counter() - a_) * 1e3)
            return {"p50_ms": float(np.percentile(ts_, 50)), "p99_ms": float(np.percentile(ts_, 99)), "calls": n}
        lo, hi = SETTLE + W, SETTLE + W + K
        latency = {"batched_step": dict(pct(lambda i: e2e_step(lo + i % (hi - lo))), batch=B, call="pgn_set_state + pgn_step (host buffers)")}
        if CONFIGS[cfg_idx]["kind"] == "coupled":
            # the reference's deployment point: ONE vehicle through the callback entry point (one packed copy in, one CUDA graph, one copy out)
            one = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1, trajectory_index=tid[:1], device=cx.local)
            one.set_stream(stream.cuda_stream)
            one.set_state(state[:1], control[:1], other[:1])
            q1 = np.ascontiguousarray(np.stack(rec[0])[:, :1]); u1 = np.ascontiguousarray(np.stack(rec[1])[:, :1])
            o5 = np.zeros((1, 5)); st1 = np.zeros(1)

            def cb(i):
                k_ = min(i, nrec - 1)
                lib.pgn_from_autobox(one._h, C.c_void_p(q1[k_].ctypes.data), C.c_void_p(u1[k_].ctypes.data), None, C.c_void_p(st1.ctypes.data), C.c_void_p(o5.ctypes.data))
            latency["single_vehicle_callback"] = dict(pct(cb), batch=1, call="pgn_from_autobox (path-tracking mode; CUDA graph)")

            def five(i):
                k_ = min(i, nrec - 1)
                lib.pgn_set_state(one._h, C.c_void_p(q1[k_].ctypes.data), C.c_void_p(u1[k_].ctypes.data), None, None)
                lib.pgn_step(one._h, C.c_void_p(tn[k_][:1].copy().ctypes.data), C.c_void_p(o5.ctypes.data))
            one.reset_solver(); one.reset_solved()
            latency["single_vehicle_step"] = dict(pct(five), batch=1, call="pgn_set_state + pgn_step (stream launches)")
            one.close()

    gathered = final_gather(cx, p, mpc, dist)

    # ---------------- the other BASELINE configs at their per-GPU size (default run of config 1 only) ----------------
    others, hji_roof = {}, None
    want = [] if (cfg_idx != 1 or args.other_configs == "none") else ([2, 3, 4] if args.other_configs == "all" else [int(x) for x in args.other_configs.split(",")])
    n, m, nnzA, nnzL, prog, Nh, nlev = mpc.n, mpc.m, mpc.nnzA, mpc.nnzL, mpc.qp_program, mpc.N, mpc.n_levels
    if cfg_idx == 3 and rank == 0:
        hji_roof = hji_roofline(cx, p, mpc)
    mpc.close()
    for oc in want:
        t_oc = time.perf_counter()
        try:
            if oc == 4:
                r, m4 = run_monte_carlo(cx, p, CONFIGS[4]["batch"], seed_shift, trajs)
                m4.close()
            else:
                r, mo, _ = run_settled(cx, p, oc, CONFIGS[oc]["batch"], 30, 3, seed_shift, trajs=trajs)
                if oc == 3 and rank == 0:
                    hji_roof = hji_roofline(cx, p, mo)
                mo.close()
            r["bench_wall_s"] = time.perf_counter() - t_oc
            others["configs[%d]" % oc] = r
        except Exception as ex:      # an out-of-memory on a shared box must not lose the headline line
            others["configs[%d]" % oc] = {"error": repr(ex)}
    _HJI.clear()

    if rank == 0:
        Nk = n + m
        rec_len = 30 * 81 + 6 + 2 + 3 + 30 if CONFIGS[cfg_idx]["kind"] == "coupled" else 30 * 43 + 4 + 1 + 30
        pk = measured_peaks()
        fp64_peak = float(pk.get("fp64_fma_tflops", 37.0))
        # algorithmic HBM bytes of one ADMM launch: per QP the piece record in, warm iterates (x|z, y) in and out, solution x,y out, stats
        bytes_per_qp = 8 * (rec_len + 4 * Nk + n + m) + 40
        # algorithmic FP64 flops per QP, from the static programs of this QP: Ruiz (10 passes over A and diag P), factor
        # (3 flops per gather entry), range inverses, dense-tail sweep, then per iteration the forward/backward entries (2 flops each),
        # the dense tail mat-vec and ~15 flops per KKT row of vector updates; residual checks every 25 iterations
        Dm = prog["tail_dim"]
        flop_factor = 3 * prog["factor_entries"] + 3 * prog["inverse_entries"] + 3 * Dm * (Dm * (Dm + 1) // 2)
        flop_iter = 2 * (prog["l_slots"] + prog["backward_entries"]) + 2 * Dm * Dm + 15 * Nk
        flop_check = 4 * nnzA + 2 * n + 10 * Nk
        flop_qp = flop_factor + 10 * 2 * (nnzA + n) + mean_iters * flop_iter + (mean_iters / 25.0) * flop_check
        hbm_peak, peak_src = cx.hbm_peak
        roof = {"kernel": "k_admm", "bound": "hbm", "achieved": B * bytes_per_qp / (admm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": peak_src, "traffic": None, "launch_ms": admm_ms,
                "share_of_step": admm_ms / max(1e-9, sum(stage[k] for k in ("nodes", "linearize", "hji", "admm", "controls", "rollout"))),
                "fp64": {"achieved_tflops": B * flop_qp / (admm_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak, "peak_source": pk.get("source"), "flop_per_qp": flop_qp},
                "note": "one QP per CTA; the solve is a chain of dependent sparse triangular solves on chip: latency/occupancy-bound, neither HBM- nor tensor-bound (SURVEY.md 8d)",
                "limiter": {"what": "dependent-instruction latency of the slowest warp in each barrier interval (about 4.7 cycles per instruction of a lone warp), then shared-memory wavefronts",
                            "evidence": "tools/ubench/*.cu (B200 latencies), DESIGN.md 4.1, profiles/r2_admm_ncu_full.md"}}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["fp64"]["frac"] = roof["fp64"]["achieved_tflops"] / fp64_peak
        tr = ncu_metric("r2_admm_ncu_full.md", "r1b_admm_ncu_full.md")
        if tr:
            roof["traffic"] = tr["dram_bytes"]
            roof["traffic_source"] = tr["source"] + " (`ncu --set full` of one B = 1024 launch of this kernel; below the algorithmic bytes because the records and iterates of 1024 vehicles stay in the 126 MB L2)"
            if "smem_pct" in tr:
                roof["smem"] = {"pct_of_peak_wavefronts": tr["smem_pct"], "source": tr["source"], "measured_peak_bytes_per_clk_sm": pk.get("smem_lds64_bytes_per_clk_sm")}
        ncpu = min(B, 256)
        cpu = cpu_closed_loop(cfg_idx, ncpu, 3, 1) if not args.no_cpu else None
        cfg = config_dict(cfg_idx, B, world, K, W)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "details": {"horizon_nodes": Nh, "qp": {"n": n, "m": m, "nnzA": nnzA, "nnzL": nnzL, "levels": nlev, "program": prog},
                            "loop": "one pgn_simulate_device call of K closed-loop steps (the reference's `simulate` loop on the device)", "pipeline_parts": res["pipeline_parts"]},
                "per_step_calls": {"value": world * B * K / (ms_calls * 1e-3), "unit": UNIT + " (rank 0 time)", "ms_per_step": ms_calls / K,
                                   "call": "pgn_step_rollout_device x K, device-resident; the pipeline parts are joined at the end of every call"},
                "roofline": roof, "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "build_flags")} if cpu else None,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (6 + 3 + 1) * 8, "d2h_bytes_per_step": B * 3 * 8,
                        "call": "pgn_step_submit + pgn_step_collect per step (host buffers in and out every step, %d steps in flight, host wall clock incl. the final drain)" % DEPTH,
                        "joined": {"value": e2e_joined, "unit": UNIT, "call": "pgn_set_state + pgn_step per step: the caller waits for every step's controls before it sends the next state"}},
                "gpu_launches": res["gpu_launches"], "clocks": clocks, "admm": res["admm"],
                "stage_ms_per_step": stage, "admm_phase_share": {k: v / max(1.0, sum(cyc.values())) for k, v in cyc.items()},
                "latency": latency, "gather": gathered,
                "cold_start": dict(res["cold_start"], note="first %d closed-loop steps from the perturbed cold start; a few QPs per step run to thousands of iterations (max_iter 4000)" % SETTLE),
                "survey_8d_timing": cold_literal}
        for k in ("hji_first_step", "hji_last_step", "by_rank"):
            if k in res:
                line[k] = res[k]
        if hji_roof:
            line["roofline_hji"] = hji_roof
        if others:
            line["other_configs"] = others
        print(json.dumps(line), flush=True)
    finish(dist)


def final_gather(cx, p, mpc, dist):
    """Final gather of controls + statistics through the library's own NCCL collective (pgn_gather), outside the timed region."""
    if dist is None:
        return None
    import torch
    lib = mpc._lib
    idb = torch.zeros(128, dtype=torch.uint8)
    if cx.rank == 0:
        buf = (C.c_char * 128)()
        rc = lib.pgn_comm_unique_id(buf)
        if rc != 0:
            return {"error": lib.pgn_last_error().decode()}
        idb = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    idb = idb.to(cx.dev)
    dist.broadcast(idb, 0)                    # the launcher's job: ship the 128-byte id to every rank
    raw = bytes(idb.cpu().numpy().tolist())
    rc = lib.pgn_comm_init_rank(mpc._h, cx.world, cx.rank, raw)
    if rc != 0:
        return {"error": lib.pgn_last_error().decode()}
    n = cx.world * mpc.B
    ctrl, it, st = np.zeros((n, 3)), np.zeros(n, np.int32), np.zeros(n, np.int32)
    t = time.perf_counter()
    rc = lib.pgn_gather(mpc._h, C.c_void_p(ctrl.ctypes.data), C.c_void_p(it.ctypes.data), C.c_void_p(st.ctypes.data))
    el = time.perf_counter() - t
    if rc != 0:
        return {"error": lib.pgn_last_error().decode()}
    lib.pgn_comm_destroy(mpc._h)
    return {"call": "pgn_gather (ncclAllGather of controls [3][B] f64 + iters/status i32 per rank)", "ranks": cx.world, "controls": int(ctrl.shape[0]), "finite_controls_pct": float(np.isfinite(ctrl).all(axis=1).mean() * 100),
            "mean_iters_all_ranks": float(it.mean()), "seconds_incl_host_copy": el}


def finish(dist):
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="index into BASELINE.json `configs` (1 = the configuration the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=0, help="vehicles per GPU (default: the config's per-GPU size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--parts", type=int, default=0, help="pipeline parts of the fused calls (0 = automatic, 1 = off)")
    ap.add_argument("--other-configs", default="all", help="configs measured beside the default config-1 run and appended as `other_configs`: all | none | e.g. 2,3")
    ap.add_argument("--seed-shift", type=int, default=0, help="added to the workload seed (rank r uses 1000*r + this): reproduces another rank's batch at N = 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the per-call latency leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
