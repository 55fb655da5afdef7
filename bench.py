#!/usr/bin/env python
"""bench.py — headline benchmark: batched coupled lat-long MPC steps/s on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over the whole batch: compute_time_steps -> compute_linearization_nodes -> update_QP
(linearisation + envelope + HJI constraint) -> solve (ADMM) -> get_next_control, followed by the plant rollout of `simulate`
(reference src/model_predictive_control.jl:87-98) so that every step solves a new, warm-started QP.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one process per GPU); the vehicle batch is sharded with no data-path collective (weak scaling,
B vehicles per GPU) and the final controls/statistics are gathered once over NCCL after the timed region.
"""
import argparse
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
WORKLOAD = ("configs[1]: batch of 1024 X1 vehicles per GPU, coupled lat-long MPC (N_short=10, N_long=20, N=31), 64 synthetic 1000-node trajectories, "
            "closed loop dt=0.01, timed after %d settling steps from the perturbed cold start" % SETTLE)


def make_workload(B, seed_shift=0):
    from pigeon.jl_b200 import synthetic
    trajs = synthetic.synthetic_trajectories(seed=synthetic.SEED, n_traj=64, n_nodes=1000, ds=0.25)
    tid, state, control, t0 = synthetic.synthetic_batch(trajs, B, seed=synthetic.SEED + 17 + seed_shift)
    other = np.tile(np.array([1e4, 1e4, 0.0, 5.0]), (B, 1))   # other car far outside the HJI grid: constraint evaluated, inactive
    return trajs, tid, state, control, t0, other


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


def cpu_baseline(B_sample, steps, warmup, nthreads=0):
    """Times the CPU oracle (oracle/, a restatement of the reference algorithm: kind "port") on the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as o
    trajs, tid, state, control, t0, other = make_workload(B_sample)
    cache = {}
    ms = []
    for i in range(B_sample):
        j = int(tid[i])
        if j not in cache:
            cache[j] = o.Trajectory(**{k: trajs[k][j] for k in o.TRAJ_FIELDS})
        m = o.Mpc(o.MPC_COUPLED)
        m.set_trajectory(cache[j])
        m.set_state(state[i], control[i], other4=other[i])
        ms.append(m)
    cores = o.max_threads() if nthreads <= 0 else nthreads
    for k in range(SETTLE + warmup):
        o.batch_step(ms, t0 + 0.01 * k, rollout=True, nthreads=cores)
    t = time.perf_counter()
    for k in range(steps):
        o.batch_step(ms, t0 + 0.01 * (SETTLE + warmup + k), rollout=True, nthreads=cores)
    el = time.perf_counter() - t
    iters = float(np.mean([m.stats()["iter"] for m in ms]))
    return {"value": B_sample * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{B_sample} vehicles x {steps} closed-loop steps after {SETTLE} settling + {warmup} warm-up steps, oracle/liboracle.so (C++ -O3, std::thread over vehicles), mean ADMM iters {iters:.1f}",
            "ms_per_step": el / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    r = cpu_baseline(B, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": B, "note": "reference's Julia cannot run here (no julia in the image); CPU arm = oracle port on all host threads"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import pigeon.jl_b200 as p
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K, W = args.batch, args.steps, args.warmup
    trajs, tid, state, control, t0, other = make_workload(B, seed_shift=1000 * rank + args.seed_shift)
    mpc = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, B, trajectory_index=tid, device=local)
    stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream: the library launches on it, the events are recorded on it
    torch.cuda.set_stream(stream)
    mpc.set_stream(stream.cuda_stream)
    parts = mpc.set_pipeline_parts(args.parts)      # vehicle ranges on their own streams inside the fused calls (0 = automatic); results do not depend on it
    dt = 0.01

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (`value`) ----------------
    mpc.set_state(state, control, other)
    d_base = torch.tensor(t0, dtype=torch.float64, device=dev)
    d_t0 = d_base.clone()
    d_out = torch.zeros(3 * B, dtype=torch.float64, device=dev)
    rec_states, rec_controls = [], []     # closed-loop replay for the e2e arm
    kstep = [0]                           # step k runs at t0 + k*dt on every path (per-step calls, the simulate loop, the e2e replay)

    def dev_step(record):
        if record:
            q, u = mpc.get_state()
            rec_states.append(q); rec_controls.append(u)
        torch.add(d_base, kstep[0] * dt, out=d_t0)
        mpc.step_rollout_device(d_t0.data_ptr(), d_out.data_ptr(), dt)      # step + plant rollout (launched beside the QP solve)
        kstep[0] += 1

    def dev_steps(n):
        """n closed-loop steps with device-resident inputs: one call of the on-device `simulate` loop (model_predictive_control.jl:87-98;
        the pipeline parts run their steps independently), or n per-step calls with --loop steps."""
        if args.loop == "simulate":
            mpc.simulate_device_async(d_base.data_ptr(), dt, n, k0=kstep[0])
            kstep[0] += n
        else:
            for _ in range(n):
                dev_step(False)

    # pass 1 (untimed): record the closed-loop states of all SETTLE+W+K steps for the e2e replay
    for _ in range(SETTLE + W + K):
        dev_step(True)
    # pass 2 (timed): identical closed loop from the same initial condition (the path is deterministic)
    mpc.reset_solver(); mpc.reset_solved()
    mpc.set_state(state, control, other)
    kstep[0] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # nvidia-smi needs ~0.2 s to deliver its first sample: started before the settling pass, read over the whole loaded region
        time.sleep(0.3)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    dev_steps(SETTLE)
    c1.record(stream)
    barrier()
    cold_ms = c0.elapsed_time(c1)
    dev_steps(W)
    barrier()
    t_timed = time.perf_counter()      # clock samples from here on are the ones reported (the GPU is warm: settling + warm-up ran just before)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mpc.stage_ms(reset=True)
    ev0.record(stream)
    dev_steps(K)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = mpc.stage_ms(reset=True)["launches"]
    st = mpc.stats()
    # the same K steps as K per-step calls (pgn_step_rollout_device: the parts are joined at the end of every call), reported beside `value`
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(K):
        dev_step(False)
    ev3.record(stream)
    barrier()
    ms_calls = ev2.elapsed_time(ev3)
    clocks = sampler.stop(t_timed) if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * K / (ms_max * 1e-3)

    # per-stage device time of one profiled pass (CUDA events on the launch stream inside the library)
    mpc.set_profiling(1)                       # CUDA events around every stage on the launch stream; the kernels are unchanged
    for _ in range(min(K, 5)):
        dev_step(False)
    nprof = min(K, 5)
    stage = mpc.stage_ms(reset=True)
    mpc.set_profiling(2)                       # + in-kernel cycle counters of the ADMM phases (one extra step, not timed)
    mpc.admm_cycles(reset=True)
    dev_step(False)
    cyc = mpc.admm_cycles(reset=True)
    mpc.stage_ms(reset=True)
    mpc.set_profiling(0)
    admm_ms = stage["admm"] / nprof
    mean_iters = float(st["iters"].mean())

    # ---------------- end-to-end arm (`e2e`): host buffers through the C ABI, copies inside the timed region ----------------
    nrec = len(rec_states)
    pin_q = torch.empty((nrec, B, 6), dtype=torch.float64).pin_memory()
    pin_u = torch.empty((nrec, B, 3), dtype=torch.float64).pin_memory()
    pin_t = torch.empty((nrec, B), dtype=torch.float64).pin_memory()
    pin_o = torch.empty((nrec, B, 3), dtype=torch.float64).pin_memory()
    pin_q.numpy()[:] = np.stack(rec_states); pin_u.numpy()[:] = np.stack(rec_controls)
    pin_t.numpy()[:] = t0[None, :] + dt * np.arange(nrec)[:, None]
    mpc.reset_solver(); mpc.reset_solved()
    qn, un_, tn, on = pin_q.numpy(), pin_u.numpy(), pin_t.numpy(), pin_o.numpy()
    import ctypes as C
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
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - tw) * 1e3)
    t_e = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t_e.item()) * 1e-3)

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
        # (a) the batched step of this workload: pgn_set_state + pgn_step on the settled closed loop (states replayed cyclically)
        lo, hi = SETTLE + W, SETTLE + W + K
        latency = {"batched_step": dict(pct(lambda i: e2e_step(lo + i % (hi - lo))), batch=B, call="pgn_set_state + pgn_step (host buffers)")}
        # (b) the reference's deployment point: ONE vehicle through the callback entry point (one packed copy in, one CUDA graph, one copy out)
        one = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 1, trajectory_index=tid[:1], device=local)
        one.set_stream(stream.cuda_stream)
        one.set_state(state[:1], control[:1], other[:1])
        q1 = np.ascontiguousarray(np.stack(rec_states)[:, :1]); u1 = np.ascontiguousarray(np.stack(rec_controls)[:, :1])
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

    # ---------------- final gather of controls + statistics (NCCL over NVLink, outside the timed region) ----------------
    gathered = None
    if dist is not None:
        ctrl = d_out.clone()
        iters_t = torch.tensor(st["iters"], dtype=torch.int32, device=dev)
        gl = [torch.empty_like(ctrl) for _ in range(world)]
        gi = [torch.empty_like(iters_t) for _ in range(world)]
        dist.all_gather(gl, ctrl); dist.all_gather(gi, iters_t)
        gathered = {"controls": int(sum(g.numel() for g in gl)), "mean_iters_all_ranks": float(torch.cat(gi).float().mean().item())}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        n, m, Nk = mpc.n, mpc.m, mpc.n + mpc.m
        rec_len = 30 * 81 + 6 + 2 + 3 + 30
        # algorithmic HBM bytes of one ADMM launch: per QP the piece record in, warm iterates (x|z, y) in and out, solution x,y out, stats
        bytes_per_qp = 8 * (rec_len + 4 * Nk + n + m) + 40
        # algorithmic FP64 flops per QP, from the static programs of this QP: Ruiz (10 passes over A and diag P), factor
        # (3 flops per gather entry), range inverses, dense-tail sweep, then per iteration the forward/backward entries (2 flops each),
        # the dense tail mat-vec and ~15 flops per KKT row of vector updates; residual checks every 25 iterations
        nnzL, nnzA = mpc.nnzL, mpc.nnzA
        prog = mpc.qp_program
        Dm = prog["tail_dim"]
        flop_factor = 3 * prog["factor_entries"] + 3 * prog["inverse_entries"] + 3 * Dm * (Dm * (Dm + 1) // 2)
        flop_iter = 2 * (prog["l_slots"] + prog["backward_entries"]) + 2 * Dm * Dm + 15 * Nk
        flop_check = 4 * nnzA + 2 * n + 10 * Nk
        flop_qp = flop_factor + 10 * 2 * (nnzA + n) + mean_iters * flop_iter + (mean_iters / 25.0) * flop_check
        roof = {"kernel": "k_admm", "bound": "hbm", "achieved": B * bytes_per_qp / (admm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": peak_src, "traffic": None, "launch_ms": admm_ms,
                "share_of_step": admm_ms / max(1e-9, (stage["nodes"] + stage["linearize"] + stage["hji"] + stage["admm"] + stage["controls"] + stage["rollout"]) / nprof),
                "fp64": {"achieved_tflops": B * flop_qp / (admm_ms * 1e-3) / 1e12, "nominal_peak_tflops": 37.0, "flop_per_qp": flop_qp},
                "note": "one QP per CTA; the solve is a chain of dependent sparse triangular solves in shared memory: latency/occupancy-bound, neither HBM- nor tensor-bound (SURVEY.md 8d)",
                "limiter": {"what": "dependent-instruction latency of the slowest warp in each barrier interval (about 4.7 cycles per instruction of a lone warp), then shared-memory wavefronts (every 8-byte load of a warp is >= 2 wavefronts of one 128 B/clk pipe)",
                            "evidence": "tools/ubench/*.cu (B200 latencies), DESIGN.md 4.1 (table of measurements, rejected variants), profiles/r1b_admm_ncu_full.md (0.42 IPC per scheduler, stalls: barrier >> wait ~ short scoreboard)"}}
        roof["frac"] = roof["achieved"] / roof["peak"]
        # DRAM traffic of one launch from the committed `ncu --set full` capture of the same workload (read + written bytes)
        try:
            import re
            txt = open(os.path.join(ROOT, "profiles", "r1_admm_ncu_full.md")).read()
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = re.search(r"dram__bytes_read.sum`\) \| ([0-9.]+) \| (\w+)", txt); wr = re.search(r"dram__bytes_write.sum`\) \| ([0-9.]+) \| (\w+)", txt)
            wf = re.search(r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed`\) \| ([0-9.]+)", txt)
            roof["traffic"] = float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)]
            roof["traffic_source"] = "profiles/r1b_admm_ncu_full.md (B = 1024; below the algorithmic bytes because the records and iterates of 1024 vehicles stay in the 126 MB L2)"
            if wf:
                roof["smem"] = {"pct_of_peak_wavefronts": float(wf.group(1)), "source": "profiles/r1b_admm_ncu_full.md: the most loaded unit of the kernel is the shared-memory pipe"}
        except Exception:
            pass
        roof["fp64"]["frac"] = roof["fp64"]["achieved_tflops"] / 37.0
        ncpu = min(B, 256)
        cpu = cpu_baseline(ncpu, 3, 1) if not args.no_cpu else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "horizon_nodes": mpc.N, "qp": {"n": n, "m": m, "nnzA": nnzA, "nnzL": nnzL, "levels": mpc.n_levels, "program": prog},
                           "l2": "per-step working set (records + iterates, ~%.0f MB per GPU) is rewritten every step; B=1024 fits L2, the ADMM kernel is not HBM-bound" % (B * bytes_per_qp / 1e6),
                           "parallelism": f"batch sharded over {world} GPU(s), no hot-path collective",
                           "loop": ("one pgn_simulate_device call of K closed-loop steps (the reference's `simulate` loop on the device)" if args.loop == "simulate" else "K pgn_step_rollout_device calls"),
                           "pipeline_parts": parts},
                "per_step_calls": {"value": world * B * K / (ms_calls * 1e-3), "unit": UNIT + " (rank 0 time)", "ms_per_step": ms_calls / K,
                                   "call": "pgn_step_rollout_device x K, device-resident; the pipeline parts are joined at the end of every call"},
                "roofline": roof, "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (6 + 3 + 1) * 8, "d2h_bytes_per_step": B * 3 * 8},
                "gpu_launches": int(launches), "clocks": clocks,
                "admm": {"mean_iters": mean_iters, "p50_iters": float(np.median(st["iters"])), "p99_iters": float(np.percentile(st["iters"], 99)), "max_iters": int(st["iters"].max()),
                         "pct_not_solved": float((st["status"] != 1).mean() * 100)},
                "stage_ms_per_step": {k: stage[k] / nprof for k in ("nodes", "linearize", "hji", "admm", "controls", "rollout")},
                "admm_phase_share": {k: v / max(1.0, sum(cyc.values())) for k, v in cyc.items()},
                "latency": latency, "gather": gathered,
                "cold_start": {"steps": SETTLE, "value": B * SETTLE / (cold_ms * 1e-3), "unit": UNIT + " (rank 0, device-resident)", "ms_per_step": cold_ms / SETTLE,
                               "note": "first %d closed-loop steps from the perturbed cold start; a few QPs per step run to thousands of iterations (max_iter 4000) and one QP occupies one SM, so single stragglers set the launch time" % SETTLE}}
        print(json.dumps(line), flush=True)
    mpc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024, help="vehicles per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--parts", type=int, default=0, help="pipeline parts of the fused calls (0 = automatic, 1 = off)")
    ap.add_argument("--loop", default="simulate", choices=["simulate", "steps"], help="timed region of `value`: one on-device simulate call of K steps, or K per-step calls")
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
