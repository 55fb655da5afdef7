"""Host-side mirror of the reference's MPC API for the batched B200 engine.

Names, argument meaning and call order follow StanfordASL/Pigeon.jl (the Julia host shim julia/PigeonB200.jl has the same
surface; Julia is absent from this image so the tests drive this Python mirror through the same C ABI):

    X1()                                         src/vehicles.jl:1-59
    CoupledControlParams / DecoupledControlParams src/coupled_lat_long.jl:23-40, src/decoupled_lat_long.jl:18-30
    TrajectoryTube, straight_trajectory          src/trajectories.jl:8-44, 96-105; TrajectoryTube.from_path: src/ros_integration.jl:13-16
    HJICache, placeholder_HJICache               src/HJI_computation.jl:26-57
    Batched{Coupled,Decoupled}TrajectoryTrackingMPC(vehicle, trajectories; control_params, N_short, N_long, dt_short, dt_long,
        use_correction_step)                     src/coupled_lat_long.jl:42-60, src/decoupled_lat_long.jl:32-50
    compute_time_steps(mpc, t0), compute_linearization_nodes(mpc), update_QP(mpc), solve(mpc), get_next_control(mpc)
                                                 src/model_predictive_control.jl:70-78   (`!` dropped: not valid in Python)
    simulate(mpc, q0, u0, dt)                    src/model_predictive_control.jl:80-100
Every field of the reference's mutable struct that the ROS callback writes (current_state, current_control, other_car_state,
time_offset, trajectory, HJI_cache, solved) is a batched property here: arrays are vehicle-major, shape (B, k).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import PGN_COUPLED, PGN_DECOUPLED, PigeonError, check, dptr, f64

VP_NAMES = ["L", "a", "b", "h", "G", "m", "Izz", "mu", "Caf", "Car", "Cd0", "Cd1", "Cd2", "fwd_frac", "rwd_frac", "fwb_frac", "rwb_frac", "Fx_max",
            "Fx_min", "Px_max", "delta_max", "kappa_max", "inv_fiala_corrected"]
CP_NAMES = ["V_min", "V_max", "k_V", "k_s", "ddelta_max", "Q_ds", "Q_dpsi", "Q_e", "W_beta", "W_r", "W_HJI", "N_HJI", "R_delta", "R_ddelta", "R_Fx", "R_dFx"]
TRAJ_FIELDS = ["t", "s", "V", "A", "E", "N", "psi", "kappa", "theta", "phi", "edge_L", "edge_R"]


def X1():
    """Vehicle parameter dict of the X1 test vehicle (reference X1(): Dict{Symbol,Float64})."""
    v = np.zeros(23)
    check(_lib.load().pgn_x1_vehicle_params(dptr(v)))
    return dict(zip(VP_NAMES, (float(x) for x in v)))


def _control_params(kind, kw):
    c = np.zeros(16)
    check(_lib.load().pgn_default_control_params(kind, dptr(c)))
    d = dict(zip(CP_NAMES, (float(x) for x in c)))
    for k, val in kw.items():
        if k not in d:
            raise TypeError(f"unknown control parameter {k!r}")
        d[k] = float(val)
    return d


def CoupledControlParams(**kw):
    return _control_params(PGN_COUPLED, kw)


def DecoupledControlParams(**kw):
    return _control_params(PGN_DECOUPLED, kw)


class TrajectoryTube:
    def __init__(self, t, s, V, A, E, N, psi, kappa, theta=None, phi=None, edge_L=None, edge_R=None):
        n = len(t)
        z = np.zeros(n)
        vals = [t, s, V, A, E, N, psi, kappa, z if theta is None else theta, z if phi is None else phi,
                np.full(n, 4.0) if edge_L is None else edge_L, np.full(n, -4.0) if edge_R is None else edge_R]
        for name, val in zip(TRAJ_FIELDS, vals):
            a = np.ascontiguousarray(val, dtype=np.float64)
            if a.shape != (n,):
                raise ValueError("all TrajectoryTube fields must have the same length")   # the reference's @assert
            setattr(self, name, a)

    def __len__(self):
        return len(self.t)

    @classmethod
    def from_path(cls, p):
        """TrajectoryTube(p::path): t = invcumtrapz(Ux_des, s), phi = 0 (src/ros_integration.jl:13-16, src/math.jl:2). `p` maps the
        .world / path-message keys to arrays."""
        s, V = np.asarray(p["s_m"], float), np.asarray(p["UxDes_mps"], float)
        t = np.concatenate([[0.0], np.cumsum(2 * np.diff(s) / (V[:-1] + V[1:]))])
        return cls(t, s, V, p["AxDes_mps2"], p["posE_m"], p["posN_m"], p["psi_rad"], p["k_1pm"], p["grade_rad"], 0 * np.asarray(p["grade_rad"], float),
                   p["edgeL_m"], p["edgeR_m"])


def straight_trajectory(length, vel):
    return TrajectoryTube([0.0, length / vel], [0.0, length], [vel, vel], [0.0, 0.0], [0.0, 0.0], [0.0, length], [0.0, 0.0], [0.0, 0.0])


class HJICache:
    """grid_knots: 7 float32 vectors; V: float32 array indexed V[i1,...,i7]; gradV: float32 array (7, n1, ..., n7)."""

    def __init__(self, grid_knots, V, gradV):
        self.grid_knots = [np.ascontiguousarray(k, dtype=np.float32) for k in grid_knots]
        dims = tuple(len(k) for k in self.grid_knots)
        self.V = np.asarray(V, dtype=np.float32)
        self.gradV = np.asarray(gradV, dtype=np.float32)
        if self.V.shape != dims or self.gradV.shape != (7,) + dims:
            raise ValueError("HJICache: V must have shape dims and gradV shape (7,)+dims")


def placeholder_HJICache():
    knots = [np.array([-1000.0, 1000.0], dtype=np.float32) for _ in range(7)]
    return HJICache(knots, np.zeros((2,) * 7, np.float32), np.zeros((7,) + (2,) * 7, np.float32))


class BatchedTrajectoryTrackingMPC:
    """Batched TrajectoryTrackingMPC (reference src/model_predictive_control.jl:32-68): B independent controllers on one GPU."""

    def __init__(self, kind, vehicle, trajectories, batch, control_params=None, N_short=10, N_long=20, dt_short=0.01, dt_long=0.2,
                 use_correction_step=True, device=-1, trajectory_index=None, **solver_settings):
        lib = _lib.load()
        self._lib = lib
        self.kind = kind
        cfg = _lib.PgnConfig()
        check(lib.pgn_default_config(C.byref(cfg), kind))
        cfg.batch, cfg.N_short, cfg.N_long, cfg.dt_short, cfg.dt_long = int(batch), int(N_short), int(N_long), float(dt_short), float(dt_long)
        cfg.use_correction_step, cfg.device = int(bool(use_correction_step)), int(device)
        for k, v in solver_settings.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown setting {k!r}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self._h = C.c_void_p()
        check(lib.pgn_create(C.byref(cfg), C.byref(self._h)))
        self.B = int(batch)
        d = np.zeros(16, dtype=np.int32)
        check(lib.pgn_qp_dims(self._h, dptr(d)))
        self.N, self.nx, self.nu, self.n, self.m, self.nnzA, self.nnzL, self.n_levels = (int(x) for x in d[:8])
        self.qp_program = dict(l_slots=int(d[8]), solve_phases=int(d[9]), factor_entries=int(d[10]), inverse_entries=int(d[11]), tail_dim=int(d[12]),
                               backward_entries=int(d[13]), admm_smem_bytes=int(d[14]), admm_threads=int(d[15]) & 0xffff,
                               admm_variant="tmem" if (int(d[15]) >> 16) & 1 else "smem", admm_ctas_per_sm=(int(d[15]) >> 20) & 0xf)
        self.T = self.N - 1
        self.vehicle = dict(vehicle)
        self.control_params = dict(control_params) if control_params is not None else _control_params(kind, {})
        self._push_params()
        self.HJI_cache = None
        if trajectories is not None:
            self.set_trajectories(trajectories, trajectory_index)

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pgn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ----
    def _push_params(self):
        vp = np.array([self.vehicle.get(k, 0.0) for k in VP_NAMES], dtype=np.float64)
        cp = np.array([self.control_params[k] for k in CP_NAMES], dtype=np.float64)
        check(self._lib.pgn_set_vehicle_params(self._h, dptr(vp)))
        check(self._lib.pgn_set_control_params(self._h, dptr(cp)))
        self.u_normalization = np.array([self.vehicle["delta_max"], max(-self.vehicle["Fx_min"], self.vehicle["Fx_max"])])

    def set_control_params(self, control_params):
        self.control_params = dict(control_params)
        self._push_params()

    def set_stream(self, cuda_stream_ptr):
        check(self._lib.pgn_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # ---- mpc.trajectory ----
    def set_trajectories(self, trajectories, trajectory_index=None):
        if isinstance(trajectories, TrajectoryTube):
            trajectories = [trajectories]
        if isinstance(trajectories, dict):            # dict of (n_traj, n_nodes) arrays (synthetic generator)
            fields = [f64(trajectories[k]) for k in TRAJ_FIELDS]
        else:
            n = len(trajectories[0])
            if any(len(t) != n for t in trajectories):
                raise ValueError("all trajectories of a batch must have the same number of nodes")
            fields = [np.ascontiguousarray(np.stack([getattr(t, k) for t in trajectories])) for k in TRAJ_FIELDS]
        n_traj, n_nodes = fields[0].shape
        self.trajectory_end_time = float(fields[0][0, -1])         # mpc.trajectory.t[end] (of the first trajectory): simulate's default horizon
        arr = (C.c_void_p * 12)(*[f.ctypes.data for f in fields])
        check(self._lib.pgn_set_trajectories(self._h, n_traj, n_nodes, arr))
        self.n_traj = n_traj
        if trajectory_index is None:
            trajectory_index = np.arange(self.B) % n_traj
        self.assign_trajectories(trajectory_index)

    def assign_trajectories(self, trajectory_index):
        idx = np.ascontiguousarray(trajectory_index, dtype=np.int32)
        if idx.shape != (self.B,):
            raise ValueError("trajectory_index must have shape (B,)")
        check(self._lib.pgn_assign_trajectories(self._h, dptr(idx)))
        self.trajectory_index = idx

    # ---- mpc.HJI_cache ----
    def set_HJI_cache(self, cache):
        dims = np.array([len(k) for k in cache.grid_knots], dtype=np.int32)
        knots = np.concatenate(cache.grid_knots).astype(np.float32)
        V = np.ascontiguousarray(cache.V.ravel(order="F"))
        g = np.ascontiguousarray(cache.gradV.ravel(order="F"))       # (7, n1..n7) Fortran order => component fastest
        check(self._lib.pgn_set_hji_cache(self._h, dptr(dims), dptr(knots), dptr(V), dptr(g)))
        self.HJI_cache = cache

    # ---- mpc.current_state etc. ----
    def set_state(self, current_state=None, current_control=None, other_car_state=None, time_offset=None):
        B = self.B
        q = None if current_state is None else f64(np.broadcast_to(np.asarray(current_state, float), (B, 6)))
        u = None if current_control is None else f64(np.broadcast_to(np.asarray(current_control, float), (B, 3)))
        o = None if other_car_state is None else f64(np.broadcast_to(np.asarray(other_car_state, float), (B, 4)))
        t = None if time_offset is None else f64(np.broadcast_to(np.asarray(time_offset, float), (B,)))
        check(self._lib.pgn_set_state(self._h, dptr(q), dptr(u), dptr(o), dptr(t)))

    def get_state(self):
        q, u = np.zeros((self.B, 6)), np.zeros((self.B, 3))
        check(self._lib.pgn_get_state(self._h, dptr(q), dptr(u)))
        return q, u

    current_state = property(lambda self: self.get_state()[0], lambda self, v: self.set_state(current_state=v))
    current_control = property(lambda self: self.get_state()[1], lambda self, v: self.set_state(current_control=v))

    def reset_solved(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        check(self._lib.pgn_reset_solved(self._h, dptr(m)))

    def reset_solver(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        check(self._lib.pgn_reset_solver(self._h, dptr(m)))

    # ---- step API ----
    def set_guards(self, nan_fallback=False, pause_below_speed=0.0):
        """Per-vehicle guards of the reference's ROS callback (src/ros_integration.jl:84-87, 134-147); off by default."""
        check(self._lib.pgn_set_guards(self._h, int(bool(nan_fallback)), C.c_double(float(pause_below_speed))))

    def from_autobox(self, state, control, stamp, other_car=None):
        """from_autobox_callback (src/ros_integration.jl:48-151) for the batch in one call: returns (B, 5) = (delta, Fxf, Fxr, s_m, e_m)."""
        q, u = f64(state, (self.B, 6)), f64(control, (self.B, 3))
        o = None if other_car is None else f64(other_car, (self.B, 4))
        out = np.zeros((self.B, 5))
        check(self._lib.pgn_from_autobox(self._h, dptr(q), dptr(u), dptr(o), dptr(self._t0(stamp)), dptr(out)))
        return out

    def _t0(self, t0):
        return f64(np.broadcast_to(np.asarray(t0, float), (self.B,)))

    def compute_time_steps(self, t0):
        check(self._lib.pgn_compute_time_steps(self._h, dptr(self._t0(t0))))

    def compute_linearization_nodes(self):
        check(self._lib.pgn_compute_linearization_nodes(self._h))

    def update_QP(self):
        check(self._lib.pgn_update_qp(self._h))

    def solve(self):
        check(self._lib.pgn_solve(self._h))

    def get_next_control(self):
        out = np.zeros((self.B, 3))
        check(self._lib.pgn_get_next_control(self._h, dptr(out)))
        return out

    def step(self, t0):
        out = np.zeros((self.B, 3))
        check(self._lib.pgn_step(self._h, dptr(self._t0(t0)), dptr(out)))
        return out

    def step_submit(self, t0, current_state=None, current_control=None, other_car_state=None):
        """Pipelined stepping: set_state + step without the wait (up to 4 steps in flight); results come from step_collect, oldest first."""
        B = self.B
        q = None if current_state is None else f64(current_state, (B, 6))
        u = None if current_control is None else f64(current_control, (B, 3))
        o = None if other_car_state is None else f64(other_car_state, (B, 4))
        check(self._lib.pgn_step_submit(self._h, dptr(q), dptr(u), dptr(o), dptr(self._t0(t0))))

    def step_collect(self):
        out = np.zeros((self.B, 3))
        check(self._lib.pgn_step_collect(self._h, dptr(out)))
        return out

    def step_device(self, d_t0_ptr, d_out_ptr=None):
        check(self._lib.pgn_step_device(self._h, C.c_void_p(d_t0_ptr), C.c_void_p(d_out_ptr) if d_out_ptr else None))

    def step_rollout_device(self, d_t0_ptr, d_out_ptr=None, dt=0.01):
        """One iteration of the `simulate` loop on device-resident data: step, then plant rollout (launched beside the QP solve)."""
        check(self._lib.pgn_step_rollout_device(self._h, C.c_void_p(d_t0_ptr), C.c_void_p(d_out_ptr) if d_out_ptr else None, float(dt)))

    def rollout(self, dt=0.01):
        check(self._lib.pgn_rollout(self._h, float(dt)))

    def simulate_device(self, t0, dt, n_steps):
        check(self._lib.pgn_simulate(self._h, dptr(self._t0(t0)), float(dt), int(n_steps)))

    def simulate_device_async(self, d_t0_ptr, dt, n_steps, k0=0):
        """pgn_simulate with t0 resident on the device, enqueued on the handle's stream (no host synchronisation): steps k0 .. k0+n_steps-1 at t0 + k*dt."""
        check(self._lib.pgn_simulate_device(self._h, C.c_void_p(d_t0_ptr), float(dt), int(k0), int(n_steps)))

    def set_history(self, capacity, stride=1):
        """Record (on the device) every `stride`-th step of simulate_device / simulate_device_async: the reference's simulate returns
        qs, xs, us, ps per step (model_predictive_control.jl:84-99).  capacity = 0 switches the recorder off."""
        check(self._lib.pgn_set_history(self._h, int(capacity), int(stride) if capacity else 0))

    def history(self):
        """(qs, xs, us, ps) of the recorded steps: arrays of shape (n, B, 6), (n, B, nx), (n, B, 3), (n, B, 4)."""
        n = C.c_int32(0)
        check(self._lib.pgn_get_history(self._h, C.byref(n), None, None, None, None))
        n = int(n.value)
        qs, us, xs, ps = np.zeros((n, self.B, 6)), np.zeros((n, self.B, 3)), np.zeros((n, self.B, self.nx)), np.zeros((n, self.B, 4))
        if n:
            check(self._lib.pgn_get_history(self._h, C.byref(C.c_int32(0)), dptr(qs), dptr(us), dptr(xs), dptr(ps)))
        return qs, xs, us, ps

    def set_solve_cap(self, iters):
        """ADMM iterations of one QP per launch inside the simulate loops (0 = unlimited): a straggler QP then delays only its own vehicle;
        results do not depend on it."""
        check(self._lib.pgn_set_solve_cap(self._h, int(iters)))

    def set_pipeline_parts(self, parts):
        """Run the fused entry points as `parts` vehicle ranges on their own streams (0: automatic, 1: off); results do not depend on it."""
        check(self._lib.pgn_set_pipeline_parts(self._h, int(parts)))
        return self.pipeline_parts

    @property
    def pipeline_parts(self):
        n = C.c_int32(0)
        check(self._lib.pgn_get_pipeline_parts(self._h, C.byref(n)))
        return int(n.value)

    def synchronize(self):
        check(self._lib.pgn_synchronize(self._h))

    # ---- introspection ----
    def time_steps(self):
        ts, dt, pts = np.zeros((self.B, self.N)), np.zeros((self.B, self.T)), np.zeros((self.B, self.N))
        check(self._lib.pgn_get_time_steps(self._h, dptr(ts), dptr(dt), dptr(pts)))
        return ts, dt, pts

    def nodes(self):
        qs, us, ps = np.zeros((self.B, self.N, self.nx)), np.zeros((self.B, self.N, 2)), np.zeros((self.B, self.N, 4))
        check(self._lib.pgn_get_nodes(self._h, dptr(qs), dptr(us), dptr(ps)))
        return qs, us, ps

    def set_nodes(self, qs, us, ps):
        check(self._lib.pgn_set_nodes(self._h, dptr(f64(qs, (self.B, self.N, self.nx))), dptr(f64(us, (self.B, self.N, 2))), dptr(f64(ps, (self.B, self.N, 4)))))

    def qp_data(self):
        B, T, nx, nu = self.B, self.T, self.nx, self.nu
        d = dict(A=np.zeros((B, T, nx, nx)), B0=np.zeros((B, T, nx, nu)), Bf=np.zeros((B, T, nx, nu)), c=np.zeros((B, T, nx)), H=np.zeros((B, T, 4, 2)),
                 G=np.zeros((B, T, 4)), dmin=np.zeros((B, T)), dmax=np.zeros((B, T)), fxmax=np.zeros((B, T)), hji=np.zeros((B, 3)))
        check(self._lib.pgn_get_qp_data(self._h, *[dptr(d[k]) for k in ("A", "B0", "Bf", "c", "H", "G", "dmin", "dmax", "fxmax", "hji")]))
        return d

    def solution(self):
        x, y = np.zeros((self.B, self.n)), np.zeros((self.B, self.m))
        check(self._lib.pgn_get_solution(self._h, dptr(x), dptr(y)))
        return x, y

    def stats(self):
        B = self.B
        it, st, ru = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
        pr, du, rho = np.zeros(B), np.zeros(B), np.zeros(B)
        check(self._lib.pgn_get_stats(self._h, dptr(it), dptr(st), dptr(pr), dptr(du), dptr(rho), dptr(ru)))
        return dict(iters=it, status=st, pri_res=pr, dua_res=du, rho=rho, rho_updates=ru)

    def hji_lookup(self, x):
        x = f64(np.atleast_2d(x))
        M = x.shape[0]
        V, g = np.zeros(M), np.zeros((M, 7))
        check(self._lib.pgn_hji_lookup(self._h, M, dptr(x), dptr(V), dptr(g)))
        return V, g

    def hji_values(self):
        """(V, gradV) of the last step's HJIRelativeState(current_state, other_car_state): what the callback logs (ros_integration.jl:57-58)."""
        V, g = np.zeros(self.B), np.zeros((self.B, 7))
        check(self._lib.pgn_get_hji_values(self._h, dptr(V), dptr(g)))
        return V, g

    def optimal_control(self, relative_state, gradV):
        """optimal_control(dynamics, relative_state, gradV) (src/HJI_computation.jl:133-158) -> (M, 2) array of (delta, Fx)."""
        x, g = f64(np.atleast_2d(relative_state)), f64(np.atleast_2d(gradV))
        if x.shape != g.shape or x.shape[1] != 7:
            raise ValueError("relative_state and gradV must both have shape (M, 7)")
        out = np.zeros((x.shape[0], 2))
        check(self._lib.pgn_hji_optimal_control(self._h, x.shape[0], dptr(x), dptr(g), dptr(out)))
        return out

    def set_path_search_window(self, half_width):
        """Windowed closest-segment search of path_coordinates (src/trajectories.jl:71-80); 0 = the reference's full scan."""
        check(self._lib.pgn_set_path_search_window(self._h, int(half_width)))

    def set_hji_policy(self, on):
        """use_HJI_policy[] of the callback (src/ros_integration.jl:47,115-118): V <= HJI_eps => the "hammer" control."""
        check(self._lib.pgn_set_hji_policy(self._h, int(bool(on))))

    def set_hji_lookup_order(self, mode):
        """Stand-alone lookups: 1 visit the queries in grid-cell order (counting sort; corners shared through L2), 2 the same with one TMA-staged
        tile of corners per block of cells, 0 input order, -1 automatic."""
        check(self._lib.pgn_set_hji_lookup_order(self._h, int(mode)))

    def hji_lookup_device(self, M, d_x, d_V, d_g):
        check(self._lib.pgn_hji_lookup_device(self._h, int(M), C.c_void_p(d_x), C.c_void_p(d_V), C.c_void_p(d_g)))

    def device_controls_ptr(self):
        p = C.c_void_p()
        check(self._lib.pgn_device_controls(self._h, C.byref(p)))
        return p.value

    def device_stats_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        check(self._lib.pgn_device_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_profiling(self, on):
        check(self._lib.pgn_set_profiling(self._h, int(on)))

    def stage_ms(self, reset=True):
        out = np.zeros(8)
        check(self._lib.pgn_get_stage_ms(self._h, dptr(out), int(reset)))
        return dict(nodes=out[0], linearize=out[1], hji=out[2], admm=out[3], controls=out[4], rollout=out[5], launches=int(out[6]), catchup_rounds=int(out[7]))


    def admm_trace(self, reset=True, max_entries=1 << 19):
        """Records of profiling mode 3 (set_profiling(3)): an (n, 3) uint64 array of (start ns, end ns, meta) — meta = part | QPs solved << 8 | SM << 32
        for an ADMM CTA, part | stage << 8 | 1 << 63 (end = 0) for a time stamp between the stages of a round."""
        buf = np.zeros(3 * max_entries, dtype=np.uint64)
        n = C.c_int32(0)
        check(self._lib.pgn_get_admm_trace(self._h, dptr(buf), int(max_entries), C.byref(n), int(reset)))
        return buf[:3 * n.value].reshape(n.value, 3)

    def admm_cycles(self, reset=True):
        out = np.zeros(512)
        check(self._lib.pgn_get_admm_cycles(self._h, dptr(out), int(reset)))
        self.level_cycles = out[16:].copy()
        return dict(zip(["gather", "ruiz", "factor", "solve", "update", "check", "store", "ticket"], out[:8]))


def BatchedCoupledTrajectoryTrackingMPC(vehicle, trajectories, batch, control_params=None, **kw):
    return BatchedTrajectoryTrackingMPC(PGN_COUPLED, vehicle, trajectories, batch, control_params=control_params, **kw)


def BatchedDecoupledTrajectoryTrackingMPC(vehicle, trajectories, batch, control_params=None, **kw):
    return BatchedTrajectoryTrackingMPC(PGN_DECOUPLED, vehicle, trajectories, batch, control_params=control_params, **kw)


def comm_init_all(mpcs):
    """One process, one controller batch per GPU: form the NCCL communicator of the final gather (pgn_comm_init_all)."""
    arr = (C.c_void_p * len(mpcs))(*[m._h.value for m in mpcs])
    check(_lib.load().pgn_comm_init_all(arr, len(mpcs)))


def gather_all(mpcs):
    """Final gather over NCCL of the last step's controls and per-QP statistics of all batches (rank-major)."""
    n, B = len(mpcs), mpcs[0].B
    arr = (C.c_void_p * n)(*[m._h.value for m in mpcs])
    c, it, st = np.zeros((n * B, 3)), np.zeros(n * B, np.int32), np.zeros(n * B, np.int32)
    check(_lib.load().pgn_gather_all(arr, n, dptr(c), dptr(it), dptr(st)))
    return c, it, st


# the reference's generic functions (model_predictive_control.jl:70-78)
def compute_time_steps(mpc, t0):
    mpc.compute_time_steps(t0)


def compute_linearization_nodes(mpc):
    mpc.compute_linearization_nodes()


def update_QP(mpc):
    mpc.update_QP()


def solve(mpc):
    mpc.solve()


def get_next_control(mpc):
    return mpc.get_next_control()


def simulate(mpc, q0, u0, dt=0.01, t0=0.0, n_steps=None, T_end=None, record=True, stride=1, on_device=True):
    """simulate(mpc, q0, u0, dt) (model_predictive_control.jl:80-100) for the whole batch: `for t in 0:dt:mpc.trajectory.t[end]`.
    T_end defaults to the end time of the (first) trajectory, as in the reference.  record=True returns the reference's
    (qs, xs, us, ps) — states / controls before each step, mpc.qs[1], mpc.ps[1] — every `stride`-th step, recorded ON THE DEVICE
    (pgn_set_history) while the loop runs as one pgn_simulate call; record=False returns only the final (state, control).
    on_device=False drives the five step calls from the host (one round trip per step; qs, us only)."""
    mpc.set_state(current_state=q0, current_control=u0)
    if n_steps is None:
        if T_end is None:
            T_end = mpc.trajectory_end_time
        n_steps = int(math.floor((T_end - 0.0) / dt + 1e-9)) + 1
    if on_device:
        if record:
            mpc.set_history((n_steps + stride - 1) // stride, stride)
        mpc.simulate_device(t0, dt, n_steps)
        if not record:
            return mpc.get_state()
        out = mpc.history()
        mpc.set_history(0)
        return out
    qs, us = [], []
    t0 = np.broadcast_to(np.asarray(t0, float), (mpc.B,)).copy()
    for k in range(n_steps):
        q, u = mpc.get_state()
        qs.append(q); us.append(u)
        mpc.compute_time_steps(t0 + k * dt)
        mpc.compute_linearization_nodes()
        mpc.update_QP()
        mpc.solve()
        mpc.get_next_control()      # leaves the new control in the device buffer consumed by rollout
        mpc.rollout(dt)
    return np.stack(qs), np.stack(us)
