"""ctypes binding of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
PARITY UNPINNED: the reference has no golden vectors (test/runtests.jl:5); see oracle/README.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MODEL_BICYCLE, MODEL_TRACKING, MODEL_LATERAL = 0, 1, 2
MPC_COUPLED, MPC_DECOUPLED = 0, 1

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
fp = C.POINTER(C.c_float)


NATIVE_FLAGS = "-O3 -march=native -std=c++17 -pthread (FMA contraction on)"
CHECKER_FLAGS = "-O3 -std=c++17 -pthread -ffp-contract=off (portable)"
build_flags = CHECKER_FLAGS


def use_native_build():
    """bench.py's timed CPU arm only: build oracle/liboracle_native.so on THIS machine with -O3 -march=native (oracle/Makefile) and bind it.
    Must be called before the first use of the library in the process.  Falls back to the portable checker build when make fails."""
    global _LIB_PATH, build_flags
    if _lib is not None:
        return build_flags
    native = os.path.join(_HERE, "liboracle_native.so")
    try:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_native.so"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _LIB_PATH, build_flags = native, NATIVE_FLAGS
    except Exception:
        pass
    return build_flags


def build(force=False):
    if _LIB_PATH.endswith("liboracle_native.so"):
        return _LIB_PATH
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], env=env, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_fiala.restype = C.c_double
        _lib.orc_fiala.argtypes = [C.c_double] * 5
        _lib.orc_invfiala.restype = C.c_double
        _lib.orc_invfiala.argtypes = [C.c_double] * 3
        _lib.orc_adiff.restype = C.c_double
        _lib.orc_adiff.argtypes = [C.c_double] * 2
        for name in ("orc_traj_create", "orc_hji_create", "orc_hji_placeholder", "orc_osqp_create", "orc_mpc_create"):
            getattr(_lib, name).restype = C.c_void_p
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def _out(*shape):
    a = np.zeros(shape, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def x1():
    a, p = _out(23)
    lib().orc_x1(p)
    return a


VP_NAMES = ["L", "a", "b", "h", "G", "m", "Izz", "mu", "Caf", "Car", "Cd0", "Cd1", "Cd2", "fwd_frac", "rwd_frac", "fwb_frac",
            "rwb_frac", "Fx_max", "Fx_min", "Px_max", "delta_max", "kappa_max", "inv_fiala_corrected"]
CP_NAMES = ["V_min", "V_max", "k_V", "k_s", "delta_dot_max", "Q_ds", "Q_dpsi", "Q_e", "W_beta", "W_r", "W_HJI", "N_HJI", "R_delta",
            "R_ddelta", "R_Fx", "R_dFx"]
ST_NAMES = ["rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf", "eps_dual_inf", "max_iter", "scaling",
            "check_termination", "adaptive_rho", "adaptive_rho_interval", "adaptive_rho_tolerance", "warm_start"]


def control_params_default(kind):
    a, p = _out(16)
    lib().orc_control_params_default(kind, p)
    return a


def osqp_settings_default(**kw):
    a, p = _out(14)
    lib().orc_osqp_settings_default(p)
    for k, v in kw.items():
        a[ST_NAMES.index(k)] = v
    return a


def vehicle_model(kind, vp, q, u2, p4):
    nx = 4 if kind == MODEL_LATERAL else 6
    o, op = _out(nx)
    lib().orc_vehicle_model(kind, _d(vp)[1], _d(q)[1], _d(u2)[1], _d(p4)[1], op)
    return o


def lateral_tire_forces(vp, q6, u3, num_iters=3):
    o, op = _out(2)
    lib().orc_lateral_tire_forces(_d(vp)[1], _d(q6)[1], _d(u3)[1], num_iters, op)
    return o


def stable_limits(vp, Ux, Fxf, Fxr):
    o, op = _out(14)
    lib().orc_stable_limits(_d(vp)[1], C.c_double(Ux), C.c_double(Fxf), C.c_double(Fxr), op)
    return dict(delta_min=o[0], delta_max=o[1], H=o[2:10].reshape(4, 2).copy(), G=o[10:14].copy())


def steady_state(vp, V, A_tan, kappa, num_iters=4, r=None, beta0=0.0, delta0=0.0, Fyf0=0.0):
    if r is None:
        r = V * kappa
    o, op = _out(8)
    lib().orc_steady_state(_d(vp)[1], C.c_double(V), C.c_double(A_tan), C.c_double(kappa), num_iters, C.c_double(r),
                           C.c_double(beta0), C.c_double(delta0), C.c_double(Fyf0), op)
    return dict(zip(["beta", "Ux", "Uy", "r", "A", "delta", "Fxf", "Fxr"], o))


def flow(kind, vp, x, dt, up0, upf=None, nsub=10):
    nx = 4 if kind == MODEL_LATERAL else 6
    if upf is None:
        upf = up0
    o, op = _out(nx)
    lib().orc_flow(kind, _d(vp)[1], _d(x)[1], C.c_double(dt), _d(up0)[1], _d(upf)[1], nsub, op)
    return o


def _lin(fn, kind, vp, x, dt, up0, upf, ramp, nk):
    nx = 4 if kind == MODEL_LATERAL else 6
    if upf is None:
        upf = up0
    A, Ap = _out(nx, nx)
    B0, B0p = _out(nx, nk)
    Bf, Bfp = _out(nx, nk)
    c, cp = _out(nx)
    fn(kind, _d(vp)[1], _d(x)[1], C.c_double(dt), _d(up0)[1], _d(upf)[1], int(ramp), nk, Ap, B0p, Bfp, cp)
    return A, B0, Bf, c


def linearize_flow(kind, vp, x, dt, up0, upf=None, ramp=False, nk=2):
    return _lin(lib().orc_linearize_flow, kind, vp, x, dt, up0, upf, ramp, nk)


def linearize_exact(kind, vp, x, dt, up0, upf=None, ramp=False, nk=1):
    return _lin(lib().orc_linearize_exact, kind, vp, x, dt, up0, upf, ramp, nk)


def linearize_continuous(kind, vp, x, up):
    nx = 4 if kind == MODEL_LATERAL else 6
    A, Ap = _out(nx, nx)
    B, Bp = _out(nx, 6)
    f, fp_ = _out(nx)
    lib().orc_linearize_continuous(kind, _d(vp)[1], _d(x)[1], _d(up)[1], Ap, Bp, fp_)
    return A, B, f


def expm(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    E, Ep = _out(*A.shape)
    lib().orc_expm(A.shape[0], _d(A)[1], Ep)
    return E


TRAJ_FIELDS = ["t", "s", "V", "A", "E", "N", "psi", "kappa", "theta", "phi", "edge_L", "edge_R"]


class Trajectory:
    def __init__(self, **f):
        n = len(f["t"])
        arrs = []
        for k in TRAJ_FIELDS:
            if k not in f:
                f[k] = np.zeros(n) if k in ("theta", "phi") else (np.full(n, 4.0) if k == "edge_L" else np.full(n, -4.0))
            arrs.append(np.ascontiguousarray(f[k], dtype=np.float64))
        self.fields = dict(zip(TRAJ_FIELDS, arrs))
        self.n = n
        self.h = C.c_void_p(lib().orc_traj_create(n, *[a.ctypes.data_as(dp) for a in arrs]))

    def __del__(self):
        try:
            lib().orc_traj_free(self.h)
        except Exception:
            pass

    def at_time(self, t):
        o, op = _out(12)
        lib().orc_traj_at_time(self.h, C.c_double(t), op)
        return dict(zip(TRAJ_FIELDS, o))

    def at_s(self, s):
        o, op = _out(12)
        lib().orc_traj_at_s(self.h, C.c_double(s), op)
        return dict(zip(TRAJ_FIELDS, o))

    def path_coordinates(self, x, y):
        o, op = _out(3)
        lib().orc_traj_path_coordinates(self.h, C.c_double(x), C.c_double(y), op)
        return tuple(o)


def invcumtrapz(y, x):
    o, op = _out(len(x))
    lib().orc_invcumtrapz(len(x), _d(y)[1], _d(x)[1], op)
    return o


class HjiCache:
    def __init__(self, knots=None, V=None, gradV=None):
        """knots: list of 7 float32 arrays; V: float32 array of shape dims (Fortran/column-major semantic: V[i1,...,i7]);
        gradV: float32 array of shape (7,)+dims."""
        if knots is None:
            self.h = C.c_void_p(lib().orc_hji_placeholder())
            return
        dims = np.array([len(k) for k in knots], dtype=np.int32)
        kc = np.concatenate([np.asarray(k, dtype=np.float32) for k in knots])
        Vf = np.asfortranarray(V, dtype=np.float32).ravel(order="F")
        gf = np.asfortranarray(gradV, dtype=np.float32).ravel(order="F")
        self.h = C.c_void_p(lib().orc_hji_create(dims.ctypes.data_as(ip), kc.ctypes.data_as(fp), Vf.ctypes.data_as(fp), gf.ctypes.data_as(fp)))

    def __del__(self):
        try:
            lib().orc_hji_free(self.h)
        except Exception:
            pass

    def lookup(self, x7):
        x7 = np.ascontiguousarray(np.atleast_2d(x7), dtype=np.float64)
        M = x7.shape[0]
        V, Vp = _out(M)
        g, gp = _out(M, 7)
        lib().orc_hji_lookup(self.h, M, x7.ctypes.data_as(dp), Vp, gp)
        return V, g


def hji_relative_state(us6, them4):
    o, op = _out(7)
    lib().orc_hji_relative_state(_d(us6)[1], _d(them4)[1], op)
    return o


def optimal_disturbance(vp, x7, gV7):
    o, op = _out(2)
    lib().orc_optimal_disturbance(_d(vp)[1], _d(x7)[1], _d(gV7)[1], op)
    return o


def optimal_control(vp, x7, gV7):
    """optimal_control (HJI_computation.jl:133-158), uMode = :max, N = 50 -> (delta, Fx)"""
    o, op = _out(2)
    lib().orc_optimal_control(_d(vp)[1], _d(x7)[1], _d(gV7)[1], op)
    return o


def reachability_constraint(vp, cache, x7, eps, uR2):
    M, Mp = _out(2)
    b = C.c_double(0)
    lib().orc_reachability_constraint(_d(vp)[1], cache.h, _d(x7)[1], C.c_double(eps), _d(uR2)[1], Mp, C.byref(b))
    return M, b.value


class Osqp:
    """Generic QP: min 1/2 x'Px + q'x s.t. l <= Ax <= u; P (upper triangle) and A as scipy.sparse CSC."""

    def __init__(self, P, q, A, l, u, settings=None):
        import scipy.sparse as sp
        P = sp.triu(sp.csc_matrix(P), format="csc")
        A = sp.csc_matrix(A)
        P.sort_indices()
        A.sort_indices()
        self.n, self.m = A.shape[1], A.shape[0]
        self.P, self.A = P, A
        st = osqp_settings_default() if settings is None else np.asarray(settings, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        Pp, Pi, Ap, Ai = i32(P.indptr), i32(P.indices), i32(A.indptr), i32(A.indices)
        self.h = C.c_void_p(lib().orc_osqp_create(self.n, self.m, Pp.ctypes.data_as(ip), Pi.ctypes.data_as(ip), _d(P.data)[1], _d(q)[1],
                                                   Ap.ctypes.data_as(ip), Ai.ctypes.data_as(ip), _d(A.data)[1], _d(l)[1], _d(u)[1], _d(st)[1]))

    def __del__(self):
        try:
            lib().orc_osqp_free(self.h)
        except Exception:
            pass

    def update(self, Px=None, Ax=None, q=None, l=None, u=None):
        f = lambda a: None if a is None else _d(a)[1]
        lib().orc_osqp_update(self.h, f(Px), f(Ax), f(q), f(l), f(u))

    def warm_start(self, x, y):
        lib().orc_osqp_warm_start(self.h, _d(x)[1], _d(y)[1])

    def cold_start(self):
        lib().orc_osqp_cold_start(self.h)

    def solve(self):
        x, xp = _out(self.n)
        y, yp = _out(self.m)
        info, infop = _out(7)
        lib().orc_osqp_solve(self.h, xp, yp, infop)
        return x, y, dict(status=int(info[0]), iter=int(info[1]), pri_res=info[2], dua_res=info[3], rho=info[4], rho_updates=int(info[5]),
                          n_factor=int(info[6]))


class Mpc:
    def __init__(self, kind, vp=None, cp=None, N_short=10, N_long=20, dt_short=0.01, dt_long=0.2, use_correction_step=True, settings=None):
        self.kind = kind
        self.vp = x1() if vp is None else np.asarray(vp, dtype=np.float64)
        self.cp = control_params_default(kind) if cp is None else np.asarray(cp, dtype=np.float64)
        st = osqp_settings_default() if settings is None else np.asarray(settings, dtype=np.float64)
        self.h = C.c_void_p(lib().orc_mpc_create(kind, _d(self.vp)[1], _d(self.cp)[1], N_short, N_long, C.c_double(dt_short), C.c_double(dt_long),
                                                  int(use_correction_step), _d(st)[1]))
        d = np.zeros(6, dtype=np.int32)
        lib().orc_mpc_dims(self.h, d.ctypes.data_as(ip))
        self.N, self.nx, self.nu, self.n, self.m, self.nnzA = [int(v) for v in d]
        self.T = self.N - 1
        self._keep = []

    def __del__(self):
        try:
            lib().orc_mpc_free(self.h)
        except Exception:
            pass

    def set_trajectory(self, traj):
        lib().orc_mpc_set_trajectory(self.h, traj.h)

    def set_hji(self, cache, eps=0.05):
        lib().orc_mpc_set_hji(self.h, cache.h, C.c_double(eps))

    def from_autobox(self, q6, u3, stamp, other4=None, pause_speed=0.0, nan_fallback=False):
        """from_autobox_callback (ros_integration.jl:48-151): returns (published, out5 = (delta, Fxf, Fxr, s, e))"""
        o, op = _out(5)
        ot = None if other4 is None else _d(other4)[1]
        r = lib().orc_mpc_from_autobox(self.h, _d(q6)[1], _d(u3)[1], ot, C.c_double(stamp), C.c_double(pause_speed), C.c_int(int(nan_fallback)), op)
        return bool(r), o

    def set_hji_policy(self, on):
        """use_HJI_policy[] of the callback (ros_integration.jl:47,115-118)"""
        lib().orc_mpc_set_hji_policy(self.h, C.c_int(int(bool(on))))

    def hji_values(self):
        V = C.c_double(0)
        g, gp = _out(7)
        lib().orc_mpc_get_hji_values(self.h, C.byref(V), gp)
        return V.value, g

    def set_state(self, q6, u3, other4=None, time_offset=float("nan")):
        o = None if other4 is None else _d(other4)[1]
        lib().orc_mpc_set_state(self.h, _d(q6)[1], _d(u3)[1], o, C.c_double(time_offset))

    def get_state(self):
        q, qp = _out(6)
        u, up = _out(3)
        lib().orc_mpc_get_state(self.h, qp, up)
        return q, u

    def set_solved(self, flag):
        lib().orc_mpc_set_solved(self.h, int(flag))

    def reset_solver(self):
        lib().orc_mpc_reset_solver(self.h)

    def compute_time_steps(self, t0):
        lib().orc_mpc_compute_time_steps(self.h, C.c_double(t0))

    def time_steps(self):
        ts, tsp = _out(self.N)
        dt, dtp = _out(self.T)
        pts, ptsp = _out(self.N)
        lib().orc_mpc_get_time_steps(self.h, tsp, dtp, ptsp)
        return ts, dt, pts

    def compute_linearization_nodes(self):
        lib().orc_mpc_compute_linearization_nodes(self.h)

    def nodes(self):
        qs, qsp = _out(self.N, self.nx)
        us, usp = _out(self.N, 2)
        ps, psp = _out(self.N, 4)
        lib().orc_mpc_get_nodes(self.h, qsp, usp, psp)
        return qs, us, ps

    def set_nodes(self, qs, us, ps):
        lib().orc_mpc_set_nodes(self.h, _d(qs)[1], _d(us)[1], _d(ps)[1])

    def update_qp(self):
        lib().orc_mpc_update_qp(self.h)

    def qp_pieces(self):
        T, nx, nu = self.T, self.nx, self.nu
        A, Ap = _out(T, nx, nx)
        B0, B0p = _out(T, nx, nu)
        Bf, Bfp = _out(T, nx, nu)
        c, cp = _out(T, nx)
        H, Hp = _out(T, 4, 2)
        G, Gp = _out(T, 4)
        dmin, dminp = _out(T)
        dmax, dmaxp = _out(T)
        fxmax, fxmaxp = _out(T)
        hji, hjip = _out(3)
        lib().orc_mpc_get_qp_pieces(self.h, Ap, B0p, Bfp, cp, Hp, Gp, dminp, dmaxp, fxmaxp, hjip)
        return dict(A=A, B0=B0, Bf=Bf, c=c, H=H, G=G, dmin=dmin, dmax=dmax, fxmax=fxmax, hji=hji)

    def qp(self):
        import scipy.sparse as sp
        Pd, Pdp = _out(self.n)
        q, qp = _out(self.n)
        l, lp = _out(self.m)
        u, up = _out(self.m)
        Ax, Axp = _out(self.nnzA)
        Ap = np.zeros(self.n + 1, dtype=np.int32)
        Ai = np.zeros(self.nnzA, dtype=np.int32)
        lib().orc_mpc_get_qp(self.h, Pdp, qp, Ap.ctypes.data_as(ip), Ai.ctypes.data_as(ip), Axp, lp, up)
        A = sp.csc_matrix((Ax, Ai, Ap), shape=(self.m, self.n))
        return dict(Pdiag=Pd, q=q, A=A, l=l, u=u)

    def solve(self):
        return int(lib().orc_mpc_solve(self.h))

    def solution(self):
        x, xp = _out(self.n)
        y, yp = _out(self.m)
        lib().orc_mpc_get_solution(self.h, xp, yp)
        return x, y

    def stats(self):
        s, sp_ = _out(6)
        lib().orc_mpc_get_stats(self.h, sp_)
        return dict(status=int(s[0]), iter=int(s[1]), pri_res=s[2], dua_res=s[3], rho=s[4], rho_updates=int(s[5]))

    def get_next_control(self):
        o, op = _out(3)
        lib().orc_mpc_get_next_control(self.h, op)
        return o

    def simulate_step(self, t, dt=0.01):
        lib().orc_mpc_simulate_step(self.h, C.c_double(t), C.c_double(dt))

    def step(self, t0):
        self.compute_time_steps(t0)
        self.compute_linearization_nodes()
        self.update_qp()
        self.solve()
        return self.get_next_control()


def batch_step(mpcs, t0, rollout=False, dt_sim=0.01, nthreads=0):
    n = len(mpcs)
    hs = (C.c_void_p * n)(*[m.h for m in mpcs])
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (n,)))
    out, outp = _out(n, 3)
    lib().orc_mpc_batch_step(hs, n, t0.ctypes.data_as(dp), outp, int(rollout), C.c_double(dt_sim), int(nthreads))
    return out


def max_threads():
    return int(lib().orc_max_threads())
