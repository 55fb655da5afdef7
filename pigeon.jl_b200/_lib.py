"""ctypes binding of libpigeon_b200.so (include/pigeon_b200.h).  The CUDA library is the ONLY compute path: if it is missing
or no sm_100 device is usable the calls raise — there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGN_LIB_PATH", os.path.join(HERE, "libpigeon_b200.so"))    # override: A/B builds of the same C ABI

PGN_COUPLED, PGN_DECOUPLED = 0, 1
STATUS_NAMES = {1: "solved", 2: "solved_inaccurate", 3: "primal_infeasible_inaccurate", 4: "dual_infeasible_inaccurate", -2: "max_iter_reached",
                -3: "primal_infeasible", -4: "dual_infeasible", -10: "unsolved"}


class PgnConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("batch", C.c_int32), ("N_short", C.c_int32), ("N_long", C.c_int32), ("dt_short", C.c_double),
                ("dt_long", C.c_double), ("use_correction_step", C.c_int32), ("device", C.c_int32), ("rho", C.c_double), ("sigma", C.c_double),
                ("alpha", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("max_iter", C.c_int32), ("scaling", C.c_int32), ("check_termination", C.c_int32), ("adaptive_rho", C.c_int32),
                ("adaptive_rho_interval", C.c_int32), ("adaptive_rho_tolerance", C.c_double), ("warm_start", C.c_int32), ("rk4_substeps", C.c_int32),
                ("hji_eps", C.c_double), ("kkt_ordering", C.c_int32), ("reserved", C.c_int32)]


# every symbol declared in include/pigeon_b200.h
SYMBOLS = ["pgn_default_config", "pgn_x1_vehicle_params", "pgn_default_control_params", "pgn_create", "pgn_destroy", "pgn_last_error", "pgn_set_stream",
           "pgn_synchronize", "pgn_set_vehicle_params", "pgn_set_control_params", "pgn_set_trajectories", "pgn_assign_trajectories", "pgn_set_hji_cache",
           "pgn_set_state", "pgn_reset_solved", "pgn_reset_solver", "pgn_set_guards", "pgn_compute_time_steps", "pgn_compute_linearization_nodes", "pgn_update_qp",
           "pgn_solve", "pgn_get_next_control", "pgn_step", "pgn_step_device", "pgn_simulate", "pgn_rollout", "pgn_qp_dims", "pgn_get_state",
           "pgn_get_time_steps", "pgn_get_nodes", "pgn_set_nodes", "pgn_get_qp_data", "pgn_get_solution", "pgn_get_stats", "pgn_hji_lookup",
           "pgn_hji_lookup_device", "pgn_device_controls", "pgn_device_stats", "pgn_set_profiling", "pgn_get_stage_ms", "pgn_get_admm_cycles", "pgn_get_admm_trace",
           "pgn_get_hji_values", "pgn_hji_optimal_control", "pgn_set_hji_policy", "pgn_from_autobox", "pgn_step_rollout_device", "pgn_set_path_search_window",
           "pgn_simulate_device", "pgn_set_pipeline_parts", "pgn_get_pipeline_parts", "pgn_set_history", "pgn_get_history", "pgn_comm_unique_id",
           "pgn_set_solve_cap", "pgn_set_hji_lookup_order", "pgn_step_submit", "pgn_step_collect", "pgn_steps_in_flight", "pgn_comm_init_rank", "pgn_comm_init_all", "pgn_comm_destroy", "pgn_gather", "pgn_gather_all"]

_lib = None


class PigeonError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PigeonError(f"{LIB_PATH} is missing: build it with `python pigeon.jl_b200/build.py` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.pgn_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(lib, name)   # AttributeError if the library does not export a declared symbol
    lib.pgn_create.argtypes = [C.POINTER(PgnConfig), C.POINTER(C.c_void_p)]
    lib.pgn_get_admm_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32]
    lib.pgn_simulate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int32]
    lib.pgn_simulate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32]
    lib.pgn_set_pipeline_parts.argtypes = [C.c_void_p, C.c_int32]
    lib.pgn_get_pipeline_parts.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.pgn_rollout.argtypes = [C.c_void_p, C.c_double]
    lib.pgn_step_rollout_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    lib.pgn_set_guards.argtypes = [C.c_void_p, C.c_int32, C.c_double]
    lib.pgn_step_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pgn_step_collect.argtypes = [C.c_void_p, C.c_void_p]
    lib.pgn_steps_in_flight.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.pgn_set_solve_cap.argtypes = [C.c_void_p, C.c_int32]
    lib.pgn_set_hji_lookup_order.argtypes = [C.c_void_p, C.c_int32]
    lib.pgn_set_history.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.pgn_get_history.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pgn_comm_unique_id.argtypes = [C.c_void_p]
    lib.pgn_comm_init_rank.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.pgn_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.pgn_comm_destroy.argtypes = [C.c_void_p]
    lib.pgn_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pgn_gather_all.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PigeonError(f"libpigeon_b200 error {rc}: {load().pgn_last_error().decode()}")


def dptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected array of shape {shape}, got {a.shape}")
    return a
