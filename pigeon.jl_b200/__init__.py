"""pigeon.jl_b200 — B200-native batched MPC engine for the per-time-step hot path of StanfordASL/Pigeon.jl.

Host-side mirror of the reference's MPC API over the C ABI of libpigeon_b200.so (include/pigeon_b200.h).
"""
from . import synthetic  # noqa: F401
from ._lib import LIB_PATH, PGN_COUPLED, PGN_DECOUPLED, STATUS_NAMES, SYMBOLS, PigeonError, load  # noqa: F401
from .mpc import (BatchedCoupledTrajectoryTrackingMPC, BatchedDecoupledTrajectoryTrackingMPC, BatchedTrajectoryTrackingMPC,  # noqa: F401
                  CoupledControlParams, DecoupledControlParams, HJICache, TrajectoryTube, X1, compute_linearization_nodes, compute_time_steps,
                  comm_init_all, gather_all, get_next_control, placeholder_HJICache, simulate, solve, straight_trajectory, update_QP)
from .world import read_msg, read_world, trajectory_from_msg, trajectory_from_world, write_msg, write_world  # noqa: F401
from .hji_io import load_hji_cache, save_hji_cache  # noqa: F401
