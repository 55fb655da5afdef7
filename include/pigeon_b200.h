/* pigeon_b200.h — C ABI of the B200-native batched MPC engine (libpigeon_b200.so).
 *
 * Drop-in boundary for the per-time-step hot path of StanfordASL/Pigeon.jl.  The reference has no FFI boundary of its
 * own on this path (its only native hop is OSQP.jl's ccall into libosqp); the de-facto operator interface is the 5-call
 * step API on `TrajectoryTrackingMPC` (reference src/model_predictive_control.jl:70-78) with inputs written as struct
 * fields (:32-58).  Each entry point below names the reference interface it replaces.  A Julia host binds these with
 * one-line `ccall`s (julia/PigeonB200.jl); the Python mirror used by the tests binds the same symbols
 * with ctypes (pigeon.jl_b200/_lib.py).
 *
 * Conventions
 *  - every function returns 0 on success or a negative pgn_status; pgn_last_error() gives a thread-local message.
 *  - host arrays are caller-owned, vehicle-major ("[B][k]" = Julia Matrix{Float64}(k, B)), read/written only during the call.
 *  - all device memory, the CUDA stream and the static factorisation schedules are owned by the opaque handle.
 *  - a handle is bound to one CUDA device and must be driven by one host thread at a time; every entry point switches to the handle's device
 *    for the duration of the call and restores the caller's current device, so ONE host thread can drive handles on several GPUs.
 *  - setters are ordered after the work already queued on the handle's stream (they synchronise it, or run on it).
 *  - there is NO CPU fallback: pgn_create fails with PGN_ECUDA when no sm_100 device is usable.
 */
#ifndef PIGEON_B200_H
#define PIGEON_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define PGN_API __attribute__((visibility("default")))
#else
#define PGN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgn_handle pgn_handle;

typedef enum pgn_status {
    PGN_OK = 0,
    PGN_EINVAL = -1,   /* bad argument / call order */
    PGN_ECUDA = -2,    /* CUDA runtime error (message in pgn_last_error) */
    PGN_ENOMEM = -3,
    PGN_ESTATE = -4,   /* required input (trajectories, state, communicator, ...) not set */
    PGN_ENCCL = -5     /* NCCL error or libnccl.so.2 not loadable (pgn_comm_*, pgn_gather only) */
} pgn_status;

enum { PGN_COUPLED = 0, PGN_DECOUPLED = 1 };

/* per-QP solver status, same codes as libosqp */
enum {
    PGN_QP_SOLVED = 1, PGN_QP_SOLVED_INACCURATE = 2, PGN_QP_PRIMAL_INFEASIBLE_INACCURATE = 3, PGN_QP_DUAL_INFEASIBLE_INACCURATE = 4,
    PGN_QP_MAX_ITER_REACHED = -2, PGN_QP_PRIMAL_INFEASIBLE = -3, PGN_QP_DUAL_INFEASIBLE = -4, PGN_QP_UNSOLVED = -10,
    PGN_QP_PENDING = -11      /* inside a simulate loop only: the solve continues in the next round (pgn_set_solve_cap) */
};

/* Mirrors the keyword arguments of CoupledTrajectoryTrackingMPC / DecoupledTrajectoryTrackingMPC
 * (reference src/coupled_lat_long.jl:42-43, src/decoupled_lat_long.jl:32-33) plus the OSQP settings the reference leaves at
 * their defaults (src/coupled_lat_long.jl:201-204). */
typedef struct pgn_config {
    int32_t kind;                 /* PGN_COUPLED | PGN_DECOUPLED */
    int32_t batch;                /* B: number of independent vehicles / scenarios on this device */
    int32_t N_short, N_long;      /* 10, 20 */
    double dt_short, dt_long;     /* 0.01, 0.2 */
    int32_t use_correction_step;  /* 1 */
    int32_t device;               /* CUDA ordinal, -1 = current device */
    /* OSQP-style ADMM settings */
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int32_t max_iter, scaling, check_termination, adaptive_rho, adaptive_rho_interval;
    double adaptive_rho_tolerance;
    int32_t warm_start;
    /* flow integrator used by the coupled linearisation and the plant rollout: RK4 sub-steps per control interval */
    int32_t rk4_substeps;         /* 10 */
    double hji_eps;               /* HJI_ϵ, 0.05 (model_predictive_control.jl:67) */
    int32_t kkt_ordering;         /* 0 = nested dissection over stages (default), 1 = minimum degree */
    int32_t reserved;
} pgn_config;

#define PGN_VEHICLE_PARAMS_LEN 23   /* L,a,b,h,G,m,Izz,mu,Caf,Car,Cd0,Cd1,Cd2,fwd,rwd,fwb,rwb,Fx_max,Fx_min,Px_max,delta_max,kappa_max,inv_fiala_corrected */
#define PGN_CONTROL_PARAMS_LEN 16   /* V_min,V_max,k_V,k_s,ddelta_max,Q_ds,Q_dpsi,Q_e,W_beta,W_r,W_HJI,N_HJI,R_delta,R_ddelta,R_Fx,R_dFx */

/* --- construction --------------------------------------------------------------------------------------------- */
/* defaults of the reference constructors + OSQP defaults */
PGN_API int pgn_default_config(pgn_config* cfg, int32_t kind);
/* X1() (reference src/vehicles.jl:1-59) and Coupled/DecoupledControlParams() defaults (coupled_lat_long.jl:23-38, decoupled_lat_long.jl:18-28) */
PGN_API int pgn_x1_vehicle_params(double* vp /*[23]*/);
PGN_API int pgn_default_control_params(int32_t kind, double* cp /*[16]*/);
/* replaces Coupled/DecoupledTrajectoryTrackingMPC(vehicle, trajectory; ...) incl. construct_*_QP + Parametron.initialize! */
PGN_API int pgn_create(const pgn_config* cfg, pgn_handle** out);
PGN_API int pgn_destroy(pgn_handle* h);
PGN_API const char* pgn_last_error(void);
/* run all work of this handle on an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
PGN_API int pgn_set_stream(pgn_handle* h, void* cuda_stream);
PGN_API int pgn_synchronize(pgn_handle* h);

/* --- inputs (fields of TrajectoryTrackingMPC, model_predictive_control.jl:32-58) --------------------------------- */
PGN_API int pgn_set_vehicle_params(pgn_handle* h, const double* vp /*[23]*/);          /* mpc.vehicle / dynamics */
PGN_API int pgn_set_control_params(pgn_handle* h, const double* cp /*[16]*/);          /* mpc.control_params */
/* mpc.trajectory: n_traj TrajectoryTubes of n_nodes nodes each; fields[k] is [n_traj][n_nodes] in the order
 * t,s,V,A,E,N,psi,kappa,theta,phi,edge_L,edge_R (trajectories.jl:8-20) */
PGN_API int pgn_set_trajectories(pgn_handle* h, int32_t n_traj, int32_t n_nodes, const double* const fields[12]);
PGN_API int pgn_assign_trajectories(pgn_handle* h, const int32_t* traj_id /*[B]*/);
/* mpc.HJI_cache (HJI_computation.jl:26-57): knots concatenated, V column-major (dim 1 fastest), gradV with the 7 components fastest */
PGN_API int pgn_set_hji_cache(pgn_handle* h, const int32_t dims[7], const float* knots, const float* V, const float* gradV);
/* mpc.current_state [B][6], mpc.current_control [B][3], mpc.other_car_state [B][4] (NULL = keep), mpc.time_offset [B] (NULL = keep; NaN = path mode) */
PGN_API int pgn_set_state(pgn_handle* h, const double* q, const double* u, const double* other_car, const double* time_offset);
/* mpc.solved = false for the masked vehicles (mask NULL = all): next node generation is the cold steady-state rollout */
PGN_API int pgn_reset_solved(pgn_handle* h, const uint8_t* mask /*[B]*/);
/* Parametron.initialize!(mpc.model) (ros_integration.jl:146): cold ADMM iterates, rho back to its setting */
PGN_API int pgn_reset_solver(pgn_handle* h, const uint8_t* mask /*[B]*/);

/* Per-vehicle guards of the reference's callback (src/ros_integration.jl:84-87 and 134-147), off by default (`simulate` has none):
 *   pause_below_speed > 0: a vehicle with current_state.Ux < pause_below_speed skips the step (the callback returns early): its QP is
 *                          not solved, solver state and mpc.solved stay as they are, the control output is the current control;
 *   nan_fallback != 0    : NaN in (delta, Fxf, Fxr) => the output is the current control, the vehicle's ADMM iterates / rho are
 *                          re-initialised (Parametron.initialize!) and mpc.solved = false (cold node generation next step). */
PGN_API int pgn_set_guards(pgn_handle* h, int32_t nan_fallback, double pause_below_speed);

/* path_coordinates (trajectories.jl:71-94) scans every segment of the trajectory for the closest one.  half_width > 0 restricts the scan of
 * a vehicle to the segments within half_width of the one found on its previous step (same result whenever the true closest segment lies
 * inside the window — always for continuous motion along a path that does not pass close to itself); the first step after pgn_set_state
 * with a new state, pgn_assign_trajectories, pgn_set_trajectories or this call scans everything.  0 (default) = the reference's full scan. */
PGN_API int pgn_set_path_search_window(pgn_handle* h, int32_t half_width);

/* --- the 5-call step API (model_predictive_control.jl:70-78) ------------------------------------------------------ */
PGN_API int pgn_compute_time_steps(pgn_handle* h, const double* t0 /*[B]*/);           /* compute_time_steps!(mpc, t0) */
PGN_API int pgn_compute_linearization_nodes(pgn_handle* h);                            /* compute_linearization_nodes!(mpc) */
PGN_API int pgn_update_qp(pgn_handle* h);                                              /* update_QP!(mpc) */
PGN_API int pgn_solve(pgn_handle* h);                                                  /* solve!(mpc) */
PGN_API int pgn_get_next_control(pgn_handle* h, double* out /*[B][3] = (delta, Fxf, Fxr)*/);  /* get_next_control(mpc) */
/* the five calls fused (host buffers; copies inside) */
PGN_API int pgn_step(pgn_handle* h, const double* t0 /*[B]*/, double* out /*[B][3]*/);
/* Pipelined host-buffer stepping.  The callback is fed MEASURED states (ros_integration.jl:50-53): step k+1's inputs do not wait for step k's
 * output, so a host that serves many vehicles may keep up to 4 steps in flight.  pgn_step_submit = pgn_set_state(q, u, other, keep time_offset)
 * + pgn_step(t0) without the wait (q / u / other may be NULL = keep); the inputs are copied before it returns.  pgn_step_collect blocks until the
 * OLDEST step in flight is done and copies its controls [B][3] out.  The pipeline parts are not joined between steps (as inside pgn_simulate),
 * which hides the per-vehicle stages behind the ADMM kernels; per vehicle the results are bit-identical to pgn_set_state + pgn_step.  Any other
 * entry point first waits for the steps in flight. */
PGN_API int pgn_step_submit(pgn_handle* h, const double* q, const double* u, const double* other_car, const double* t0 /*[B]*/);
PGN_API int pgn_step_collect(pgn_handle* h, double* out /*[B][3]*/);
PGN_API int pgn_steps_in_flight(pgn_handle* h, int32_t* n);
/* from_autobox_callback (ros_integration.jl:48-151) for the whole batch in ONE call — the low-latency entry point (B = 1 is the
 * reference's deployment).  Writes current_state q [B][6] and current_control u [B][3] (and other_car_state [B][4] unless NULL), picks
 * the MPC time per vehicle: stamp[v] - time_offset[v], or the path_coordinates time when time_offset is NaN (path-tracking mode,
 * :72-75); a vehicle whose time lies outside [0, trajectory.t[end]] (:77-80), or that is paused by pgn_set_guards (:84-87), returns
 * early: no solve, output = current control.  Then the five step calls and out [B][5] = (delta, Fxf, Fxr, s_m, e_m) (:110-116), with
 * the NaN fallback of pgn_set_guards and the policy of pgn_set_hji_policy applied.  Internally one packed H2D copy from pinned
 * memory, one CUDA-graph launch of all kernels, one D2H copy. */
PGN_API int pgn_from_autobox(pgn_handle* h, const double* q, const double* u, const double* other_car, const double* stamp, double* out /*[B][5]*/);
/* same with inputs already resident: t0 and out are DEVICE pointers ([B] and [3][B] field-major); out may be NULL */
PGN_API int pgn_step_device(pgn_handle* h, const double* d_t0, double* d_out);
/* pgn_step_device followed by pgn_rollout(dt), i.e. one iteration of the `simulate` loop (model_predictive_control.jl:87-98), with the plant
 * step launched beside the QP solve: propagate() needs only the state and the control applied during the interval, both known before
 * the solve, so it runs on a side stream into a shadow state and is committed together with the new control.  Same results as the two calls. */
PGN_API int pgn_step_rollout_device(pgn_handle* h, const double* d_t0, double* d_out, double dt);
/* simulate (model_predictive_control.jl:80-100): n_steps closed-loop steps fully on the device:
 * step at t0 + k*dt, plant rollout propagate(dynamics, state, StepControl(dt, control)), apply the new control */
PGN_API int pgn_simulate(pgn_handle* h, const double* t0 /*[B]*/, double dt, int32_t n_steps);
/* the same loop with t0 a DEVICE pointer, enqueued on the handle's stream without a host synchronisation (results are stream-ordered);
 * runs the steps k = k0 .. k0 + n_steps - 1 at t0 + k*dt, so that a long loop can be issued in pieces on one time axis */
PGN_API int pgn_simulate_device(pgn_handle* h, const double* d_t0 /*[B]*/, double dt, int32_t k0, int32_t n_steps);
/* Pipeline parts of the fused entry points (pgn_step, pgn_step_device, pgn_step_rollout_device, pgn_simulate, pgn_simulate_device): the batch is
 * run as `parts` contiguous vehicle ranges, each on its own stream, so that the per-vehicle stages (nodes, linearisation, HJI, controls, plant
 * step) of one range run while the ADMM kernel of another drains; inside pgn_simulate every range runs all its steps without waiting for the
 * others (a vehicle's step k+1 depends only on its own step k: the `for` loop of model_predictive_control.jl:87-98 per vehicle).  Every vehicle's
 * results are bit-identical for any part count.  parts = 0 (what pgn_create sets): chosen from the batch size (4 from 64 vehicles up, else 1);
 * 1: one range on the caller's stream; <= 8.
 * The five single-stage calls and pgn_from_autobox always run the whole batch on the caller's stream. */
PGN_API int pgn_set_pipeline_parts(pgn_handle* h, int32_t parts);
/* Deferred solves inside pgn_simulate / pgn_simulate_device.  Iteration counts of the batch spread from 25 to max_iter = 4000, and one QP is
 * one CTA: a single 4000-iteration QP used to hold every vehicle of its range for the whole solve.  With a cap, a QP that has not terminated
 * after `iters` ADMM iterations of one launch saves its iterates (the warm-start buffers) and continues in the NEXT round's launch while only
 * ITS vehicle waits; every vehicle counts its own steps, and the vehicles that fell behind are caught up at the end of pgn_simulate (or, for
 * pgn_simulate_device, at the next entry point that needs results, pgn_synchronize included).  The continued solve recomputes scaling and
 * factor from the unchanged QP data and resumes at iteration k + 1: every vehicle's results are bit-identical with and without the cap.
 * iters must be a multiple of check_termination and adaptive_rho_interval; 0 = off; -1 (default) = automatic: 400 when a range's ADMM launch is at
 * most two waves of CTAs (there one long solve is the launch time), off for large ranges (their launches absorb such a solve in their many waves,
 * and a deferred one would surface in the catch-up rounds).  The step entry points are never capped. */
PGN_API int pgn_set_solve_cap(pgn_handle* h, int32_t iters);
PGN_API int pgn_get_pipeline_parts(pgn_handle* h, int32_t* parts);
/* one plant rollout + control application (the tail of the simulate loop) */
PGN_API int pgn_rollout(pgn_handle* h, double dt);
/* The return values of simulate (model_predictive_control.jl:84-99: qs, xs, us, ps pushed once per step).  With capacity > 0 the loops of
 * pgn_simulate / pgn_simulate_device record, on the device, every step k with k % stride == 0 (record k / stride, up to `capacity`):
 * current_state and current_control BEFORE the step (qs, us), mpc.qs[1] and mpc.ps[1] of the step (xs, ps).  pgn_get_history copies the
 * records out: qs [n][B][6], us [n][B][3], xs [n][B][nx], ps [n][B][4] (any may be NULL); n_records is always written.  0, 0 = off. */
PGN_API int pgn_set_history(pgn_handle* h, int32_t capacity, int32_t stride);
PGN_API int pgn_get_history(pgn_handle* h, int32_t* n_records, double* qs, double* us, double* xs, double* ps);

/* --- multi-GPU: the batch is sharded over one handle per GPU with NO hot-path traffic; the only collective is this final gather ------------
 * (the loop that is sharded: model_predictive_control.jl:87-98).  NCCL is loaded at run time (dlopen "libnccl.so.2").
 *   one process, one handle per GPU (the Julia deployment):   pgn_comm_init_all(handles, n);  ...steps on every handle...;  pgn_gather_all(...)
 *   one process per GPU:   rank 0: pgn_comm_unique_id(id), ship the 128 bytes to every rank;  each rank: pgn_comm_init_rank(h, n, rank, id);  pgn_gather(h, ...)
 * Gathered arrays are rank-major: controls [n*B][3] = (delta, Fxf, Fxr) of the last step, iters / status [n*B] (any may be NULL).
 * Every rank receives the full result (all-gather over NVLink). */
PGN_API int pgn_comm_unique_id(char* id /*[128]*/);
PGN_API int pgn_comm_init_rank(pgn_handle* h, int32_t nranks, int32_t rank, const char* id /*[128]*/);
PGN_API int pgn_comm_init_all(pgn_handle* const* handles, int32_t n);
PGN_API int pgn_comm_destroy(pgn_handle* h);
PGN_API int pgn_gather(pgn_handle* h, double* controls, int32_t* iters, int32_t* status);            /* collective: every rank calls it */
PGN_API int pgn_gather_all(pgn_handle* const* handles, int32_t n, double* controls, int32_t* iters, int32_t* status);

/* --- outputs / introspection (used by the parity tests) ----------------------------------------------------------- */
PGN_API int pgn_qp_dims(pgn_handle* h, int32_t* out /*[16] = N, nx, nu, n, m, nnz(A), nnz(L), n_levels, L slots, solve phases, factor entries, inverse entries, tail dim, backward entries, ADMM smem bytes, ADMM threads | tensor-memory variant << 16 | resident CTAs per SM << 20*/);
PGN_API int pgn_get_state(pgn_handle* h, double* q /*[B][6]*/, double* u /*[B][3]*/);
PGN_API int pgn_get_time_steps(pgn_handle* h, double* ts /*[B][N]*/, double* dt /*[B][N-1]*/, double* prev_ts /*[B][N]*/);
PGN_API int pgn_get_nodes(pgn_handle* h, double* qs /*[B][N][nx]*/, double* us /*[B][N][2]*/, double* ps /*[B][N][4]*/);
PGN_API int pgn_set_nodes(pgn_handle* h, const double* qs, const double* us, const double* ps);
/* A [B][T][nx][nx], B0/Bf [B][T][nx][nu] (already multiplied by u_normalization for the coupled QP), c [B][T][nx],
 * H [B][T][4][2], G [B][T][4], dmin/dmax/fxmax [B][T], hji [B][3] = (M1*un1, M2*un2, b) */
PGN_API int pgn_get_qp_data(pgn_handle* h, double* A, double* B0, double* Bf, double* c, double* H, double* G, double* dmin, double* dmax,
                    double* fxmax, double* hji);
PGN_API int pgn_get_solution(pgn_handle* h, double* x /*[B][n]*/, double* y /*[B][m]*/);
PGN_API int pgn_get_stats(pgn_handle* h, int32_t* iters, int32_t* status, double* pri_res, double* dua_res, double* rho, int32_t* rho_updates);
/* stand-alone HJI lookup cache[x] (HJI_computation.jl:66-72); x [M][7] host, V [M], gradV [M][7] */
PGN_API int pgn_hji_lookup(pgn_handle* h, int32_t M, const double* x, double* V, double* gradV);
/* V and gradV of the last step's relative state HJIRelativeState(current_state, other_car_state) — what the callback logs and
 * publishes (ros_integration.jl:57-58); (Inf, 0) outside the grid or before the first coupled step.  V [B], gradV [B][7] */
PGN_API int pgn_get_hji_values(pgn_handle* h, double* V, double* gradV);
/* optimal_control(dynamics, relative_state, gradV) (HJI_computation.jl:133-158, uMode = :max, N = 50): x [M][7], gradV [M][7] host,
 * out [M][2] = (delta, Fx) */
PGN_API int pgn_hji_optimal_control(pgn_handle* h, int32_t M, const double* x, const double* gradV, double* out);
/* use_HJI_policy[] of the callback (ros_integration.jl:47,115-118), coupled controllers only: when on, a vehicle whose V <= HJI_eps
 * gets BicycleControl(longitudinal_params, optimal_control(...)) from get_next_control / step instead of the QP's node-2 control
 * (the QP is still solved, as in the callback).  Off by default. */
PGN_API int pgn_set_hji_policy(pgn_handle* h, int32_t on);
/* Visiting order of the stand-alone lookups: 1 = the queries are counting-sorted by grid cell first, so that queries of the same and of
 * neighbouring cells find their corners in L2 (a random query moves 6.6 KB through DRAM for its 4 KB of corners; in cell order the whole set
 * reads about the table size); 2 = cell order, and every block of cells stages ONE tile of corners in shared memory with TMA box loads
 * (cp.async.bulk.tensor.5d over a 5-D view of the table; built and parity-tested, but measured slower than mode 1: DESIGN.md 4.4);
 * 0 = input order; -1 (default) = mode 1 from 2^19 queries up.  Results are bit-identical in every mode. */
PGN_API int pgn_set_hji_lookup_order(pgn_handle* h, int32_t mode);
/* device variant for the HBM roofline micro-benchmark: d_x [7][M] field-major, d_V [M], d_gradV [7][M] */
PGN_API int pgn_hji_lookup_device(pgn_handle* h, int32_t M, const double* d_x, double* d_V, double* d_gradV);
/* device pointers of library-owned buffers (for zero-copy gathers through torch.distributed / NCCL) */
PGN_API int pgn_device_controls(pgn_handle* h, double** d_out /* [3][B] */);
PGN_API int pgn_device_stats(pgn_handle* h, int32_t** d_iters, int32_t** d_status);
/* per-stage device time accumulated with CUDA events since the last reset, milliseconds:
 * [0] time steps + nodes, [1] linearisation + envelope, [2] HJI, [3] ADMM, [4] controls, [5] rollout, [6] launches counted,
 * [7] catch-up rounds run for vehicles whose solves were deferred (pgn_set_solve_cap).
 * on: 0 off; 1 stage timers (stages serial, one pipeline part); 2 = 1 + cycle counters inside the ADMM kernel (pgn_get_admm_cycles);
 * 3 = the cycle counters alone, pipeline parts and graphs untouched (counters of the free-running loop) */
PGN_API int pgn_set_profiling(pgn_handle* h, int32_t on);
/* profiling 3: one record (start ns, end ns, part | QPs solved << 8 | SM << 32; %globaltimer) per ADMM CTA launched since the last reset */
PGN_API int pgn_get_admm_trace(pgn_handle* h, unsigned long long* out /* [3 * max_entries] */, int32_t max_entries, int32_t* n, int32_t reset);
PGN_API int pgn_get_stage_ms(pgn_handle* h, double* out /*[8]*/, int32_t reset);
/* SM cycles spent by the ADMM CTAs per phase while profiling is on (summed over CTAs):
 * [0] gather, [1] Ruiz scaling, [2] LDL' factorisation, [3] triangular solves, [4] x/z/y update, [5] residuals/termination/rho, [6] store, [7] ticket */
PGN_API int pgn_get_admm_cycles(pgn_handle* h, double* out /*[512]: [0..7] phases (all CTAs); CTA 0 only: [16+p] solve phase p, [116..117] dense tail, [126..128] factor init / range inverses / tail, [136+l] factor level l*/, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* PIGEON_B200_H */
