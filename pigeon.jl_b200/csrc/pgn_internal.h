// pgn_internal.h — handle layout and device-side table views shared by the translation units of libpigeon_b200.so.
//
// HBM layout (B = batch):
//   state / control / other car / time offset / flags      SoA  [field][B]      (thread-per-vehicle kernels, coalesced)
//   ts, dt, prev_ts                                         [B][N], [B][T], [B][N]
//   nodes qs / us / ps                                      [B][N][nx], [B][N][2], [B][N][4]
//   QP piece record (A,B0,Bf,c,H,G,limits per interval + q_curr,u_curr,hji,dt)   [B][rec_len]   (one CTA gathers one record)
//   ADMM warm iterates in KKT-position space (x|z and y)    [B][Nk] each, rho [B]
//   QP solution x [B][n], y [B][m]; stats                   SoA [B]
//   trajectories                                            [12][n_traj][n_nodes]
//   HJI grid                                                V f32 [n1..n7] (dim 1 fastest), gradV f32 [8][n1..n7]-interleaved (7 -> 8 padded)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pigeon_b200.h"
#include "pgn_device.cuh"
#include "pgn_structure.h"

#define PGN_MAX_PARTS 8
#define PGN_RING 4          // steps in flight of the pipelined host-buffer API (pgn_step_submit / pgn_step_collect)

namespace pgn {

struct TrajView {
    const double* f[12];   // t,s,V,A,E,N,psi,kappa,theta,phi,edge_L,edge_R  each [n_traj][n_nodes]
    int n_traj, n_nodes;
};

struct HjiView {
    int dims[7];
    int kofs[7];             // offset of each dimension's knots in `knots`
    const float* knots;
    const float* V;          // [n1..n7], dim 1 fastest
    const float* gV;         // 8 floats per node (7 components + pad), dim 1 fastest over nodes
    long long stride[7];
    int valid;
};

// device views of the static QP tables
struct QpDev {
    int kind, N, T, Ns, nx, nu, n, m, Nk, nnzA, nnzL, nlev, rec_len, o_dt;
    const int32_t* a_src;
    const uint16_t *a_rowpos, *a_colpos, *a_lpos;
    const uint8_t *l_type, *u_type;
    const int32_t *l_idx, *u_idx;
    const uint8_t *P_mode, *q_mode, *P_w, *q_w;
    const uint16_t *P_t, *q_t, *q_hji_t;
    const uint16_t *pos_var, *pos_con, *pos2idx;
    const uint8_t* is_con;
    const uint16_t *kadj_ptr, *kadj_e, *kadj_nb;
    const uint16_t* rz_pos;                     // Ruiz norm program (pgn_structure.h, RZP_*)
    const uint32_t* rz_idx;
    int rz_prog;
    const uint32_t* a_rc;                       // per A entry: row position | column position << 16
    const uint16_t* a_slot;                     // per A entry: L slot
    // warp programs (pgn_structure.h): packed task descriptors {ebase/32 | rbase << 16, K | nrows << 8 | sh << 16 | flags << 24}
    const uint2 *sol_task, *fac_task, *inv_task;
    const uint16_t *sol_orow, *fidx, *sol_tcol, *bsrc;
    const uint32_t* bent;
    const uint32_t *fac_lvl_ptr, *fac_tgt, *inv_lvl_ptr, *inv_tgt;
    const unsigned long long *fac_ent, *inv_ent;
    const uint16_t* bwd_k0; int n_bwd_k0, bwd_k0_phase, bwd_k0_warp0, fwd_k0_end, fac_k0_end;      // task-less rows of the first range (pgn_structure.h)
    int nslots, zslot, rhs_tmp_end, n_fwd_ph, n_bwd_ph, n_sol_task, n_fac_task, n_inv_task, n_bent, n_orow, n_orow_fwd, n_fac_lvl, n_inv_levels;
    int tail_level, tail_start, tail_dim;
    const double* ctab;      // [CT_LEN]
    const double* wtab;      // [W_LEN] cost weights
    int n_hji;               // N_HJI
    int var_u1_delta, var_u1_fx;
};

struct AdmmSettings {
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
    int max_iter, scaling, check_termination, adaptive_rho, adaptive_rho_interval, warm_start;
};

}  // namespace pgn

struct pgn_handle {
    int prio_least, prio_greatest, admm_low_priority;      // stream priority range of the device; 1 = part streams high, ADMM launches low
    pgn_config cfg;
    int device;
    cudaStream_t stream, own_stream;
    int B, N, T, nx, nu;
    pgn::VehParams veh;
    pgn::CtrlParams ctl;
    double un[2];
    pgn::QpTables tab;
    pgn::QpDev qd;
    pgn::AdmmSettings st;
    std::vector<void*> allocs;        // every device allocation (freed in pgn_destroy)
    // per-vehicle device buffers
    double *d_state, *d_control, *d_other, *d_toff;      // [6][B], [3][B], [4][B], [B]
    uint8_t* d_solved;                                   // [B]
    int32_t* d_traj_id;                                  // [B]
    double *d_ts, *d_dt, *d_prev_ts;                     // [B][N], [B][T], [B][N]
    double *d_qs, *d_us, *d_ps;                          // nodes
    double* d_rec;                                       // [B][rec_len]
    double *d_ws_xz, *d_ws_y, *d_rho;                    // warm iterates
    double *d_sol_x, *d_sol_y;                           // [B][n], [B][m]
    int32_t *d_iters, *d_status, *d_rho_updates;
    double *d_pri_res, *d_dua_res;
    double* d_controls;                                  // [3][B]
    double *d_t0, *d_t0_base;                            // [B]
    double* d_hji_val;                                   // [8][B]: gradV[0..6], V of the step's relative state (written by the HJI constraint kernel)
    int hji_policy;                                      // use_HJI_policy[] (ros_integration.jl:47): V <= HJI_eps => optimal_control replaces the QP control
    uint8_t *d_skip, *d_cold;                            // guards: vehicle paused this step / ADMM iterates to be re-initialised
    int guard_nan; double guard_pause;
    int path_window; int32_t* d_last_seg;                // windowed closest-segment search: half-width in segments (0 = full scan), previous segment per vehicle (-1 = none)
    // callback entry point (pgn_from_autobox): packed message buffer (pinned host + device), time-interval flags, path coordinates,
    // and the CUDA graph of the whole call (H2D copy, unpack, the five step stages, pack, D2H copy), re-captured when a setter bumps `epoch`
    // plant rollout beside the ADMM launch (pgn_step_rollout_device / pgn_simulate): shadow state, side stream, fork / join events
    double* d_state_next; cudaStream_t side_stream; cudaEvent_t ev_fork, ev_join;
    // pipeline parts (pgn_set_pipeline_parts): the fused entry points run the batch as `parts` contiguous vehicle ranges, each on its own
    // stream, so that the thread-per-vehicle stages of one range run while the ADMM kernel of another drains.  Launchers read the range
    // of the current launch from (v0, nv, part); outside the fused entry points it is the whole batch (0, B, 0).
    int parts, parts_created, v0, nv, part;
    cudaStream_t part_stream[PGN_MAX_PARTS], part_side[PGN_MAX_PARTS];
    cudaEvent_t part_begin, part_done[PGN_MAX_PARTS], part_evf[PGN_MAX_PARTS], part_evj[PGN_MAX_PARTS];
    double* d_se0;                                       // decoupled node generation: (s0, e0) handed from the scan kernel to the rollout kernel
    double *h_io, *d_io, *d_se; uint8_t* d_tskip; int in_callback;
    cudaGraph_t cb_graph; cudaGraphExec_t cb_exec; long long epoch, cb_epoch, cb_launches; cudaStream_t cb_stream; int cb_has_exec;
    int32_t* d_order;                                    // ticket -> vehicle order of the ADMM launch
    int* d_counter;                                      // work-queue ticket for the persistent ADMM kernel
    unsigned long long* d_trace; int* d_trace_n; int trace_cap;      // profiling 3: (start ns, end ns, part | QPs << 8 | SM << 32) of every ADMM CTA
    unsigned long long* d_cycles;                        // [8] per-phase cycle counters of the ADMM kernel (profiling only)
    double* d_stage;                                     // AoS<->SoA staging
    size_t stage_bytes;
    // host inputs of pgn_set_state / pgn_step: ONE packed pinned buffer [q 6B | u 3B | other 4B | toff B | t0 B] -> one H2D copy -> one unpack
    // kernel; results [B][3] come back through the pinned tail.  ev_in fences the reuse of the pinned buffer.
    double *h_in, *d_in; cudaEvent_t ev_in; int in_pending;
    // pipelined host-buffer stepping (pgn_step_submit / pgn_step_collect): a ring of PGN_RING steps in flight.  Slot s owns a pinned input block
    // [q 6B | u 3B | other 4B | t0 B] + a pinned output block [B][3], their device twins, the event of its H2D copy and one completion event
    // per pipeline part (kernels + the part's D2H copy).  ring_head = next slot to submit, ring_tail = oldest slot not yet collected.
    double *h_ring, *d_ring; cudaEvent_t ring_h2d[PGN_RING], ring_done[PGN_RING][PGN_MAX_PARTS]; int ring_flags[PGN_RING], ring_parts[PGN_RING];
    int ring_head, ring_tail, ring_count, ring_created;
    uint8_t* d_mask;                                     // [B] staging of the masks of pgn_reset_solved / pgn_reset_solver
    // history recorder of simulate (model_predictive_control.jl:84-99 returns qs, xs, us, ps per step): [n_rec][6 + 3 + nx + 4][B], written on
    // the device every `hist_stride` steps
    double* d_hist; int hist_cap, hist_stride, hist_n;
    // final gather (pgn_gather): NCCL communicator (ncclComm_t) of this handle, its rank / size, device staging of the gathered arrays
    void* comm; int comm_rank, comm_size; double* d_gath_c; int32_t* d_gath_i;
    // trajectories / HJI
    pgn::TrajView traj; bool have_traj, have_assign;
    pgn::HjiView hji;
    int hji_sort;                                        // stand-alone lookups: -1 automatic (cell order from 2^19 queries up, TMA-staged tiles when the blocks are well filled), 0 input order, 1 cell order, 2 cell order + TMA tiles
    void* d_hji_ws; size_t hji_ws_bytes;                 // work space of the cell-ordered lookup
    alignas(64) unsigned char hji_tmap[128];             // CUtensorMap of the record table (5-D view) for the TMA-staged lookup
    int hji_tma_valid, hji_tma_tile_bytes; long long hji_tma_blocks;
    // profiling
    int profiling; cudaEvent_t ev[2]; double stage_ms[8]; long long launches;
    int admm_smem_bytes, admm_threads, num_sms;
    // Deferred solves inside the simulate loops (pgn_set_solve_cap): a QP that has not terminated after `solve_cap` iterations of one ADMM
    // launch saves its iterates and continues in the launch of the NEXT round, while its vehicle holds (no new step) and all others go on; every
    // vehicle therefore counts its own steps.  d_hold: 0 steps normally, 1 solve continues, 2 reached the target step count.
    int solve_cap, sim_cap, round_cap, hold_on, sim_target, sim_open, sim_axis_valid; double sim_dt; long long catchup_rounds;
    // one simulate round per pipeline part as a CUDA graph (captured once, replayed every round)
    cudaGraph_t rg_graph2[PGN_MAX_PARTS]; cudaGraphExec_t rg_exec2[PGN_MAX_PARTS]; int split_rounds;      // split rounds: the part of a round after the QP solve
    cudaGraph_t rg_graph[PGN_MAX_PARTS]; cudaGraphExec_t rg_exec[PGN_MAX_PARTS]; long long rg_epoch[PGN_MAX_PARTS], rg_launches[PGN_MAX_PARTS];
    double rg_dt[PGN_MAX_PARTS]; int rg_cap[PGN_MAX_PARTS], rg_rec[PGN_MAX_PARTS], rg_v0[PGN_MAX_PARTS], rg_nv[PGN_MAX_PARTS];
    uint8_t* d_hold; int32_t *d_kstep, *d_iters_acc; int *d_lag, *h_lag;
    int admm_tmem, admm_ctas_per_sm;                     // tensor-memory variant of the ADMM kernel (two coupled N = 31 QPs per SM); resident CTAs per SM
    double* d_admm_scratch;                              // its per-CTA global scratch
};

namespace pgn {
// kernels (defined in the .cu files)
void launch_time_steps(pgn_handle* h, const double* d_t0);
void launch_nodes(pgn_handle* h);
void launch_linearize(pgn_handle* h);
void launch_hji_constraint(pgn_handle* h);
void launch_admm(pgn_handle* h);
void launch_controls(pgn_handle* h, double* d_out);
void launch_callback_in(pgn_handle* h);
void launch_callback_out(pgn_handle* h);
void launch_rollout(pgn_handle* h, double dt);
void launch_propagate_shadow(pgn_handle* h, double dt, cudaStream_t side);
void launch_commit_rollout(pgn_handle* h);
void launch_hji_optimal_control(pgn_handle* h, int M, const double* d_x, const double* d_gV, double* d_out);   // [M][7], [M][7] -> [M][2]
void launch_hji_lookup(pgn_handle* h, int M, const double* d_x, double* d_V, double* d_gV);
bool hji_make_tensor_map(pgn_handle* h);
void launch_transpose_in(pgn_handle* h, const double* d_aos, double* d_soa, int k);    // [B][k] -> [k][B]
void launch_transpose_out(pgn_handle* h, const double* d_soa, double* d_aos, int k);   // [k][B] -> [B][k]
void launch_time_axpy(pgn_handle* h, const double* d_base, double k, double dt, double* d_v, int n);
void launch_round_begin(pgn_handle* h, double dt);      // target step count: d_lag[1]
bool admm_is_kernel(const void* f);                   // is this device function one of the ADMM kernel instantiations? (graph node priorities)
void launch_stamp(pgn_handle* h, int stage);      // profiling 3: time stamp record between the stages of a round
void launch_fill_i32(pgn_handle* h, int32_t* d, int value, int n);      // deferred solves: per-vehicle step time and hold flags of the current range
void launch_count_lag(pgn_handle* h, int target);                   // vehicles with fewer than `target` completed steps -> d_lag
void launch_unpack_range(pgn_handle* h, const double* d_in, int flags);   // the same for the current vehicle range, from a ring slot
void launch_unpack_state(pgn_handle* h, int flags);                 // d_in -> SoA state / control / other / toff / t0 (flags: 1 q, 2 u, 4 other, 8 toff, 16 t0)
void launch_masked_reset(pgn_handle* h, const uint8_t* d_mask, int what);   // what: 1 solved = 0, 2 ADMM iterates = 0 and rho = setting (mask nullptr = all)
void launch_record(pgn_handle* h, int slot);                        // history recorder: (state, control, node 1, params 1) of the current range -> slot
void launch_pack_out(pgn_handle* h, const double* d_soa, double* d_aos, int k);        // [k][B] -> [B][k] for the current vehicle range only
size_t admm_smem_bytes(const QpTables& t, int nthreads, bool tables_in_smem);
size_t admm_smem_bytes_tmem(const QpTables& t, int nthreads);
bool admm_tmem_fits(const QpTables& t, int nthreads);
size_t admm_scratch_doubles(const pgn_handle* h);
int admm_orow_fwd(const QpTables& t);
int admm_configure(pgn_handle* h);   // sets the max dynamic shared memory attribute; returns cudaError
}  // namespace pgn
