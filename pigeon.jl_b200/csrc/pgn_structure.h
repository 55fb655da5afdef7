// pgn_structure.h — host-side static analysis of the tracking QP (done once per handle in pgn_create).
//
// The reference freezes the QP sparsity pattern at construction (construct_coupled_tracking_QP, coupled_lat_long.jl:197-313;
// construct_lateral_tracking_QP, decoupled_lat_long.jl:134-226) and only pushes new *values* each step (update_QP!).  The same
// holds here, taken further: because the pattern is identical for every vehicle and every step, the elimination ordering of
// the OSQP KKT matrix, its symbolic LDL' factor, the level schedule of the triangular solves and the gather lists of the
// numeric factorisation are all computed once on the host and uploaded as read-only tables shared by every CTA.
#pragma once
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#define PGN_HOSTDEV __host__ __device__
#else
#define PGN_HOSTDEV
#endif

// threads per ADMM CTA: the step programs of the triangular solves are laid out for exactly this many lanes
#ifndef ADMM_THREADS
#define ADMM_THREADS 512
#endif

// in-place range inverse: a warp keeps the results of at most this many tasks of one level in registers before writing them back
#define INV_MAX_TASKS_PER_WARP 4
#define FAC_TGT_PIVOT 0x80000000u
// solve phases (two per level range and direction) whose task ranges travel in the kernel parameters
#define ADMM_MAX_PHASES 48

namespace pgn {

// dense tail sweep of the ADMM kernel: eight-element row segments of the packed lower triangle of dimension D (one thread each)
inline int tail_segments(int D) { int n = 0; for (int i = 0; i < D; i++) n += i / 8 + 1; return n; }

// Ruiz norm program of the 256-thread ADMM kernels: thread t owns the positions rz_pos[u * 256 + t], u < RZP_U, handed out in order of
// decreasing adjacency length, so class u needs at most RZP_K[u] entry slots; the A value indices of a thread's RZP_SLOTS slots (lists padded
// by repeating their first entry: a maximum does not mind) sit in registers over the ten equilibration passes
#define RZP_U 5
#define RZP_NT 256
#define RZP_SLOTS 38
PGN_HOSTDEV constexpr int rzp_k(int u) { return u == 0 ? 16 : u == 1 ? 12 : u == 4 ? 2 : 4; }

// layout of the per-vehicle "QP piece record" written by the linearisation / HJI kernels and gathered by the ADMM kernel
struct RecLayout {
    int nx, nu, T;
    int piece_len;                      // doubles per interval
    int oA, oB0, oBf, oc, oH, oG, odmin, odmax, ofxmax;   // offsets inside a piece
    int o_qcurr, o_ucurr, o_hji, o_dt;  // offsets of the global part
    int rec_len;
    PGN_HOSTDEV int piece(int t) const { return t * piece_len; }
};

// constant table shared by all vehicles
enum { CT_ZERO = 0, CT_PINF, CT_NINF, CT_VMIN, CT_VMAX, CT_FXMIN_N, CT_DDELTA_N, CT_LEN };

// sources of bound values
enum { BND_CONST = 0, BND_REC = 1, BND_NEG_REC = 2, BND_DT_SCALED = 3, BND_NEG_DT_SCALED = 4 };
// modes of the cost tables
enum { PQ_ZERO = 0, PQ_TIMES_DT = 1, PQ_OVER_DT = 2, PQ_CONST = 3 };
// flags of a solve task (see WTask): out[r] = f(in[r], acc) with acc = sum of the task's products
//   default            out = in - acc
//   TASK_ADD           out = in + acc
//   TASK_SCALE_OUT     out = Dinv[r] * (in +- acc)
//   TASK_SCALE_ACC     out = in +- Dinv[r] * acc
enum { TASK_SRC_TMP = 1 << 0, TASK_DST_TMP = 1 << 1, TASK_ADD = 1 << 2, TASK_SCALE_OUT = 1 << 3, TASK_SCALE_ACC = 1 << 4 };
// weight ids (resolved against the control parameters at run time so that pgn_set_control_params needs no re-analysis)
enum { W_NONE = 0, W_Q_DS, W_Q_DPSI, W_Q_E, W_R_DELTA, W_R_FX, W_R_DDELTA, W_R_DFX, W_W_BETA, W_W_R, W_W_HJI, W_LEN };

struct QpTables {
    int kind, N, T, Ns, nx, nu;
    int n, m, Nk, nnzA, nnzL, nlev;
    RecLayout rec;
    // canonical QP (construction order of the reference)
    std::vector<int32_t> a_row, a_col, a_src;           // A entries: src >= 0 -> rec index, -1 -> +1.0, -2 -> -1.0
    std::vector<uint8_t> l_type, u_type;                // per constraint
    std::vector<int32_t> l_idx, u_idx;
    std::vector<uint8_t> P_mode, q_mode, P_w, q_w;      // per variable: mode + weight id
    std::vector<uint16_t> P_t, q_t;                     // index of the interval whose dt scales the weight
    std::vector<uint16_t> q_hji_t;                      // for W_HJI: short-step index (weight active iff t < N_HJI), 0xFFFF otherwise
    // KKT ordering: position of variable j / constraint i in the elimination order (sorted by level)
    std::vector<uint16_t> pos_var, pos_con;
    std::vector<uint8_t> is_con;                        // per position
    std::vector<uint16_t> pos2idx;                      // per position: j or i
    // L (unit lower triangular) in CSR by rows (values are stored in this order) and its CSC view
    std::vector<uint16_t> lrow_ptr, lrow_col, lcol_ptr, lcol_row, lcol_val, lvl_ptr;
    // A entries in position space
    std::vector<uint16_t> a_rowpos, a_colpos, a_lpos;
    // off-diagonal KKT adjacency in position space: for position p, entries (A value index, neighbour position)
    std::vector<uint16_t> kadj_ptr, kadj_e, kadj_nb;
    // Ruiz norm program (see RZP_*): rz_prog = 1 if the adjacency lengths fit its slot classes; rz_idx[k * 256 + t] = slots 2k | 2k+1 << 16 of thread t
    int rz_prog;
    std::vector<uint16_t> rz_pos;
    std::vector<uint32_t> rz_idx;
    // level ranges [la, lb) of the sparse part whose in-range block of L is replaced by its explicit inverse after factorisation
    std::vector<int> range_lvl;
    // ---- warp programs -----------------------------------------------------------------------------------------------------------
    // Every data-parallel phase of the factorisation and of the triangular solves is a list of warp tasks.  One task = one warp:
    // 32 >> sh rows (targets), 2^sh adjacent lanes per row, K entry slots per lane; the entries of a task are stored slot-major
    // (index ebase + k * 32 + lane) so that every access to a program array is one coalesced / conflict-free warp access.
    //   word 0: ebase          word 1: rbase | nrows << 16 | sh << 24          word 2: K | flags << 16          word 3: spare
    // Padding entries multiply the always-zero L slot `zslot` with the always-zero vector element Nk.
    //
    // The factor is kept in the unscaled form W = L D (W_ij = L_ij d_j): W_ij = K_ij - sum_k W_ik W_jk / d_k needs no column scaling
    // phase, and the solves absorb 1/d (flags above).  L values live in the order the FORWARD solve consumes them (`nslots` slots).
    int nwarps;                                        // warps the programs below were scheduled for
    int nslots, zslot;
    std::vector<uint32_t> sol_task;                    // solve tasks, 4 words each: forward phases then backward phases
    std::vector<uint16_t> sol_ph_ptr;                  // phase -> first task; n_fwd_ph forward phases, then n_bwd_ph backward phases
    int n_fwd_ph, n_bwd_ph;
    std::vector<uint16_t> sol_orow;                    // output position per (task, row)
    std::vector<uint16_t> fidx;                        // forward: source position per L slot (size nslots)
    std::vector<uint32_t> bent;                        // backward entries in program order: L slot | source position << 16
    int rhs_tmp_end;                                   // positions [0, rhs_tmp_end) = first range: their right-hand side goes to the scratch vector
    // Rows without entries are kept out of the biggest phases (they are 13 + 15 + 7 of the 28 + 28 + 27 tasks of the first range of the coupled
    // N = 31 QP):  positions [0, fwd_k0_end) — level 0: y^ = t / d — are written by whoever forms the right-hand side;  the columns bwd_k0 of the
    // first range have nothing below the range (v = y^): they are copied sol -> scratch by the warps that phase bwd_k0_phase leaves idle
    // (bwd_k0_warp0 = number of tasks of that phase), and those among them without in-range successors either (x = v = y^) need no task at all
    // likewise the factorisation: the pivots of level 0 have no update terms, [0, fac_k0_end) are inverted where K_jj is written and level 0 has no pass
    int fac_k0_end;
    int fwd_k0_end, bwd_k0_phase, bwd_k0_warp0;
    std::vector<uint16_t> bwd_k0;
    // Tensor-memory layout of the L values (tmem_layout = 1, the two-QPs-per-SM variant for QPs whose factor does not fit shared memory twice):
    // task t of a phase runs on warp (t - first task of the phase) % nwarps, and a warp reaches only the 32 TMEM lanes of quadrant warp % 4; the
    // K slot rows of task t occupy the 2 K 32-bit columns [sol_tcol[t], sol_tcol[t] + 2 K) of that quadrant — lane l, slot k of the task is the
    // double at (lane 32 q + l, column sol_tcol[t] + 2 k): the slot-major program layout IS the TMEM layout.  Backward tasks get their own
    // copy in program order (bsrc[e] = source position of backward entry e; the L slot it copies from stays in `bent`).  Tasks are permuted
    // inside every round of `nwarps` so that the four quadrants fill evenly; tmem_cols = columns the fullest quadrant needs (+ read slack).
    int tmem_layout, tmem_cols;
    std::vector<uint16_t> sol_tcol, bsrc;
    // numeric factorisation: per level a list of tasks; target: FAC_TGT_PIVOT | position, or a slot (an L slot, or nslots + packed lower
    // index of the dense tail Schur complement, gathered by the last pass); entry: a | b << 16 | k << 32  (product W[a] * W[b] / d_k)
    std::vector<uint32_t> fac_task, fac_lvl_ptr, fac_tgt;
    std::vector<uint64_t> fac_ent;
    // in-place inversion of the range blocks (unit lower M = L_RR^-1), level by level: target (slot | col << 16),
    //   M_ij = -(W_ij / d_j + sum_{j<k<i} W_ik / d_k * M_kj),   entry a | b << 16 | k << 32  (W[a] * M[b] / d_k)
    std::vector<uint32_t> inv_task, inv_lvl_ptr, inv_tgt;
    std::vector<uint64_t> inv_ent;
    int inv_max_tasks_per_warp;
    // A entries -> L slot
    std::vector<uint16_t> a_slot;
    // dense tail: the Schur complement of the last `tail_dim` positions (levels >= tail_level, the top separators) is formed densely
    // (packed lower, with diagonal) and inverted on chip by a symmetric sweep after every numeric factorisation; in the solves it
    // replaces tail_dim narrow levels by one dense symmetric mat-vec
    int tail_level, tail_start, tail_dim;
    // where the solution components consumed by the host-side API live
    int var_u1_delta, var_u1_fx;                        // variable indices of u[:,2] (node 2)
};

// ordering: 0 nested dissection over stages, 1 minimum degree
// nwarps: warps of the ADMM CTA the warp programs are laid out for (ADMM_THREADS / 32 by default; 8 for the two-CTAs-per-SM variant)
// tmem_layout: 1 = balance the solve tasks over the TMEM quadrants and emit sol_tcol / bsrc (see QpTables)
bool build_qp_tables(int kind, int N_short, int N_long, int ordering, QpTables& out, char* err, int errlen, int nwarps = ADMM_THREADS / 32, int tmem_layout = 0);

}  // namespace pgn
