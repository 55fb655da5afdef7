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
#define ADMM_THREADS 1024
#endif

namespace pgn {

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
// step flags (word 1 >> 8)
enum { STEP_LAST = 1, STEP_SEG_FWD_EXT = 0 << 1, STEP_SEG_FWD_IN = 1 << 1, STEP_SEG_BWD_IN = 2 << 1, STEP_SEG_BWD_EXT = 3 << 1, STEP_SEG_MASK = 3 << 1,
       STEP_SRC_TMP = 1 << 3, STEP_DST_TMP = 1 << 4, STEP_ADD = 1 << 5, STEP_SCALE = 1 << 6 };
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
    // numeric factorisation program (left-looking gathers, level scheduled)
    std::vector<uint32_t> ftgt_ptr;                     // per level -> range of targets
    std::vector<uint16_t> ftgt_id, ftgt_col;            // target: L value index (< nnzL) or nnzL + column for a diagonal; its column
    std::vector<uint32_t> fac_ptr;                      // per target -> range of pairs
    std::vector<uint16_t> fac_a, fac_b, fac_k;          // pair: L value indices (row i col k), (row j col k) and the column k
    // lanes cooperating on one row / column / factor target, per level (powers of two)
    std::vector<uint8_t> lvl_gf, lvl_gb, lvl_gfac;
    // level ranges [la, lb) of the sparse part whose in-range block of L is replaced by its explicit inverse after factorisation
    std::vector<int> range_lvl;
    // per-row segment descriptors (first entry | count << 16): forward CSR row = [external | in-range], backward CSC column = [in-range | external]
    std::vector<uint32_t> fwd_ext, fwd_in, bwd_in, bwd_ext;
    // flattened step programs of the triangular solves.  One step = one pass of the CTA: rows [r0, r0+rows) handled by 2^sh lanes
    // each, <= 4 entries per lane; word 0 = r0 | rows << 16, word 1 = sh | STEP_* flags << 8
    std::vector<uint32_t> step_f, step_b;
    // in-place inversion program of the range blocks: targets (L value indices) per in-range level, pairs (S_ik, M_kj) per target
    std::vector<uint32_t> itgt_ptr, inv_ptr;
    std::vector<uint16_t> itgt_id, inv_a, inv_b;
    // dense tail: the last `tail_dim` positions (levels >= tail_level) form a (nearly dense) unit lower triangular block whose explicit
    // inverse is rebuilt after every numeric factorisation; it replaces tail_dim narrow levels by two dense mat-vec levels
    int tail_level, tail_start, tail_dim;
    std::vector<uint16_t> tl_src, tl_dst;               // sparse L entry -> packed strictly-lower dense index i*(i-1)/2 + j
    // where the solution components consumed by the host-side API live
    int var_u1_delta, var_u1_fx;                        // variable indices of u[:,2] (node 2)
};

// ordering: 0 nested dissection over stages, 1 minimum degree
bool build_qp_tables(int kind, int N_short, int N_long, int ordering, QpTables& out, char* err, int errlen);

}  // namespace pgn
