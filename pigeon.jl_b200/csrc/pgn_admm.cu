// pgn_admm.cu — solve!: batched OSQP-style ADMM, one QP per CTA, persistent CTAs pulling vehicles from an atomic ticket.
//
// Replaces Parametron.solve! -> OSQP.update!/osqp_solve (reference src/model_predictive_control.jl:76 with the settings of
// src/coupled_lat_long.jl:201-204; libosqp 0.4.x is not vendored — algorithm per Stellato et al. 2020 and restated on the
// CPU in oracle/osqp_port.hpp).  Per vehicle and per MPC step the CTA
//   1. gathers the QP values (P, q, A, l, u) from the vehicle's piece record through the static source tables,
//   2. re-equilibrates them (modified Ruiz, `scaling` passes, cost scaling) — osqp_update_P_A rescales from scratch,
//   3. factors the quasi-definite KKT matrix [P+sigma I, A'; A, -1/rho] = L D L' with a static elimination order
//      (nested dissection over the horizon stages) and a level-scheduled gather program computed once on the host,
//   4. iterates  (x~,nu) = K^-1 rhs;  x,z,y update with relaxation alpha;  every `check_termination` iterations the unscaled
//      residuals, tolerances and infeasibility certificates; every `adaptive_rho_interval` the rho estimate with refactorisation,
//   5. stores the unscaled solution, the (scaled) warm-start iterates and rho for the next step, and the per-QP statistics.
// Everything the iteration touches lives in shared memory (~180 KB for the coupled N=31 QP): L values, 1/D, the scaled A
// values, seven KKT-length work vectors and the 16-bit index tables of the triangular solves.  All vectors are indexed by KKT
// *position* (elimination order), so no permutation gathers happen inside the loop.  FP64 throughout: the KKT matrix mixes
// sigma = 1e-6 with 1/rho up to 1e6 and the parity target (1e-4 on controls at eps = 1e-3) does not survive an FP32 factor.
#include "pgn_internal.h"

namespace pgn {

#define ADMM_NCYC 256
#define ADMM_TMEM_COLS 256      // tensor-memory columns per CTA of the vtm variant (half an SM)

static const double OSQP_INFTY = 1e20;

struct AdmmArgs {
    QpDev q;
    AdmmSettings st;
    int B;
    const double* rec;
    double *ws_xz, *ws_y, *rho;
    double *sol_x, *sol_y;
    int32_t *iters, *status, *rho_updates;
    double *pri_res, *dua_res;
    uint8_t* solved;
    int* counter;
    const int32_t* order;          // ticket -> vehicle (longest previous solve first)
    const uint8_t* skip;           // guard: paused vehicles are not solved (nullptr = guard off)
    uint8_t* cold;                 // guard: vehicles whose iterates / rho start from scratch (Parametron.initialize! after a NaN)
    uint8_t* hold;                 // deferred solves (nullptr = off): 1 = this QP continues an earlier launch, 2 = vehicle finished; written 0 / 1 here
    int32_t* iters_acc;            // iterations a continued QP has behind it
    int iter_cap;                  // iterations of one QP per launch (0 = unlimited); a multiple of check_termination and adaptive_rho_interval
    uint16_t ph_ptr[ADMM_MAX_PHASES + 1];   // first task of every solve phase (forward phases, then backward phases): uniform constant-bank reads
    double* scratch;              // tensor-memory variant: per-CTA global scratch (scaled A, scalings, spilled vectors), tm_scratch_doubles each
    unsigned long long* trace; int* trace_n; int trace_cap, part;      // optional CTA trace (profiling 3)
    unsigned long long* cycles;   // optional per-phase cycle counters (profiling builds of the host call): gather, ruiz, factor, solve, update, check, store
};

struct Smem {
    double *Lval, *S, *Dinv, *Aval, *xz, *sol, *dxy, *yq, *lo, *hi, *sc, *red;   // S: packed lower dense tail block, directly behind the L slots
    const uint2 *sol_task, *fac_task, *inv_task;      // packed warp-task descriptors
    const uint32_t *bent, *fac_lvl, *inv_lvl;
    const uint16_t *fidx, *orow;
    uint8_t* flag;   // 0 variable, 1 inequality, 2 equality, 3 loose
    // aliases inside the Lval region, valid between the gather and the first factorisation of a QP (Ruiz equilibration)
    uint16_t *kptr, *ke;
    uint32_t* arc;
    // views used while a QP is gathered / equilibrated (identical to Aval, lo, hi, sc, yq except in the tensor-memory variant, where the big
    // shared-memory region is time-shared)
    double *rz_Aval, *rz_lo, *rz_hi, *rz_sc, *rz_yq;
    // tensor-memory variant: spilled copies in the CTA's global scratch, writable index tables of the iteration views, this warp's TMEM base
    // address (lane field = first lane of its quadrant), per-task TMEM column and backward source positions
    double *g_lo, *g_hi, *g_yq;
    uint16_t *fidx_w, *bsrc_w, *orow_bw, *orow_fw, *tcol_w;
    uint2* task_w;
    const uint16_t *bsrc, *tcol, *orow_f, *orow_b;
    uint32_t tm_base;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int vec_len(int Nk) { return (Nk + 2) & ~1; }                 // Nk values + the always-zero element Nk
__host__ __device__ inline int a_pad_len(int nnzA) { return (nnzA + 2) & ~1; }           // A values + at least one always-zero double (padding slots of the Ruiz norm program)
__host__ __device__ inline size_t lval_region_doubles(int nslots, int tail_dim, int Nk, int nnzA) {
    const size_t alias = align_up((size_t)(Nk + 1) * 2 + (size_t)nnzA * 4, 4) + (size_t)nnzA * 4;      // kptr, ke (u16) + arc (u32)
    const size_t a = (alias + 7) / 8, l = (size_t)nslots + (size_t)tail_dim * (tail_dim + 1) / 2;
    return a > l ? a : l;
}

// tensor-memory variant: the time-shared region must hold (a) the factor + dense tail during a factorisation, (b) scaled A | lo | hi | sc |
// adjacency aliases while a QP is equilibrated, (c) y | lo | hi | fidx | bsrc in front of the dense tail during the iterations
__host__ __device__ inline size_t tm_region_doubles(int nslots, int tail_dim, int Nk, int nnzA, int n_bent) {
    const size_t V = vec_len(Nk), A2 = a_pad_len(nnzA);
    const size_t fac = (size_t)nslots + (size_t)tail_dim * (tail_dim + 1) / 2;
    const size_t alias = align_up((size_t)(Nk + 1) * 2 + (size_t)nnzA * 4, 4) + (size_t)nnzA * 4;
    const size_t ruiz = A2 + 3 * V + (alias + 7) / 8;
    return fac > ruiz ? fac : ruiz;
}
__host__ __device__ inline size_t tm_iter_view_doubles(int nslots, int Nk, int n_bent, int n_orow_bwd) {
    return 3 * (size_t)vec_len(Nk) + ((size_t)(nslots + 128) * 2 + (size_t)(n_bent + 128) * 2 + (size_t)n_orow_bwd * 2 + 7) / 8;
}
static int orow_fwd_count(const QpTables& t) {       // output rows of the forward phases = row base of the first backward task
    const size_t first_bwd = t.sol_ph_ptr[t.n_fwd_ph];
    return first_bwd * 4 + 1 < t.sol_task.size() ? (int)(t.sol_task[4 * first_bwd + 1] & 0xffff) : (int)t.sol_orow.size();
}
__host__ __device__ inline size_t tm_scratch_doubles(int Nk, int nnzA) { return (size_t)a_pad_len(nnzA) + 4 * (size_t)vec_len(Nk); }
size_t admm_smem_bytes_tmem(const QpTables& t, int nthreads) {
    const size_t V = vec_len(t.Nk);
    const size_t ntask = t.sol_task.size() / 4;
    const size_t d = tm_region_doubles(t.nslots, t.tail_dim, t.Nk, t.nnzA, (int)t.bent.size()) + 4 * V + 16 * (nthreads / 32) + 8 + ntask +
                     (((ntask + 3) & ~(size_t)3) + orow_fwd_count(t) + 3) / 4;
    return d * 8 + align_up((size_t)t.Nk, 8) + 64;
}
// whether the tensor-memory variant can run this QP: iteration views in front of the dense tail, TMEM columns within one CTA's half of an SM
bool admm_tmem_fits(const QpTables& t, int nthreads) {
    return t.tmem_layout && t.tmem_cols <= 256 && t.nwarps * 32 == nthreads &&
           tm_iter_view_doubles(t.nslots, t.Nk, (int)t.bent.size(), (int)t.sol_orow.size() - orow_fwd_count(t)) <= (size_t)t.nslots &&
           2 * (admm_smem_bytes_tmem(t, nthreads) + 1024) <= (size_t)228 * 1024;
}

// nthreads: threads of the CTA (the reduction scratch holds 16 doubles per warp); tables_in_smem: whether the static warp programs of the
// solves / factor are copied into shared memory (large QPs, one CTA per SM) or read through L1 (small QPs, two CTAs per SM)
size_t admm_smem_bytes(const QpTables& t, int nthreads, bool tables_in_smem) {
    const size_t V = vec_len(t.Nk);
    size_t d = lval_region_doubles(t.nslots, t.tail_dim, t.Nk, t.nnzA) + V + a_pad_len(t.nnzA) + 7 * V + 16 * (nthreads / 32) + 8;
    size_t u64 = (t.sol_task.size() + t.fac_task.size() + t.inv_task.size()) / 4;
    size_t u32 = t.bent.size() + t.fac_lvl_ptr.size() + t.inv_lvl_ptr.size() + 4;
    size_t u16 = (size_t)t.nslots + t.sol_orow.size() + 8;
    if (!tables_in_smem) { u64 = 0; u32 = 4; u16 = 8; }
    return d * 8 + u64 * 8 + align_up(u32 * 4, 8) + align_up(u16 * 2, 8) + align_up((size_t)t.Nk, 8) + 64;
}

// ---- build variants of the kernel ---------------------------------------------------------------------------------------------------
// v512: one QP per SM (the coupled N = 31 QP fills 224 KB of shared memory and, at 128 registers x 512 threads, the register file).
// v256: half the threads and the static tables left in global memory (L1-resident), for QPs small enough that TWO CTAs fit an SM
//       (deployed horizon N = 16, decoupled controller): at 128 registers a 512-thread CTA owns the whole register file, so the small
//       QPs gained nothing from their smaller shared-memory footprint before.
// vtm:  256 threads, two CTAs per SM, L values of the solves in TENSOR MEMORY (tcgen05.ld / tcgen05.st, 256 columns per CTA): the coupled
//       N = 31 QP, whose 60 KB factor cannot sit in shared memory twice.
#define ADMM_TMEM 0
#define ADMM_NT 512
#define ADMM_MINCTAS 1
#define ADMM_TABSMEM 1
namespace v512 {
#include "pgn_admm_kernel.inc"
}
#undef ADMM_NT
#undef ADMM_MINCTAS
#undef ADMM_TABSMEM
#define ADMM_NT 256
#define ADMM_MINCTAS 2
#define ADMM_TABSMEM 0
namespace v256 {
#include "pgn_admm_kernel.inc"
}
#undef ADMM_NT
#undef ADMM_MINCTAS
#undef ADMM_TABSMEM
#undef ADMM_TMEM
#define ADMM_TMEM 1
#define ADMM_MINCTAS 2
#define ADMM_TABSMEM 0
#define ADMM_NT 256
namespace vtm {
#include "pgn_admm_kernel.inc"
}
#undef ADMM_NT
// Measured with 12 and 16 warps per QP (384 / 512 threads, 80 / 64 registers per thread with 72 / 230 bytes of spills, still two CTAs per SM):
// 487 k / 456 k steps/s against 519 k with 8 warps — more warps shorten a lone QP (cold start 255 k against 216 k) but two QPs of 8 warps
// fill the issue slots better than they do (tools/gpu_tmem_threads.sh).  Only the 256-thread build is kept.
#undef ADMM_MINCTAS
#undef ADMM_TABSMEM
#undef ADMM_TMEM

// Ticket order of the persistent CTAs: vehicles whose previous solve took the most iterations go first (longest-processing-time-first),
// so that a slow QP starts at the beginning of the launch instead of becoming its tail.  Counting sort on iters / 25 in one CTA.
__global__ void __launch_bounds__(1024) k_admm_order(const int32_t* __restrict__ iters, int32_t* __restrict__ order, int B, int v0) {
    __shared__ int hist[257];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < B; v += blockDim.x) atomicAdd(&hist[255 - min(iters[v] / 25, 255)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < 256; i++) { const int c = hist[i]; hist[i] = acc; acc += c; }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < B; v += blockDim.x) order[atomicAdd(&hist[255 - min(iters[v] / 25, 255)], 1)] = v0 + v;
}

int admm_configure(pgn_handle* h) {
    const bool small = h->admm_threads == 256 && !h->admm_tmem;
    cudaError_t e;
    if (h->admm_tmem) {
        h->admm_smem_bytes = (int)admm_smem_bytes_tmem(h->tab, h->admm_threads);
        e = cudaSuccess;
        // two CTAs per SM need the whole shared-memory carve-out (2 x 112 KB)
#define TM_ATTR(ns)                                                                                                                                   \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ns::k_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);          \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ns::k_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);           \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ns::k_admm<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ns::k_admm<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        TM_ATTR(vtm)
#undef TM_ATTR
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that contains tcgen05.alloc (it cannot know the column
        // count), but two CTAs that allocate 256 of the 512 columns each do share an SM: measured with a per-SM counter, tools/ubench/occ_tmem.cu.
        // Resident CTAs follow from shared memory (2 x (bytes + 1 KB) <= 228 KB) and registers (<= 128 x 256 x 2), both checked at build / create time.
        h->admm_ctas_per_sm = 2;
        return (int)e;
    }
    h->admm_smem_bytes = (int)admm_smem_bytes(h->tab, h->admm_threads, !small);
    h->admm_ctas_per_sm = small ? 2 : 1;
    if (small) {
        e = cudaFuncSetAttribute(v256::k_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(v256::k_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
    } else {
        e = cudaFuncSetAttribute(v512::k_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(v512::k_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
    }
    return (int)e;
}

void launch_admm(pgn_handle* h) {
    AdmmArgs a;
    a.q = h->qd; a.st = h->st; a.B = h->nv;          // tickets of this launch; `order` holds global vehicle indices
    a.rec = h->d_rec; a.ws_xz = h->d_ws_xz; a.ws_y = h->d_ws_y; a.rho = h->d_rho;
    a.sol_x = h->d_sol_x; a.sol_y = h->d_sol_y;
    a.iters = h->d_iters; a.status = h->d_status; a.rho_updates = h->d_rho_updates; a.pri_res = h->d_pri_res; a.dua_res = h->d_dua_res;
    a.solved = h->d_solved; a.counter = h->d_counter + h->part;
    a.trace = h->profiling == 3 ? h->d_trace : nullptr; a.trace_n = h->d_trace_n; a.trace_cap = h->trace_cap; a.part = h->part;
    a.cycles = h->profiling >= 2 ? h->d_cycles : nullptr;      // 1: stage timers only, 2: + in-kernel cycle counters
    cudaMemsetAsync(h->d_counter + h->part, 0, sizeof(int), h->stream);
    k_admm_order<<<1, 1024, 0, h->stream>>>(h->d_iters + h->v0, h->d_order + h->v0, h->nv, h->v0);
    h->launches++;
    a.order = h->d_order + h->v0;
    a.skip = (h->guard_pause > 0.0 || h->in_callback) ? h->d_skip : nullptr; a.cold = h->d_cold;
    a.hold = h->hold_on ? h->d_hold : nullptr; a.iters_acc = h->d_iters_acc; a.iter_cap = h->hold_on ? h->round_cap : 0;
    for (size_t i = 0; i < h->tab.sol_ph_ptr.size() && i <= ADMM_MAX_PHASES; i++) a.ph_ptr[i] = h->tab.sol_ph_ptr[i];
    const bool small = h->admm_threads == 256 && !h->admm_tmem;
    const int full = h->num_sms * h->admm_ctas_per_sm;
    int grid = full;
    if (grid > h->nv) grid = h->nv;
    // launches of different pipeline parts may overlap: every part owns a slice of the scratch
    a.scratch = h->admm_tmem ? h->d_admm_scratch + (size_t)h->part * full * tm_scratch_doubles(h->tab.Nk, h->tab.nnzA) : nullptr;
    // The pipeline parts run on high-priority streams and the ADMM kernel is launched at the lowest priority: two resident ADMM CTAs hold the
    // whole register file of an SM, so without this the short kernels of the other parts (nodes, linearisation, HJI, propagation) queue behind
    // every pending ADMM CTA instead of slipping into the slot an exiting one frees (pgn_set_pipeline_parts).
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.dynamicSmemBytes = h->admm_smem_bytes; cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority; attr[0].val.priority = h->prio_least;
    cfg.attrs = attr; cfg.numAttrs = h->admm_low_priority ? 1 : 0;
#define ADMM_LAUNCH(ns, nt)                                                                   \
    do {                                                                                      \
        cfg.blockDim = dim3(nt);                                                              \
        if (a.cycles) cudaLaunchKernelEx(&cfg, ns::k_admm<true>, a);                          \
        else cudaLaunchKernelEx(&cfg, ns::k_admm<false>, a);                                  \
    } while (0)
    if (h->admm_tmem) ADMM_LAUNCH(vtm, 256);
    else if (small) ADMM_LAUNCH(v256, 256);
    else ADMM_LAUNCH(v512, 512);
#undef ADMM_LAUNCH
    h->launches++;
}

// graph kernel nodes do not take the priority of the stream they were captured on: ensure_round_graph sets it node by node and asks here
bool admm_is_kernel(const void* f) {
    return f == (const void*)vtm::k_admm<false> || f == (const void*)vtm::k_admm<true> || f == (const void*)v256::k_admm<false> || f == (const void*)v256::k_admm<true> ||
           f == (const void*)v512::k_admm<false> || f == (const void*)v512::k_admm<true>;
}
int admm_orow_fwd(const QpTables& t) { return orow_fwd_count(t); }
size_t admm_scratch_doubles(const pgn_handle* h) {
    return h->admm_tmem ? (size_t)PGN_MAX_PARTS * h->num_sms * 2 * tm_scratch_doubles(h->tab.Nk, h->tab.nnzA) : 0;
}

}  // namespace pgn
