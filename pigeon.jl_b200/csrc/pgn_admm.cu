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
    uint16_t ph_ptr[ADMM_MAX_PHASES + 1];   // first task of every solve phase (forward phases, then backward phases): uniform constant-bank reads
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
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int vec_len(int Nk) { return (Nk + 2) & ~1; }                 // Nk values + the always-zero element Nk
__host__ __device__ inline size_t lval_region_doubles(int nslots, int tail_dim, int Nk, int nnzA) {
    const size_t alias = align_up((size_t)(Nk + 1) * 2 + (size_t)nnzA * 4, 4) + (size_t)nnzA * 4;      // kptr, ke (u16) + arc (u32)
    const size_t a = (alias + 7) / 8, l = (size_t)nslots + (size_t)tail_dim * (tail_dim + 1) / 2;
    return a > l ? a : l;
}

// nthreads: threads of the CTA (the reduction scratch holds 16 doubles per warp); tables_in_smem: whether the static warp programs of the
// solves / factor are copied into shared memory (large QPs, one CTA per SM) or read through L1 (small QPs, two CTAs per SM)
size_t admm_smem_bytes(const QpTables& t, int nthreads, bool tables_in_smem) {
    const size_t V = vec_len(t.Nk);
    size_t d = lval_region_doubles(t.nslots, t.tail_dim, t.Nk, t.nnzA) + V + align_up(t.nnzA, 2) + 7 * V + 16 * (nthreads / 32) + 8;
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
    const bool small = h->admm_threads == 256;
    h->admm_smem_bytes = (int)admm_smem_bytes(h->tab, h->admm_threads, !small);
    cudaError_t e;
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
    a.cycles = h->profiling >= 2 ? h->d_cycles : nullptr;      // 1: stage timers only, 2: + in-kernel cycle counters
    cudaMemsetAsync(h->d_counter + h->part, 0, sizeof(int), h->stream);
    k_admm_order<<<1, 1024, 0, h->stream>>>(h->d_iters + h->v0, h->d_order + h->v0, h->nv, h->v0);
    h->launches++;
    a.order = h->d_order + h->v0;
    a.skip = (h->guard_pause > 0.0 || h->in_callback) ? h->d_skip : nullptr; a.cold = h->d_cold;
    for (size_t i = 0; i < h->tab.sol_ph_ptr.size() && i <= ADMM_MAX_PHASES; i++) a.ph_ptr[i] = h->tab.sol_ph_ptr[i];
    const bool small = h->admm_threads == 256;
    int grid = h->num_sms * (small ? 2 : 1);
    if (grid > h->nv) grid = h->nv;
    if (small) {
        if (a.cycles) v256::k_admm<true><<<grid, 256, h->admm_smem_bytes, h->stream>>>(a);
        else v256::k_admm<false><<<grid, 256, h->admm_smem_bytes, h->stream>>>(a);
    } else {
        if (a.cycles) v512::k_admm<true><<<grid, 512, h->admm_smem_bytes, h->stream>>>(a);
        else v512::k_admm<false><<<grid, 512, h->admm_smem_bytes, h->stream>>>(a);
    }
    h->launches++;
}

}  // namespace pgn
