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

#define NW (ADMM_THREADS / 32)
#define ADMM_NCYC 256
#define SWEEP_EPT ((64 * 65 / 2 + ADMM_THREADS - 1) / ADMM_THREADS)      // elements of the packed dense tail per thread

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

size_t admm_smem_bytes(const QpTables& t) {
    const size_t V = vec_len(t.Nk);
    size_t d = lval_region_doubles(t.nslots, t.tail_dim, t.Nk, t.nnzA) + V + align_up(t.nnzA, 2) + 7 * V + 16 * NW + 8;
    size_t u64 = (t.sol_task.size() + t.fac_task.size() + t.inv_task.size()) / 4;
    size_t u32 = t.bent.size() + t.fac_lvl_ptr.size() + t.inv_lvl_ptr.size() + 4;
    size_t u16 = (size_t)t.nslots + t.sol_orow.size() + 8;
    return d * 8 + u64 * 8 + align_up(u32 * 4, 8) + align_up(u16 * 2, 8) + align_up((size_t)t.Nk, 8) + 64;
}

__device__ __forceinline__ void carve(const QpDev& q, unsigned char* base, Smem& s, uint2*& w_task, uint32_t*& w_u32, uint16_t*& w_u16) {
    const int V = vec_len(q.Nk);
    double* d = reinterpret_cast<double*>(base);
    s.Lval = d; s.S = d + q.nslots; d += lval_region_doubles(q.nslots, q.tail_dim, q.Nk, q.nnzA);
    s.Dinv = d; d += V;
    s.Aval = d; d += (q.nnzA + 1) & ~1;
    s.xz = d; d += V;
    s.sol = d; d += V;
    s.dxy = d; d += V;
    s.yq = d; d += V;
    s.lo = d; d += V;
    s.hi = d; d += V;
    s.sc = d; d += V;
    s.red = d; d += 16 * NW + 8;
    w_task = reinterpret_cast<uint2*>(d);
    s.sol_task = w_task; s.fac_task = s.sol_task + q.n_sol_task; s.inv_task = s.fac_task + q.n_fac_task;
    w_u32 = reinterpret_cast<uint32_t*>(w_task + q.n_sol_task + q.n_fac_task + q.n_inv_task);
    s.bent = w_u32; s.fac_lvl = s.bent + q.n_bent; s.inv_lvl = s.fac_lvl + q.n_fac_lvl + 1;
    size_t off = align_up((size_t)(reinterpret_cast<const unsigned char*>(s.inv_lvl + q.n_inv_levels + 1) - base), 8);
    w_u16 = reinterpret_cast<uint16_t*>(base + off);
    s.fidx = w_u16; s.orow = s.fidx + q.nslots;
    off = align_up((size_t)(reinterpret_cast<const unsigned char*>(s.orow + q.n_orow) - base), 8);
    s.flag = base + off;
    s.kptr = reinterpret_cast<uint16_t*>(s.Lval);
    s.ke = s.kptr + q.Nk + 1;
    s.arc = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.Lval) + align_up((size_t)(q.Nk + 1) * 2 + (size_t)q.nnzA * 4, 4));
}

// warp-level butterfly over all 32 lanes
template <bool IS_MAX>
__device__ __forceinline__ double warp_all(double a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double b = __shfl_xor_sync(0xffffffffu, a, o);
        a = IS_MAX ? fmax(a, b) : a + b;
    }
    return a;
}
// block-wide max / sum of NV values per thread; every thread returns with the results in v[]
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_reduce(double* v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_all<IS_MAX>(v[k]);
    __syncthreads();   // red[] free
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) red[k * NW + w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_all<IS_MAX>(lane < NW ? red[k * NW + lane] : (IS_MAX ? -1e300 : 0.0));
}
// block-wide (sum, max) pair in one pass
__device__ __forceinline__ void block_reduce_sum_max(double& sum, double& mx, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sum = warp_all<false>(sum); mx = warp_all<true>(mx);
    __syncthreads();
    if (lane == 0) { red[w] = sum; red[NW + w] = mx; }
    __syncthreads();
    sum = warp_all<false>(lane < NW ? red[lane] : 0.0);
    mx = warp_all<true>(lane < NW ? red[NW + lane] : -1e300);
}

__device__ __forceinline__ double limit_scaling(double a) {
    a = a < 1e-4 ? 1.0 : a;
    return a > 1e4 ? 1e4 : a;
}
__device__ __forceinline__ double rho_of(uint8_t flag, double rho) { return flag == 2 ? 1e3 * rho : (flag == 3 ? 1e-6 : rho); }
// rho_inv_vec of OSQP: reciprocals are formed once per rho value and multiplied in
struct RhoInv { double in, eq, loose; };
__device__ __forceinline__ RhoInv make_rho_inv(double rho) { RhoInv r; r.in = 1.0 / rho; r.eq = 1.0 / (1e3 * rho); r.loose = 1.0 / 1e-6; return r; }
__device__ __forceinline__ double rinv_of(uint8_t flag, const RhoInv& r) { return flag == 2 ? r.eq : (flag == 3 ? r.loose : r.in); }

// sum over the 2^sh adjacent lanes of a group (sh warp-uniform); every lane of the warp must call.  A plain loop: a switch here becomes
// a jump table (LDC + BRX), which costs a lone warp far more than the shuffles themselves.
__device__ __forceinline__ double group_sum_sh(double v, int sh) {
#pragma unroll 1
    for (int o = (1 << sh) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_sum_c(double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp task of a gather program (factorisation / range inverse): returns, in the lanes with sub == 0, the sum over the row's
// entries of W[a] * W[b] / d_k.  Entries stream from global memory (static, coalesced: slot k of lane l is at ebase + 32 k + l) in batches
// of four; the last (partial) batch reads up to three slots past the task — always inside the padded array — and replaces them by the
// zero entry, so the task runs without data-dependent branches.
#define GATHER_TERM(e) (s.Lval[(e) & 0xffff] * s.Lval[((e) >> 16) & 0xffff] * s.Dinv[(e) >> 32])
__device__ __forceinline__ double gather_task(const Smem& s, const uint2 d, const unsigned long long* __restrict__ ents, int lane, unsigned long long pad) {
    const int K = d.y & 0xff, sh = (d.y >> 16) & 0xff;
    const unsigned long long* e = ents + ((size_t)(d.x & 0xffff) << 5) + lane;
    double acc0 = 0.0, acc1 = 0.0;
    int k = 0;
#pragma unroll 1
    for (; k + 4 <= K; k += 4, e += 128) {
        const unsigned long long e0 = __ldg(e), e1 = __ldg(e + 32), e2 = __ldg(e + 64), e3 = __ldg(e + 96);
        const double t0 = GATHER_TERM(e0), t1 = GATHER_TERM(e1), t2 = GATHER_TERM(e2), t3 = GATHER_TERM(e3);
        acc0 += t0; acc1 += t1; acc0 += t2; acc1 += t3;
    }
    {
        const int rem = K - k;
        unsigned long long e0 = __ldg(e), e1 = __ldg(e + 32), e2 = __ldg(e + 64);
        e0 = rem > 0 ? e0 : pad; e1 = rem > 1 ? e1 : pad; e2 = rem > 2 ? e2 : pad;
        const double t0 = GATHER_TERM(e0), t1 = GATHER_TERM(e1), t2 = GATHER_TERM(e2);
        acc0 += t0; acc1 += t1; acc0 += t2;
    }
    return group_sum_sh(acc0 + acc1, sh);
}

// Same with the first four entries and the row's target already in registers (loaded before the previous level's barrier).  The second
// and third batch (K <= 12 covers every task but the Schur gather of the dense tail) are requested together before the first is
// consumed, so a task pays at most one exposed L2 round trip; the branches are warp-uniform.
#define GATHER_BATCH(b0, b1, b2, b3, kbase)                                                                            \
    do {                                                                                                              \
        const unsigned long long q0__ = (kbase) < K ? (b0) : pad, q1__ = (kbase) + 1 < K ? (b1) : pad;                  \
        const unsigned long long q2__ = (kbase) + 2 < K ? (b2) : pad, q3__ = (kbase) + 3 < K ? (b3) : pad;              \
        const double t0__ = GATHER_TERM(q0__), t1__ = GATHER_TERM(q1__), t2__ = GATHER_TERM(q2__), t3__ = GATHER_TERM(q3__); \
        acc0 += t0__; acc1 += t1__; acc0 += t2__; acc1 += t3__;                                                         \
    } while (0)
__device__ __forceinline__ double gather_task_pre(const Smem& s, const uint2 d, const unsigned long long* __restrict__ ents, int lane, unsigned long long pad,
                                                  unsigned long long p0, unsigned long long p1, unsigned long long p2, unsigned long long p3) {
    const int K = d.y & 0xff, sh = (d.y >> 16) & 0xff;
    const unsigned long long* e = ents + ((size_t)(d.x & 0xffff) << 5) + lane + 128;
    double acc0 = 0.0, acc1 = 0.0;
    if (K > 4) {
        const unsigned long long a0 = __ldg(e), a1 = __ldg(e + 32), a2 = __ldg(e + 64), a3 = __ldg(e + 96);
        unsigned long long b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        if (K > 8) { b0 = __ldg(e + 128); b1 = __ldg(e + 160); b2 = __ldg(e + 192); b3 = __ldg(e + 224); }
        GATHER_BATCH(p0, p1, p2, p3, 0);
        GATHER_BATCH(a0, a1, a2, a3, 4);
        if (K > 8) {
            GATHER_BATCH(b0, b1, b2, b3, 8);
            e += 256;
#pragma unroll 1
            for (int k = 12; k < K; k += 4, e += 128) {
                const unsigned long long c0 = __ldg(e), c1 = __ldg(e + 32), c2 = __ldg(e + 64), c3 = __ldg(e + 96);
                GATHER_BATCH(c0, c1, c2, c3, k);
            }
        }
    } else {
        GATHER_BATCH(p0, p1, p2, p3, 0);
    }
    return group_sum_sh(acc0 + acc1, sh);
}

// numeric LDL' of K = [P + sigma I, A'; A, -1/rho] (position space) in the unscaled form W = L D with the static gather programs:
//   d_j = K_jj - sum_k W_jk^2 / d_k,     W_ij = K_ij - sum_k W_ik W_jk / d_k        (one pass and one barrier per level)
// then the explicit inverses of the level ranges (in place) and of the dense tail.
#define FAC_T(idx)                                                                         \
    do {                                                                                   \
        if (PROF && lvl_cyc && threadIdx.x == 0 && blockIdx.x == 0) {                              \
            const long long now__ = clock64();                                             \
            lvl_cyc[(idx)] += (unsigned int)(now__ - t_lvl);                               \
            t_lvl = now__;                                                                 \
        }                                                                                  \
    } while (0)
template <bool PROF>
__device__ void factor(const QpDev& q, const Smem& s, double sigma, double rho, unsigned int* lvl_cyc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_lvl = clock64();
    const RhoInv ri = make_rho_inv(rho);
    const unsigned long long pad = (unsigned long long)q.zslot | ((unsigned long long)q.zslot << 16);      // 0 * 0 / d_0
    const int ts = q.tail_start, Dm = q.tail_dim, npk = Dm * (Dm + 1) / 2;
    for (int e = tid; e < q.nslots + npk; e += ADMM_THREADS) s.Lval[e] = 0.0;
    __syncthreads();
    // Dinv[p] holds K_pp until the pivot of p is formed; tail positions put K_pp on the diagonal of the dense block
    for (int p = tid; p < q.Nk; p += ADMM_THREADS) {
        const double kpp = s.flag[p] ? -rinv_of(s.flag[p], ri) : s.lo[p] + sigma;
        if (p >= ts && Dm > 0) { const int i = p - ts; s.S[i * (i + 1) / 2 + i] = kpp; }
        else s.Dinv[p] = kpp;
    }
    for (int e = tid; e < q.nnzA; e += ADMM_THREADS) s.Lval[__ldg(q.a_slot + e)] = s.Aval[e];
    __syncthreads();
    FAC_T(110);
    // The static program of a level (task descriptor, the row's target, the first four entries: all in L2, ~300 cycles away) does not
    // depend on the numbers: every warp fetches its first task of level l+1 before the barrier of level l, so the round trip overlaps
    // the barrier wait instead of starting the level.
    uint2 dn = make_uint2(0, 0);
    unsigned long long pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0;
    uint32_t tgn = 0;
    bool have = false;
#define FAC_PREFETCH(lvl)                                                                                   \
    do {                                                                                                    \
        const int t__ = s.fac_lvl[lvl] + warp;                                                              \
        have = t__ < (int)s.fac_lvl[(lvl) + 1];                                                             \
        if (have) {                                                                                         \
            dn = s.fac_task[t__];                                                                           \
            const unsigned long long* e__ = q.fac_ent + ((size_t)(dn.x & 0xffff) << 5) + lane;              \
            pe0 = __ldg(e__); pe1 = __ldg(e__ + 32); pe2 = __ldg(e__ + 64); pe3 = __ldg(e__ + 96);          \
            const int sh__ = (dn.y >> 16) & 0xff, rr__ = lane >> sh__;                                      \
            tgn = __ldg(q.fac_tgt + (dn.x >> 16) + min(rr__, (int)((dn.y >> 8) & 0xff) - 1));               \
        }                                                                                                   \
    } while (0)
    if (q.n_fac_lvl > 0) FAC_PREFETCH(0);
    for (int l = 0; l < q.n_fac_lvl; l++) {
        const int t1 = s.fac_lvl[l + 1];
        if (have) {
            const uint2 d = dn;
            const int sh = (d.y >> 16) & 0xff, rr = lane >> sh;
            const bool writer = (lane & ((1 << sh) - 1)) == 0 && rr < (int)((d.y >> 8) & 0xff);
            const uint32_t tg = tgn;
            const double acc = gather_task_pre(s, d, q.fac_ent, lane, pad, pe0, pe1, pe2, pe3);
            if (writer) {
                if (tg & FAC_TGT_PIVOT) { const int j = tg & 0x7fffffff; s.Dinv[j] = 1.0 / (s.Dinv[j] - acc); }
                else s.Lval[tg] -= acc;
            }
        }
        for (int t = s.fac_lvl[l] + warp + NW; t < t1; t += NW) {
            const uint2 d = s.fac_task[t];
            const int sh = (d.y >> 16) & 0xff, rr = lane >> sh;
            const bool writer = (lane & ((1 << sh) - 1)) == 0 && rr < (int)((d.y >> 8) & 0xff);
            uint32_t tg = 0;
            if (writer) tg = __ldg(q.fac_tgt + (d.x >> 16) + rr);          // issued before the gather: its latency overlaps the entry stream
            const double acc = gather_task(s, d, q.fac_ent, lane, pad);
            if (writer) {
                if (tg & FAC_TGT_PIVOT) { const int j = tg & 0x7fffffff; s.Dinv[j] = 1.0 / (s.Dinv[j] - acc); }
                else s.Lval[tg] -= acc;
            }
        }
        if (l + 1 < q.n_fac_lvl) FAC_PREFETCH(l + 1); else have = false;
        __syncthreads();
        FAC_T(120 + (l < 100 ? l : 99));
    }
#undef FAC_PREFETCH
    // level ranges: replace the in-range block of W by the explicit inverse M of the unit lower block L[range, range], level by level:
    //   M_ij = -(W_ij / d_j + sum_{j<k<i} W_ik / d_k M_kj)     (targets of one level are computed into registers before any is written)
    // (the first task of the next level is fetched before the barriers of the current one, as in the factor levels)
    uint2 idn = make_uint2(0, 0);
    unsigned long long ie0 = 0, ie1 = 0, ie2 = 0, ie3 = 0;
    uint32_t itg = 0;
    bool ihave = false;
#define INV_PREFETCH(lvl)                                                                                   \
    do {                                                                                                    \
        const int t__ = s.inv_lvl[lvl] + warp;                                                              \
        ihave = t__ < (int)s.inv_lvl[(lvl) + 1];                                                            \
        if (ihave) {                                                                                        \
            idn = s.inv_task[t__];                                                                          \
            const unsigned long long* e__ = q.inv_ent + ((size_t)(idn.x & 0xffff) << 5) + lane;             \
            ie0 = __ldg(e__); ie1 = __ldg(e__ + 32); ie2 = __ldg(e__ + 64); ie3 = __ldg(e__ + 96);          \
            const int sh__ = (idn.y >> 16) & 0xff, rr__ = lane >> sh__;                                     \
            itg = __ldg(q.inv_tgt + (idn.x >> 16) + min(rr__, (int)((idn.y >> 8) & 0xff) - 1));             \
        }                                                                                                   \
    } while (0)
    if (q.n_inv_levels > 0) INV_PREFETCH(0);
    for (int l = 0; l < q.n_inv_levels; l++) {
        const int t0 = s.inv_lvl[l] + warp, t1 = s.inv_lvl[l + 1];
        double v[INV_MAX_TASKS_PER_WARP];
        int id[INV_MAX_TASKS_PER_WARP];
#pragma unroll
        for (int k = 0; k < INV_MAX_TASKS_PER_WARP; k++) {
            const int t = t0 + k * NW;
            id[k] = -1;
            if (t < t1) {
                const uint2 d = k == 0 ? idn : s.inv_task[t];
                const int sh = (d.y >> 16) & 0xff, rr = lane >> sh;
                const bool writer = (lane & ((1 << sh) - 1)) == 0 && rr < (int)((d.y >> 8) & 0xff);
                uint32_t tg = itg;
                if (k > 0 && writer) tg = __ldg(q.inv_tgt + (d.x >> 16) + rr);
                const double acc = k == 0 ? gather_task_pre(s, d, q.inv_ent, lane, pad, ie0, ie1, ie2, ie3) : gather_task(s, d, q.inv_ent, lane, pad);
                if (writer) {
                    id[k] = tg & 0xffff;
                    v[k] = -(s.Lval[id[k]] * s.Dinv[tg >> 16] + acc);
                }
            }
        }
        if (l + 1 < q.n_inv_levels) INV_PREFETCH(l + 1); else ihave = false;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < INV_MAX_TASKS_PER_WARP; k++)
            if (id[k] >= 0) s.Lval[id[k]] = v[k];
        __syncthreads();
    }
#undef INV_PREFETCH
    FAC_T(111);
    // dense tail: symmetric sweep of the packed lower Schur complement S over all pivots, in place:  S <- -S^-1.
    // Pivot p: S_ik -= S_ip S_kp / d,  S_ip <- S_ip / d,  S_pp <- -1 / d.  The pivot column (and 1/d) of step p+1 is staged into a small
    // buffer by the threads that produce it during step p, so one barrier per pivot suffices.
    if (Dm > 0) {
        double* col = s.red;                 // [2][64] pivot columns + [2] pivots
        double* dpiv = s.red + 128;
        const int ept = (npk + ADMM_THREADS - 1) / ADMM_THREADS;
        int ei[SWEEP_EPT], ek[SWEEP_EPT];
#pragma unroll
        for (int x = 0; x < SWEEP_EPT; x++) {
            const int e = min(tid + x * ADMM_THREADS, npk - 1);          // surplus threads shadow the last element (their stores are masked)
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            i += ((i + 1) * (i + 2) / 2 <= e);
            i -= (i * (i + 1) / 2 > e);
            ei[x] = i; ek[x] = e - i * (i + 1) / 2;
        }
        if (tid < Dm) col[tid] = s.S[tid * (tid + 1) / 2];
        if (tid == 0) dpiv[0] = s.S[0];
        __syncthreads();
        for (int p = 0; p < Dm; p++) {
            const double* cc = col + (p & 1) * 64;
            double* cn = col + ((p + 1) & 1) * 64;
            const double dinv = 1.0 / dpiv[p & 1];           // every thread forms the reciprocal itself: no serial hand-off through one lane
#pragma unroll
            for (int x = 0; x < SWEEP_EPT; x++) {
                if (x < ept) {                                // uniform
                    const int i = ei[x], k = ek[x], e = tid + x * ADMM_THREADS;
                    const bool live = e < npk;
                    const double ci = cc[i], ck = cc[k], old = s.S[live ? e : 0];
                    const double upd = old - ci * ck * dinv;
                    const double onrow = (k == p) ? -dinv : ck * dinv;       // i == p
                    const double v = (i == p) ? onrow : ((k == p) ? ci * dinv : upd);
                    if (live) s.S[e] = v;
                    if (live && k == p + 1) cn[i] = v;                        // column p+1 below (and on) the diagonal ...
                    if (live && i == p + 1) cn[k] = v;                        // ... and its mirror image left of the diagonal
                    if (live && i == p + 1 && k == p + 1) dpiv[(p + 1) & 1] = v;
                }
            }
            __syncthreads();
        }
    }
    FAC_T(112);
}

// One phase of a triangular solve: warp w runs tasks t0 + w, t0 + w + NW, ... < t1,   out[r] = f(in[r], sum_e W_e * in[c_e])  with the
// phase's flags FL known at compile time (pgn_structure.h).  Forward phases read the L values in slot order (conflict-free); backward
// phases gather them through (slot, source) pairs.  Loads are issued in batches of four so that their latencies overlap.
#ifdef PGN_PHASE_PROBE
#define PROBE(i) do { if (probe && threadIdx.x == 0) { const long long n__ = clock64(); probe[i] += (unsigned int)(n__ - tp); tp = n__; } } while (0)
#else
#define PROBE(i)
#endif
template <int FL, bool BWD>
__device__ __forceinline__ void run_phase(const Smem& s, int t0, int t1, int zidx, uint32_t zpair, unsigned int* probe = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PGN_PHASE_PROBE
    long long tp = clock64();
#endif
    const double* __restrict__ in = (FL & TASK_SRC_TMP) ? s.dxy : s.sol;
    double* __restrict__ out = (FL & TASK_DST_TMP) ? s.dxy : s.sol;
    for (int t = t0 + warp; t < t1; t += NW) {
        // straight-line on purpose: with one or two live warps per scheduler every branch costs a full resolve latency
        const uint2 d = s.sol_task[t];
        PROBE(0);
        const int K = d.y & 0xff, nrows = (d.y >> 8) & 0xff, sh = (d.y >> 16) & 0xff;
        const int rr = lane >> sh;
        const bool writer = ((lane & ((1 << sh) - 1)) == 0) & (rr < nrows);
        const int r = s.orow[(d.x >> 16) + min(rr, nrows - 1)];
        double x = in[r];
        const double di = (FL & (TASK_SCALE_ACC | TASK_SCALE_OUT)) ? s.Dinv[r] : 0.0;
        const int e = ((d.x & 0xffff) << 5) + lane;
        double acc0 = 0.0, acc1 = 0.0;
        // full batches of four with constant offsets, then one partial batch that reads up to three slots past the task (always inside
        // the padded arrays) and redirects them to the zero entries: no data-dependent branches
        int k = 0;
        if (!BWD) {
            const double* lv = s.Lval + e;
            const uint16_t* ix = s.fidx + e;
#pragma unroll 1
            for (; k + 4 <= K; k += 4, lv += 128, ix += 128) {
                const int i0 = ix[0], i1 = ix[32], i2 = ix[64], i3 = ix[96];
                const double l0 = lv[0], l1 = lv[32], l2 = lv[64], l3 = lv[96];
                const double x0 = in[i0], x1 = in[i1], x2 = in[i2], x3 = in[i3];
                acc0 += l0 * x0; acc1 += l1 * x1; acc0 += l2 * x2; acc1 += l3 * x3;
            }
            const int rem = K - k;
            int i0 = ix[0], i1 = ix[32], i2 = ix[64];
            double l0 = lv[0], l1 = lv[32], l2 = lv[64];
            i0 = rem > 0 ? i0 : zidx; i1 = rem > 1 ? i1 : zidx; i2 = rem > 2 ? i2 : zidx;
            l0 = rem > 0 ? l0 : 0.0; l1 = rem > 1 ? l1 : 0.0; l2 = rem > 2 ? l2 : 0.0;
            const double x0 = in[i0], x1 = in[i1], x2 = in[i2];
            acc0 += l0 * x0; acc1 += l1 * x1; acc0 += l2 * x2;
        } else {
            const uint32_t* be = s.bent + e;
#pragma unroll 1
            for (; k + 4 <= K; k += 4, be += 128) {
                const uint32_t b0 = be[0], b1 = be[32], b2 = be[64], b3 = be[96];
                const double l0 = s.Lval[b0 & 0xffff], l1 = s.Lval[b1 & 0xffff], l2 = s.Lval[b2 & 0xffff], l3 = s.Lval[b3 & 0xffff];
                const double x0 = in[b0 >> 16], x1 = in[b1 >> 16], x2 = in[b2 >> 16], x3 = in[b3 >> 16];
                acc0 += l0 * x0; acc1 += l1 * x1; acc0 += l2 * x2; acc1 += l3 * x3;
            }
            const int rem = K - k;
            uint32_t b0 = be[0], b1 = be[32], b2 = be[64];
            b0 = rem > 0 ? b0 : zpair; b1 = rem > 1 ? b1 : zpair; b2 = rem > 2 ? b2 : zpair;
            const double l0 = s.Lval[b0 & 0xffff], l1 = s.Lval[b1 & 0xffff], l2 = s.Lval[b2 & 0xffff];
            const double x0 = in[b0 >> 16], x1 = in[b1 >> 16], x2 = in[b2 >> 16];
            acc0 += l0 * x0; acc1 += l1 * x1; acc0 += l2 * x2;
        }
        double acc = acc0 + acc1;
        PROBE(1);
        acc = group_sum_sh(acc, sh);
        PROBE(2);
        if (FL & TASK_SCALE_ACC) acc *= di;
        x = (FL & TASK_ADD) ? x + acc : x - acc;
        if (FL & TASK_SCALE_OUT) x *= di;
        if (writer) out[r] = x;
        PROBE(3);
    }
    PROBE(4);
    __syncthreads();
    PROBE(5);
}

#define LVL_T(idx)                                                                         \
    do {                                                                                   \
        if (PROF && lvl_cyc && threadIdx.x == 0 && blockIdx.x == 0) {                              \
            const long long now__ = clock64();                                             \
            lvl_cyc[(idx)] += (unsigned int)(now__ - t_lvl);                               \
            t_lvl = now__;                                                                 \
        }                                                                                  \
    } while (0)
// sol <- K^-1 rhs (rhs in sol, except the first range whose rhs is in the scratch vector):  forward over the level ranges (first range:
// in-range explicit inverse only; others: external part, then in-range inverse), the dense tail (external part as the last forward
// phase, then one symmetric mat-vec with -S^-1), and the mirror image backwards.
template <bool PROF>
__device__ __forceinline__ void kkt_solve(const AdmmArgs& a, const Smem& s, unsigned int* lvl_cyc) {
    const QpDev& q = a.q;
    const int tid = threadIdx.x;
    long long t_lvl = clock64();
    const int zidx = q.Nk;                                                    // the always-zero vector element
    const uint32_t zpair = (uint32_t)q.zslot | ((uint32_t)q.Nk << 16);       // (always-zero L slot, always-zero vector element)
    for (int ph = 0; ph < q.n_fwd_ph; ph++) {
        if (ph & 1) run_phase<TASK_DST_TMP, false>(s, a.ph_ptr[ph], a.ph_ptr[ph + 1], zidx, zpair, nullptr);      // t = b - W_ext y^
        else run_phase<TASK_SRC_TMP | TASK_ADD | TASK_SCALE_OUT, false>(s, a.ph_ptr[ph], a.ph_ptr[ph + 1], zidx, zpair);              // y^ = (t + M t) / d
        LVL_T(ph);
    }
    const int ts = q.tail_start, Dm = q.tail_dim;
    if (Dm > 0) {
        // x_T = S^-1 t_T with the packed lower -S^-1: 8 lanes per row
        for (int i8 = tid; i8 < ((Dm * 8 + 31) & ~31); i8 += ADMM_THREADS) {
            const int i = i8 >> 3, sub = i8 & 7;
            double acc0 = 0.0, acc1 = 0.0;
            {
                const double* tv = s.dxy + ts;
                const int ii = min(i, Dm - 1), rb = ii * (ii + 1) / 2;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {             // tail_dim <= 64: eight columns per lane, all loads in flight together
                    const int k0 = sub + 8 * j, k1 = k0 + 8;
                    const int c0 = min(k0, Dm - 1), c1 = min(k1, Dm - 1);
                    const double a0 = s.S[c0 <= ii ? rb + c0 : c0 * (c0 + 1) / 2 + ii], a1 = s.S[c1 <= ii ? rb + c1 : c1 * (c1 + 1) / 2 + ii];
                    const double t0 = tv[c0], t1 = tv[c1];
                    acc0 += (k0 < Dm ? a0 : 0.0) * t0; acc1 += (k1 < Dm ? a1 : 0.0) * t1;
                }
            }
            double acc = acc0 + acc1;
            acc = group_sum_c<8>(acc);
            if (i < Dm && sub == 0) s.sol[ts + i] = -acc;
        }
        __syncthreads();
        LVL_T(100);
    }
    for (int ph = 0; ph < q.n_bwd_ph; ph++) {
        const int pp = q.n_fwd_ph + ph;
        if (ph & 1) run_phase<TASK_SRC_TMP | TASK_ADD, true>(s, a.ph_ptr[pp], a.ph_ptr[pp + 1], zidx, zpair);                         // x = v + M' v
        else run_phase<TASK_DST_TMP | TASK_SCALE_ACC, true>(s, a.ph_ptr[pp], a.ph_ptr[pp + 1], zidx, zpair);                          // v = y^ - (W_below' x) / d
        LVL_T(pp);
    }
}

// out[p] = sum over the off-diagonal KKT entries of row p:  constraints get (A x)_i, variables get (A' y)_j
__device__ __forceinline__ void kadj_product(const QpDev& q, const Smem& s, const double* vin_var, const double* vin_con, double* out) {
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS) {
        const bool con = s.flag[p] != 0;
        const double* in = con ? vin_var : vin_con;
        double acc = 0.0;
        const int e1 = __ldg(q.kadj_ptr + p + 1);
        for (int e = __ldg(q.kadj_ptr + p); e < e1; e++) acc += s.Aval[__ldg(q.kadj_e + e)] * in[__ldg(q.kadj_nb + e)];
        out[p] = acc;
    }
}

struct Resid { double pri_res, dua_res, eps_pri_n, eps_dua_n, s_pri, s_dua, s_pn, s_dn; };

// residuals in unscaled norms (termination) and scaled norms (rho estimate). s.sol receives [A'y ; Ax] by position.
__device__ __forceinline__ Resid residuals(const QpDev& q, const Smem& s, double cinv) {
    kadj_product(q, s, s.xz, s.yq, s.sol);
    double v[14];
#pragma unroll
    for (int k = 0; k < 14; k++) v[k] = 0.0;
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS) {
        const double t = s.sol[p], sc = s.sc[p], isc = 1.0 / sc;
        if (s.flag[p]) {
            const double z = s.xz[p], r = t - z;
            v[0] = fmax(v[0], fabs(r * isc)); v[1] = fmax(v[1], fabs(z * isc)); v[2] = fmax(v[2], fabs(t * isc));
            v[7] = fmax(v[7], fabs(r)); v[8] = fmax(v[8], fabs(z)); v[9] = fmax(v[9], fabs(t));
        } else {
            const double px = s.lo[p] * s.xz[p], qq = s.yq[p], r = px + qq + t;
            v[3] = fmax(v[3], fabs(r * isc)); v[4] = fmax(v[4], fabs(px * isc)); v[5] = fmax(v[5], fabs(t * isc)); v[6] = fmax(v[6], fabs(qq * isc));
            v[10] = fmax(v[10], fabs(r)); v[11] = fmax(v[11], fabs(px)); v[12] = fmax(v[12], fabs(t)); v[13] = fmax(v[13], fabs(qq));
        }
    }
    block_reduce<14, true>(v, s.red);
    Resid R;
    R.pri_res = v[0]; R.eps_pri_n = fmax(v[1], v[2]);
    R.dua_res = cinv * v[3]; R.eps_dua_n = cinv * fmax(v[6], fmax(v[5], v[4]));
    R.s_pri = v[7]; R.s_pn = fmax(v[8], v[9]);
    R.s_dua = v[10]; R.s_dn = fmax(v[13], fmax(v[12], v[11]));
    return R;
}

// is_primal_infeasible / is_dual_infeasible of OSQP on the increments stored in s.dxy (delta_x at variables, delta_y at constraints)
__device__ bool primal_infeasible(const QpDev& q, const Smem& s, double eps) {
    const double thr = OSQP_INFTY * 1e-4;
    double v[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (s.flag[p]) {
            double dy = s.dxy[p];
            const double l = s.lo[p], u = s.hi[p];
            if (u > thr) { dy = (l < -thr) ? 0.0 : fmin(dy, 0.0); }
            else if (l < -thr) dy = fmax(dy, 0.0);
            s.dxy[p] = dy;
            v[0] = fmax(v[0], fabs(s.sc[p] * dy));
        }
    block_reduce<1, true>(v, s.red);
    const double norm_dy = v[0];
    if (!(norm_dy > eps)) return false;
    double w[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (s.flag[p]) { const double dy = s.dxy[p]; w[0] += s.hi[p] * fmax(dy, 0.0) + s.lo[p] * fmin(dy, 0.0); }
    block_reduce<1, false>(w, s.red);
    if (!(w[0] < -eps * norm_dy)) return false;
    // ||Dinv A' dy||
    kadj_product(q, s, s.dxy, s.dxy, s.sol);
    double n2[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (!s.flag[p]) n2[0] = fmax(n2[0], fabs(s.sol[p] / s.sc[p]));
    block_reduce<1, true>(n2, s.red);
    return n2[0] < eps * norm_dy;
}
__device__ bool dual_infeasible(const QpDev& q, const Smem& s, double eps, double c) {
    const double thr = OSQP_INFTY * 1e-4;
    double v[2] = {0.0, 0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (!s.flag[p]) { v[0] = fmax(v[0], fabs(s.sc[p] * s.dxy[p])); }
    block_reduce<1, true>(v, s.red);
    const double norm_dx = v[0];
    if (!(norm_dx > eps)) return false;
    double w[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (!s.flag[p]) w[0] += s.yq[p] * s.dxy[p];
    block_reduce<1, false>(w, s.red);
    if (!(w[0] < -c * eps * norm_dx)) return false;
    double n2[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (!s.flag[p]) n2[0] = fmax(n2[0], fabs(s.lo[p] * s.dxy[p] / s.sc[p]));
    block_reduce<1, true>(n2, s.red);
    if (!(n2[0] < c * eps * norm_dx)) return false;
    kadj_product(q, s, s.dxy, s.dxy, s.sol);   // constraints: (A dx)_i
    __syncthreads();
    double bad[1] = {0.0};
    for (int p = threadIdx.x; p < q.Nk; p += ADMM_THREADS)
        if (s.flag[p]) {
            const double a = s.sol[p] / s.sc[p];
            if ((s.hi[p] < thr && a > eps * norm_dx) || (s.lo[p] > -thr && a < -eps * norm_dx)) bad[0] = 1.0;
        }
    block_reduce<1, true>(bad, s.red);
    return bad[0] == 0.0;
}

#define PHASE(idx)                                                                  \
    do {                                                                            \
        if (PROF && a.cycles && threadIdx.x == 0) {                                         \
            const long long now__ = clock64();                                      \
            s_cyc[(idx)] += (unsigned int)(now__ - t_phase);                        \
            t_phase = now__;                                                        \
        }                                                                           \
    } while (0)

// PROF: in-kernel cycle counters (profiling level 2).  A template parameter because even an untaken counter predicate (parameter load +
// blockIdx read + compare) costs ~50 cycles per barrier interval (tools/ubench/sreg.cu).
template <bool PROF>
__global__ void __launch_bounds__(ADMM_THREADS, 1) k_admm(const AdmmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_vehicle;
    __shared__ unsigned int s_cyc[ADMM_NCYC];      // profiling only: cycle counters kept on chip, flushed once when the CTA retires
    const QpDev& q = a.q;
    const AdmmSettings& st = a.st;
    Smem s;
    const int tid = threadIdx.x;
    {   // static tables of the warp programs: global -> shared, once per CTA
        uint2* w_task; uint32_t* w_u32; uint16_t* w_u16;
        carve(q, smem_raw, s, w_task, w_u32, w_u16);
        for (int i = tid; i < q.n_sol_task; i += ADMM_THREADS) w_task[i] = q.sol_task[i];
        for (int i = tid; i < q.n_fac_task; i += ADMM_THREADS) w_task[q.n_sol_task + i] = q.fac_task[i];
        for (int i = tid; i < q.n_inv_task; i += ADMM_THREADS) w_task[q.n_sol_task + q.n_fac_task + i] = q.inv_task[i];
        for (int i = tid; i < q.n_bent; i += ADMM_THREADS) w_u32[i] = q.bent[i];
        for (int i = tid; i <= q.n_fac_lvl; i += ADMM_THREADS) w_u32[q.n_bent + i] = q.fac_lvl_ptr[i];
        for (int i = tid; i <= q.n_inv_levels; i += ADMM_THREADS) w_u32[q.n_bent + q.n_fac_lvl + 1 + i] = q.inv_lvl_ptr[i];
        for (int i = tid; i < q.nslots; i += ADMM_THREADS) w_u16[i] = q.fidx[i];
        for (int i = tid; i < q.n_orow; i += ADMM_THREADS) w_u16[q.nslots + i] = q.sol_orow[i];
        if (tid == 0) { s.sol[q.Nk] = 0.0; s.dxy[q.Nk] = 0.0; }      // the always-zero element read by the padding entries
    }
    for (int i = threadIdx.x; i < ADMM_NCYC; i += ADMM_THREADS) s_cyc[i] = 0;
    __syncthreads();
    long long t_phase = clock64();
    for (;;) {
        if (tid == 0) s_vehicle = atomicAdd(a.counter, 1);
        __syncthreads();
        const int ticket = s_vehicle;
        __syncthreads();
        if (ticket >= a.B) break;
        const int v = a.order[ticket];
        if (a.skip && a.skip[v]) {                                  // CTA-uniform: every thread reads the same flag
            if (tid == 0) { a.iters[v] = 0; a.status[v] = PGN_QP_UNSOLVED; }
            continue;
        }
        const bool warm = st.warm_start && !a.cold[v];
        const double* rec = a.rec + (size_t)v * q.rec_len;
        PHASE(7);

        // ---- 1. gather the QP values --------------------------------------------------------------------------------
        for (int e = tid; e < q.nnzA; e += ADMM_THREADS) {
            const int src = __ldg(q.a_src + e);
            s.Aval[e] = src >= 0 ? rec[src] : (src == -1 ? 1.0 : -1.0);
            s.arc[e] = __ldg(q.a_rc + e);
        }
        for (int e = tid; e < 2 * q.nnzA; e += ADMM_THREADS) s.ke[e] = __ldg(q.kadj_e + e);
        for (int p = tid; p <= q.Nk; p += ADMM_THREADS) s.kptr[p] = __ldg(q.kadj_ptr + p);
        for (int p = tid; p < q.Nk; p += ADMM_THREADS) {
            const int idx = __ldg(q.pos2idx + p);
            if (__ldg(q.is_con + p)) {
                double b[2];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int ty = k == 0 ? __ldg(q.l_type + idx) : __ldg(q.u_type + idx);
                    const int ix = k == 0 ? __ldg(q.l_idx + idx) : __ldg(q.u_idx + idx);
                    double val;
                    if (ty == BND_CONST) val = __ldg(q.ctab + ix);
                    else if (ty == BND_REC) val = rec[ix];
                    else if (ty == BND_NEG_REC) val = -rec[ix];
                    else if (ty == BND_DT_SCALED) val = __ldg(q.ctab + CT_DDELTA_N) * rec[ix];
                    else val = -__ldg(q.ctab + CT_DDELTA_N) * rec[ix];
                    b[k] = val;
                }
                s.lo[p] = fmax(b[0], -OSQP_INFTY);
                s.hi[p] = fmin(b[1], OSQP_INFTY);
                s.flag[p] = 1;
                s.yq[p] = warm ? a.ws_y[(size_t)v * q.Nk + p] : 0.0;
            } else {
                const int pm = __ldg(q.P_mode + idx), qm = __ldg(q.q_mode + idx);
                double Pv = 0.0, qv = 0.0;
                if (pm == PQ_TIMES_DT) Pv = 2.0 * __ldg(q.wtab + __ldg(q.P_w + idx)) * rec[__ldg(q.P_t + idx)];
                else if (pm == PQ_OVER_DT) Pv = 2.0 * __ldg(q.wtab + __ldg(q.P_w + idx)) / rec[__ldg(q.P_t + idx)];
                if (qm == PQ_TIMES_DT) qv = __ldg(q.wtab + __ldg(q.q_w + idx)) * rec[__ldg(q.q_t + idx)];
                else if (qm == PQ_CONST) qv = (__ldg(q.q_hji_t + idx) < q.n_hji) ? __ldg(q.wtab + __ldg(q.q_w + idx)) : 0.0;
                s.lo[p] = Pv;      // P_jj
                s.hi[p] = 0.0;
                s.yq[p] = qv;      // q_j
                s.flag[p] = 0;
            }
            s.sc[p] = 1.0;         // D_j | E_i
            s.xz[p] = warm ? a.ws_xz[(size_t)v * q.Nk + p] : 0.0;
        }
        double rho = warm ? a.rho[v] : st.rho;
        double c = 1.0;
        __syncthreads();
        PHASE(0);

        // ---- 2. modified Ruiz equilibration (scale_data of OSQP) -------------------------------------------------------
        for (int it = 0; it < st.scaling; it++) {
            // inf-norms of the KKT columns (variables: P_jj and column j of A; constraints: row i of A), four gathers in flight per thread;
            // the square roots and reciprocals (~220 cycles each, serial) of a thread's positions are formed together after the scans
            for (int p0 = tid; p0 < q.Nk; p0 += 3 * ADMM_THREADS) {
                double nrm[3];
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int p = p0 + u * ADMM_THREADS;
                    if (p >= q.Nk) { nrm[u] = 1.0; continue; }
                    double n0 = s.flag[p] ? 0.0 : fabs(s.lo[p]), n1 = 0.0;
                    int e = s.kptr[p];
                    const int e1 = s.kptr[p + 1];
#pragma unroll 1
                    for (; e + 4 <= e1; e += 4) {
                        const int j0 = s.ke[e], j1 = s.ke[e + 1], j2 = s.ke[e + 2], j3 = s.ke[e + 3];
                        const double a0 = s.Aval[j0], a1 = s.Aval[j1], a2 = s.Aval[j2], a3 = s.Aval[j3];
                        n0 = fmax(n0, fmax(fabs(a0), fabs(a1))); n1 = fmax(n1, fmax(fabs(a2), fabs(a3)));
                    }
                    if (e < e1) {
                        const int last = e1 - 1;
                        const int j0 = s.ke[e], j1 = s.ke[min(e + 1, last)], j2 = s.ke[min(e + 2, last)];
                        const double a0 = s.Aval[j0], a1 = s.Aval[j1], a2 = s.Aval[j2];
                        n0 = fmax(n0, fmax(fabs(a0), fabs(a1))); n1 = fmax(n1, fabs(a2));
                    }
                    nrm[u] = limit_scaling(fmax(n0, n1));
                }
                const double r0 = 1.0 / sqrt(nrm[0]), r1 = 1.0 / sqrt(nrm[1]), r2 = 1.0 / sqrt(nrm[2]);
                s.sol[p0] = r0;
                if (p0 + ADMM_THREADS < q.Nk) s.sol[p0 + ADMM_THREADS] = r1;
                if (p0 + 2 * ADMM_THREADS < q.Nk) s.sol[p0 + 2 * ADMM_THREADS] = r2;
            }
            __syncthreads();
            {
                int e = tid;
                for (; e + ADMM_THREADS < q.nnzA; e += 2 * ADMM_THREADS) {       // two independent entries per trip
                    const uint32_t rc0 = s.arc[e], rc1 = s.arc[e + ADMM_THREADS];
                    const double f0 = s.sol[rc0 & 0xffff] * s.sol[rc0 >> 16], f1 = s.sol[rc1 & 0xffff] * s.sol[rc1 >> 16];
                    const double a0 = s.Aval[e], a1 = s.Aval[e + ADMM_THREADS];
                    s.Aval[e] = a0 * f0; s.Aval[e + ADMM_THREADS] = a1 * f1;
                }
                if (e < q.nnzA) { const uint32_t rc = s.arc[e]; s.Aval[e] *= s.sol[rc & 0xffff] * s.sol[rc >> 16]; }
            }
            double sumP = 0.0, maxq = 0.0;   // sum |P_jj|, max |q_j|
            for (int p = tid; p < q.Nk; p += ADMM_THREADS) {
                const double d = s.sol[p];
                s.sc[p] *= d;
                if (!s.flag[p]) {
                    s.lo[p] *= d * d;
                    s.yq[p] *= d;
                    sumP += fabs(s.lo[p]);
                    maxq = fmax(maxq, fabs(s.yq[p]));
                }
            }
            block_reduce_sum_max(sumP, maxq, s.red);
            double c_temp = sumP / q.n;
            const double inf_q = limit_scaling(maxq);
            c_temp = limit_scaling(fmax(c_temp, inf_q));
            c_temp = 1.0 / c_temp;
            for (int p = tid; p < q.Nk; p += ADMM_THREADS)
                if (!s.flag[p]) { s.lo[p] *= c_temp; s.yq[p] *= c_temp; }
            c *= c_temp;
            // no barrier here: every thread owns the same positions p in all per-position loops, and the A values (scaled by other threads
            // above) are fenced from the next pass's scans by the two barriers of the block reduction
        }
        __syncthreads();
        const double cinv = 1.0 / c;
        // bounds scaled by E; constraint classes (set_rho_vec of OSQP)
        for (int p = tid; p < q.Nk; p += ADMM_THREADS)
            if (s.flag[p]) {
                const double E = s.sc[p];
                const double l = s.lo[p] * E, u = s.hi[p] * E;
                s.lo[p] = l; s.hi[p] = u;
                s.flag[p] = (l < -OSQP_INFTY * 1e-4 && u > OSQP_INFTY * 1e-4) ? 3 : ((u - l < 1e-4) ? 2 : 1);
            }
        __syncthreads();

        PHASE(1);
        // ---- 3. factor ----------------------------------------------------------------------------------------------------
        factor<PROF>(q, s, st.sigma, rho, PROF ? s_cyc + 16 : nullptr);
        PHASE(2);

        // ---- 4. ADMM iterations ---------------------------------------------------------------------------------------------
        int iter = 0, status = PGN_QP_UNSOLVED, n_rho_upd = 0;
        double pri_res = 0.0, dua_res = 0.0;
        const double alpha = st.alpha;
        RhoInv rinv = make_rho_inv(rho);
        // right-hand side of the KKT system; the first range reads it from the scratch vector
#define ADMM_RHS(p, f)                                                                                              \
        do {                                                                                                        \
            const double b__ = (f) ? s.xz[p] - rinv_of((f), rinv) * s.yq[p] : st.sigma * s.xz[p] - s.yq[p];         \
            if ((p) < q.rhs_tmp_end) s.dxy[p] = b__; else s.sol[p] = b__;                                           \
        } while (0)
        for (int p = tid; p < q.Nk; p += ADMM_THREADS) { const uint8_t f = s.flag[p]; ADMM_RHS(p, f); }
        __syncthreads();
        // iterations until the next residual check / rho estimate (countdowns: a run-time modulo is ~25 instructions on every warp)
        int chk_left = st.check_termination > 0 ? st.check_termination : 0x7fffffff;
        int adp_left = (st.adaptive_rho && st.adaptive_rho_interval > 0) ? st.adaptive_rho_interval : 0x7fffffff;
        for (iter = 1; iter <= st.max_iter; iter++) {
            const bool check = --chk_left == 0, adapt = --adp_left == 0;
            if (check) chk_left = st.check_termination;
            if (adapt) adp_left = st.adaptive_rho_interval;
            const bool need_delta = check || adapt;
            kkt_solve<PROF>(a, s, PROF ? s_cyc + 16 : nullptr);
            PHASE(3);
            if (!need_delta) {
                // x, z, y updates fused with the next right-hand side.  Branch-free (both the constraint and the variable form are
                // evaluated, selects pick one) with three positions per thread in flight: the pass is a latency chain of each warp's
                // own instructions, so independent positions are interleaved instead of run one after the other.
                for (int p0 = tid; p0 < q.Nk; p0 += 3 * ADMM_THREADS) {
                    int pp[3]; bool ok[3]; uint8_t f[3];
                    double xz[3], y[3], so[3], lo[3], hi[3];
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        const int p = p0 + u * ADMM_THREADS;
                        ok[u] = p < q.Nk; pp[u] = ok[u] ? p : p0;
                        f[u] = s.flag[pp[u]]; xz[u] = s.xz[pp[u]]; y[u] = s.yq[pp[u]]; so[u] = s.sol[pp[u]]; lo[u] = s.lo[pp[u]]; hi[u] = s.hi[pp[u]];
                    }
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        const bool con = f[u] != 0;
                        const double r = rho_of(f[u], rho), ri = rinv_of(f[u], rinv);
                        const double zt = xz[u] + ri * (so[u] - y[u]);
                        const double zr = alpha * zt + (1.0 - alpha) * xz[u];
                        const double zn = fmin(fmax(zr + ri * y[u], lo[u]), hi[u]);
                        const double dy = r * (zr - zn);
                        const double xn = alpha * so[u] + (1.0 - alpha) * xz[u];
                        const double nx = con ? zn : xn, ny = con ? y[u] + dy : y[u];
                        const double b = con ? nx - ri * ny : st.sigma * nx - ny;
                        if (ok[u]) {
                            s.xz[pp[u]] = nx;
                            if (con) s.yq[pp[u]] = ny;
                            if (pp[u] < q.rhs_tmp_end) s.dxy[pp[u]] = b; else s.sol[pp[u]] = b;
                        }
                    }
                }
                __syncthreads();
                PHASE(4);
                continue;
            }
            for (int p = tid; p < q.Nk; p += ADMM_THREADS) {
                const uint8_t f = s.flag[p];
                if (f) {
                    const double r = rho_of(f, rho), ri = rinv_of(f, rinv);
                    const double zp = s.xz[p], y = s.yq[p];
                    const double zt = zp + ri * (s.sol[p] - y);
                    const double zr = alpha * zt + (1.0 - alpha) * zp;
                    const double zn = fmin(fmax(zr + ri * y, s.lo[p]), s.hi[p]);
                    const double dy = r * (zr - zn);
                    s.xz[p] = zn;
                    s.yq[p] = y + dy;
                    s.dxy[p] = dy;
                } else {
                    const double xp = s.xz[p];
                    const double xn = alpha * s.sol[p] + (1.0 - alpha) * xp;
                    s.xz[p] = xn;
                    s.dxy[p] = xn - xp;
                }
            }
            __syncthreads();
            PHASE(4);
            Resid R = residuals(q, s, cinv);
            pri_res = R.pri_res; dua_res = R.dua_res;
            if (check) {
                const double eps_prim = st.eps_abs + st.eps_rel * R.eps_pri_n;
                const double eps_dual = st.eps_abs + st.eps_rel * R.eps_dua_n;
                const bool prim_ok = R.pri_res < eps_prim, dual_ok = R.dua_res < eps_dual;
                bool pinf = false, dinf = false;
                if (!prim_ok) pinf = primal_infeasible(q, s, st.eps_prim_inf);
                if (!dual_ok) dinf = dual_infeasible(q, s, st.eps_dual_inf, c);
                if (prim_ok && dual_ok) { status = PGN_QP_SOLVED; break; }
                if (pinf) { status = PGN_QP_PRIMAL_INFEASIBLE; break; }
                if (dinf) { status = PGN_QP_DUAL_INFEASIBLE; break; }
            }
            if (adapt) {
                // compute_rho_estimate / adapt_rho of OSQP (scaled norms)
                const double pr = R.s_pri / (R.s_pn + 1e-10), du = R.s_dua / (R.s_dn + 1e-10);
                double rho_new = rho * sqrt(pr / (du + 1e-10));
                rho_new = fmin(fmax(rho_new, 1e-6), 1e6);
                if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
                    rho = rho_new;
                    rinv = make_rho_inv(rho);
                    n_rho_upd++;
                    __syncthreads();
                    PHASE(5);
                    factor<PROF>(q, s, st.sigma, rho, PROF ? s_cyc + 16 : nullptr);
                    PHASE(2);
                }
            }
            __syncthreads();       // the checks used sol / the scratch vector as work space
            for (int p = tid; p < q.Nk; p += ADMM_THREADS) { const uint8_t f = s.flag[p]; ADMM_RHS(p, f); }
            __syncthreads();
            PHASE(5);
        }
        if (iter > st.max_iter) {
            iter = st.max_iter;
            // approximate termination test (10x tolerances) before declaring max_iter_reached
            Resid R = residuals(q, s, cinv);
            pri_res = R.pri_res; dua_res = R.dua_res;
            const double eps_prim = 10 * st.eps_abs + 10 * st.eps_rel * R.eps_pri_n, eps_dual = 10 * st.eps_abs + 10 * st.eps_rel * R.eps_dua_n;
            status = (R.pri_res < eps_prim && R.dua_res < eps_dual) ? PGN_QP_SOLVED_INACCURATE : PGN_QP_MAX_ITER_REACHED;
        }

        PHASE(5);
        // ---- 5. store ------------------------------------------------------------------------------------------------------------
        const bool infeas = (status == PGN_QP_PRIMAL_INFEASIBLE || status == PGN_QP_DUAL_INFEASIBLE);
        for (int p = tid; p < q.Nk; p += ADMM_THREADS) {
            const int idx = __ldg(q.pos2idx + p);
            if (s.flag[p]) {
                a.sol_y[(size_t)v * q.m + idx] = infeas ? NAN : s.sc[p] * s.yq[p] * cinv;
                a.ws_y[(size_t)v * q.Nk + p] = infeas ? 0.0 : s.yq[p];
            } else {
                a.sol_x[(size_t)v * q.n + idx] = infeas ? NAN : s.sc[p] * s.xz[p];
                a.ws_y[(size_t)v * q.Nk + p] = 0.0;
            }
            a.ws_xz[(size_t)v * q.Nk + p] = infeas ? 0.0 : s.xz[p];
        }
        if (tid == 0) {
            a.rho[v] = rho;
            a.iters[v] = iter; a.status[v] = status; a.rho_updates[v] = n_rho_upd;
            a.pri_res[v] = pri_res; a.dua_res[v] = dua_res;
            a.solved[v] = 1;
            a.cold[v] = 0;
        }
        __syncthreads();
        PHASE(6);
    }
    if (PROF && a.cycles) {
        __syncthreads();
        for (int i = threadIdx.x; i < ADMM_NCYC; i += ADMM_THREADS)
            if (s_cyc[i]) atomicAdd(a.cycles + i, (unsigned long long)s_cyc[i]);
    }
}

// Ticket order of the persistent CTAs: vehicles whose previous solve took the most iterations go first (longest-processing-time-first),
// so that a slow QP starts at the beginning of the launch instead of becoming its tail.  Counting sort on iters / 25 in one CTA.
__global__ void __launch_bounds__(1024) k_admm_order(const int32_t* __restrict__ iters, int32_t* __restrict__ order, int B) {
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
    for (int v = threadIdx.x; v < B; v += blockDim.x) order[atomicAdd(&hist[255 - min(iters[v] / 25, 255)], 1)] = v;
}

int admm_configure(pgn_handle* h) {
    h->admm_smem_bytes = (int)admm_smem_bytes(h->tab);
    h->admm_threads = ADMM_THREADS;
    cudaError_t e = cudaFuncSetAttribute(k_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->admm_smem_bytes);
    return (int)e;
}

void launch_admm(pgn_handle* h) {
    AdmmArgs a;
    a.q = h->qd; a.st = h->st; a.B = h->B;
    a.rec = h->d_rec; a.ws_xz = h->d_ws_xz; a.ws_y = h->d_ws_y; a.rho = h->d_rho;
    a.sol_x = h->d_sol_x; a.sol_y = h->d_sol_y;
    a.iters = h->d_iters; a.status = h->d_status; a.rho_updates = h->d_rho_updates; a.pri_res = h->d_pri_res; a.dua_res = h->d_dua_res;
    a.solved = h->d_solved; a.counter = h->d_counter;
    a.cycles = h->profiling >= 2 ? h->d_cycles : nullptr;      // 1: stage timers only, 2: + in-kernel cycle counters
    cudaMemsetAsync(h->d_counter, 0, sizeof(int), h->stream);
    k_admm_order<<<1, 1024, 0, h->stream>>>(h->d_iters, h->d_order, h->B);
    h->launches++;
    a.order = h->d_order;
    a.skip = (h->guard_pause > 0.0 || h->in_callback) ? h->d_skip : nullptr; a.cold = h->d_cold;
    for (size_t i = 0; i < h->tab.sol_ph_ptr.size() && i <= ADMM_MAX_PHASES; i++) a.ph_ptr[i] = h->tab.sol_ph_ptr[i];
    int ctas_per_sm = 1;
    if (h->admm_smem_bytes * 2 + 2048 <= 227 * 1024) ctas_per_sm = 2;
    if (h->admm_smem_bytes * 3 + 3072 <= 227 * 1024) ctas_per_sm = 3;
    int grid = h->num_sms * ctas_per_sm;
    if (grid > h->B) grid = h->B;
    if (a.cycles) k_admm<true><<<grid, ADMM_THREADS, h->admm_smem_bytes, h->stream>>>(a);
    else k_admm<false><<<grid, ADMM_THREADS, h->admm_smem_bytes, h->stream>>>(a);
    h->launches++;
}

}  // namespace pgn
