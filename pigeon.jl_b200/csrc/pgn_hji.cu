// pgn_hji.cu — HJI value/gradient lookup and the reachability (safety) constraint of the coupled controller.
//   HJIRelativeState          reference src/HJI_computation.jl:20-24
//   cache[x] (7-D multilinear) src/HJI_computation.jl:66-72  (Interpolations.jl Gridded(Linear()), Float32 table, Float64 weights)
//   optimal_disturbance       src/HJI_computation.jl:90-131
//   compute_reachability_constraint  src/HJI_computation.jl:160-170, call site src/coupled_lat_long.jl:341-346
//
// HBM layout: the reference keeps V (Float32) and gradV (SVector{7,Float32}) as two tables; here every grid node is ONE
// 32-byte record {gradV[0..6], V} (the 7->8 padding slot carries V), dimension 1 fastest.  A query touches the 2^7 corners of
// its cell = 128 records = 128 sectors of 32 B = exactly the algorithmic 4096 B, fetched as 256 x LDG.128 with all of a
// thread's loads for one dim-1 pair issued back to back (64 B contiguous).
#include <cuda.h>

#include "pgn_internal.h"

namespace pgn {

struct HjiCell { int idx[7]; double w[7]; bool inside; };

__device__ __forceinline__ HjiCell hji_locate(const HjiView& H, const double* x) {
    HjiCell c;
    c.inside = true;
#pragma unroll
    for (int d = 0; d < 7; d++) {
        const float* k = H.knots + H.kofs[d];
        const int nk = H.dims[d];
        const double lo = (double)__ldg(k), hi = (double)__ldg(k + nk - 1);
        if (!(lo <= x[d] && x[d] <= hi)) c.inside = false;
        int a = 0, b = nk;
        while (a < b) { int mid = (a + b) >> 1; if ((double)__ldg(k + mid) <= x[d]) a = mid + 1; else b = mid; }
        int i = min(max(a, 1), nk - 1);
        c.idx[d] = i - 1;
        const double k0 = (double)__ldg(k + i - 1), k1 = (double)__ldg(k + i);
        c.w[d] = (x[d] - k0) / (k1 - k0);
    }
    return c;
}

// multilinear interpolation of the 8-float node records; out[0..6] = gradV, out[7] = V
__device__ __forceinline__ void hji_interp(const HjiView& H, const HjiCell& c, double* out) {
    long long base = 0;
#pragma unroll
    for (int d = 0; d < 7; d++) base += (long long)c.idx[d] * H.stride[d];
    const float4* tab = reinterpret_cast<const float4*>(H.gV);
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.0;
    // dims 2..7 enumerate 64 corner pairs; the pair along dim 1 is 64 contiguous bytes
    for (int cnr = 0; cnr < 64; cnr++) {
        long long off = base;
        double wgt = 1.0;
#pragma unroll
        for (int d = 1; d < 7; d++) {
            const int bit = (cnr >> (d - 1)) & 1;
            off += bit ? H.stride[d] : 0;
            wgt *= bit ? c.w[d] : (1.0 - c.w[d]);
        }
        const float4 a0 = __ldg(tab + 2 * off), a1 = __ldg(tab + 2 * off + 1), b0 = __ldg(tab + 2 * off + 2), b1 = __ldg(tab + 2 * off + 3);
        const double w0 = wgt * (1.0 - c.w[0]), w1 = wgt * c.w[0];
        acc[0] += w0 * a0.x + w1 * b0.x; acc[1] += w0 * a0.y + w1 * b0.y; acc[2] += w0 * a0.z + w1 * b0.z; acc[3] += w0 * a0.w + w1 * b0.w;
        acc[4] += w0 * a1.x + w1 * b1.x; acc[5] += w0 * a1.y + w1 * b1.y; acc[6] += w0 * a1.z + w1 * b1.z; acc[7] += w0 * a1.w + w1 * b1.w;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = acc[k];
}

// stand-alone policy evaluation: x [M][7] relative states, gV [M][7], out [M][2] = (delta, Fx)
__global__ void __launch_bounds__(128) k_hji_optimal_control(VehParams P, int M, const double* __restrict__ x, const double* __restrict__ gV, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double g[7];
#pragma unroll
    for (int k = 0; k < 7; k++) g[k] = gV[(size_t)i * 7 + k];
    double d, Fx;
    hji_optimal_control(P, x[(size_t)i * 7 + 3], x[(size_t)i * 7 + 4], x[(size_t)i * 7 + 6], g, d, Fx);
    out[(size_t)i * 2] = d; out[(size_t)i * 2 + 1] = Fx;
}

// stand-alone lookup: x [7][M] field-major, V [M], gradV [7][M]
__global__ void __launch_bounds__(128) k_hji_lookup(HjiView H, int M, const double* __restrict__ x, double* __restrict__ V, double* __restrict__ gV) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double xq[7];
#pragma unroll
    for (int d = 0; d < 7; d++) xq[d] = x[(size_t)d * M + i];
    HjiCell c = hji_locate(H, xq);
    double out[8];
    if (c.inside) hji_interp(H, c, out);
    else {
#pragma unroll
        for (int k = 0; k < 7; k++) out[k] = 0.0;
        out[7] = INFINITY;
    }
    V[i] = out[7];
#pragma unroll
    for (int k = 0; k < 7; k++) gV[(size_t)k * M + i] = out[k];
}

// reachability constraint M u + b >= -sigma for every vehicle; writes (M1*un1, M2*un2, b) into the record
__global__ void __launch_bounds__(128) k_hji_constraint(HjiView H, int B, int v0, int nv, VehParams P, double eps, double un0, double un1, const double* __restrict__ state,
                                                        const double* __restrict__ control, const double* __restrict__ other, double* __restrict__ rec,
                                                        int rec_len, int o_hji, double* __restrict__ hji_val /*[8][B]: gradV[0..6], V*/, const uint8_t* __restrict__ hold) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    if (hold && hold[v]) return;          // deferred solve in progress: the record keeps the constraint of the step being solved
    const double E = state[0 * B + v], N = state[1 * B + v], psi = state[2 * B + v], Ux = state[3 * B + v], Uy = state[4 * B + v], r = state[5 * B + v];
    const double oE = other[0 * B + v], oN = other[1 * B + v], opsi = other[2 * B + v], oV = other[3 * B + v];
    // HJIRelativeState: `cψ, sψ = sincos(-ψ)` binds cψ <- sin(-ψ), sψ <- cos(-ψ) (HJI_computation.jl:21-22)
    double sn, cs;
    sincos(-psi, &sn, &cs);
    const double cpsi = sn, spsi = cs;
    double x7[7];
    x7[0] = cpsi * (oE - E) + spsi * (oN - N);
    x7[1] = -spsi * (oE - E) + cpsi * (oN - N);
    x7[2] = adiff(opsi, psi);
    x7[3] = Ux; x7[4] = Uy; x7[5] = oV; x7[6] = r;
    double M0 = 0, M1 = 0, b = 1.0;
    double g[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, INFINITY};      // cache[x] outside the grid: (V = Inf, gradV = 0)
    bool active = false;
    if (H.valid) {
        HjiCell c = hji_locate(H, x7);
        if (c.inside) { hji_interp(H, c, g); active = !(g[7] > eps); }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) hji_val[(size_t)k * B + v] = g[k];    // V, gradV of this step: published by the callback, read by the policy override
    if (active) {
        // optimal_disturbance (dMode = :min). Deviation: other-car speed <= 0 gives (0,0) instead of NaN
        double uH0 = 0, uH1 = 0;
        const double Vo = x7[5];
        if (Vo > 0) {
            const double Ax_max = P.Fx_max / P.m, Pmx_max = P.Px_max / P.m, maxA = 0.9 * P.mu * P.G;
            const double lam_Ax = g[5], lam_Ay = g[2] / Vo;
            const double lam_norm = hypot(lam_Ax, lam_Ay);
            if (!(lam_norm < 1e-3)) {
                const double desAx = -lam_Ax * maxA / lam_norm, desAy = -lam_Ay * maxA / lam_norm;
                double maxAx = fmin(Ax_max, Pmx_max / Vo), maxAy = P.kappa_max * Vo * Vo;
                if (desAx > maxAx) {
                    if (fabs(desAy) < maxAy) maxAy = fmin(maxAy, sqrt(maxA * maxA - maxAx * maxAx));
                    uH0 = copysign(maxAy, desAy) / Vo; uH1 = maxAx;
                } else if (fabs(desAy) > maxAy) {
                    if (desAx > 0) { maxAx = fmin(sqrt(maxA * maxA - maxAy * maxAy), maxAx); uH0 = copysign(maxAy, desAy) / Vo; uH1 = maxAx; }
                    else { uH0 = copysign(maxAy, desAy) / Vo; uH1 = -sqrt(maxA * maxA - maxAy * maxAy); }
                } else { uH0 = desAy / Vo; uH1 = maxAx; }
            }
        }
        // H(uR) = gradV . relative_dynamics(x, uR, uH); M = dH/duR by forward AD over (delta, Fx); b = H - M.uR
        typedef Dual<2> D;
        const double d0 = control[0 * B + v], Fx0 = control[1 * B + v] + control[2 * B + v];
        D q[6] = {D(x7[0]), D(x7[1]), D(x7[2]), D(x7[3]), D(x7[4]), D(x7[6])}, out[6];
        D du(d0), dF(Fx0);
        du.d[0] = 1.0; dF.d[1] = 1.0;
        vehicle_model<MODEL_BICYCLE, D>(P, q, du, dF, D(0.0), D(0.0), out);
        double s3, c3;
        sincos(x7[2], &s3, &c3);
        const double f0 = x7[5] * c3 - x7[3] + x7[1] * x7[6];
        const double f1 = x7[5] * s3 - x7[4] - x7[0] * x7[6];
        const double f2 = uH0 - x7[6];
        const double Hv = g[0] * f0 + g[1] * f1 + g[2] * f2 + g[3] * out[3].v + g[4] * out[4].v + g[5] * uH1 + g[6] * out[5].v;
        M0 = g[3] * out[3].d[0] + g[4] * out[4].d[0] + g[6] * out[5].d[0];
        M1 = g[3] * out[3].d[1] + g[4] * out[4].d[1] + g[6] * out[5].d[1];
        b = Hv - (M0 * d0 + M1 * Fx0);
    }
    double* glob = rec + (size_t)v * rec_len + o_hji;
    glob[0] = M0 * un0; glob[1] = M1 * un1; glob[2] = b;
}

void launch_hji_constraint(pgn_handle* h) {
    const int B = h->B;
    k_hji_constraint<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(h->hji, B, h->v0, h->nv, h->veh, h->cfg.hji_eps, h->un[0], h->un[1], h->d_state, h->d_control, h->d_other,
                                                             h->d_rec, h->tab.rec.rec_len, h->tab.rec.o_hji, h->d_hji_val, h->hold_on ? h->d_hold : nullptr);
    h->launches++;
}
void launch_hji_optimal_control(pgn_handle* h, int M, const double* d_x, const double* d_gV, double* d_out) {
    k_hji_optimal_control<<<(M + 127) / 128, 128, 0, h->stream>>>(h->veh, M, d_x, d_gV, d_out);
    h->launches++;
}
// ---- large query sets: visit the queries in CELL ORDER ------------------------------------------------------------------------------------
// A random query moves 6.6 KB through DRAM for its 4 KB of corners: the 64 corner pairs are 64-byte segments and every random sub-128-byte
// access costs this memory system ~100 bytes (tools/ubench/hji_fetch.cu: 64-byte aligned segments read 1.64x, unaligned pairs 2.0x, single
// 32-byte records 3.0x their bytes), so no layout of the same records cures it.  What does: queries of the same and of neighbouring cells
// share their corners, and visited in cell order they find them in L2 — the same benchmark reads the TABLE SIZE instead of 34 GB and runs
// 9x faster.  From 2^19 queries up (measured break-even: 2^17 loses 20 %, 2^20 gains 47 %, 2^24 gains 94 %) the lookup therefore counting-sorts the query indices by cell (dimension 1 fastest, as the table): locate +
// histogram, exclusive scan over the cells, scatter, then the gather walks the permutation.  Results are bit-identical to the direct kernel.
__global__ void __launch_bounds__(256) k_hji_keys(HjiView H, int M, long long ncell, const double* __restrict__ x, uint32_t* __restrict__ keys, uint32_t* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double xq[7];
#pragma unroll
    for (int d = 0; d < 7; d++) xq[d] = x[(size_t)d * M + i];
    const HjiCell c = hji_locate(H, xq);
    long long key = 0, cs = 1;
#pragma unroll
    for (int d = 0; d < 7; d++) { key += (long long)c.idx[d] * cs; cs *= H.dims[d] - 1; }
    const uint32_t k = c.inside ? (uint32_t)key : (uint32_t)ncell;      // queries outside the grid: one bucket behind the last cell
    keys[i] = k;
    atomicAdd(cnt + k, 1u);
}
// exclusive scan of n counters in three passes (2048 per block)
__global__ void __launch_bounds__(256) k_scan_blocks(uint32_t* __restrict__ a, long long n, uint32_t* __restrict__ totals) {
    __shared__ uint32_t wsum[8];
    const long long base = (long long)blockIdx.x * 2048 + threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = base + k < n ? a[base + k] : 0u; s += v[k]; }
    uint32_t incl = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < w; k++) woff += wsum[k];
    uint32_t run = woff + incl - s;
#pragma unroll
    for (int k = 0; k < 8; k++) { if (base + k < n) a[base + k] = run; run += v[k]; }
    if (threadIdx.x == 255) totals[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t* __restrict__ totals, int nb) {
    __shared__ uint32_t carry, wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? totals[i] : 0u;
        uint32_t incl = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; k++) woff += wsum[k];
        if (i < nb) totals[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += woff + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t* __restrict__ a, long long n, const uint32_t* __restrict__ totals) {
    const long long base = (long long)blockIdx.x * 2048 + threadIdx.x * 8;
    const uint32_t off = totals[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; k++) if (base + k < n) a[base + k] += off;
}
__global__ void __launch_bounds__(256) k_hji_scatter(int M, const uint32_t* __restrict__ keys, uint32_t* __restrict__ offs, uint32_t* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    perm[atomicAdd(offs + keys[i], 1u)] = (uint32_t)i;
}
__global__ void __launch_bounds__(128) k_hji_lookup_perm(HjiView H, int M, const uint32_t* __restrict__ perm, const double* __restrict__ x, double* __restrict__ V, double* __restrict__ gV) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int i = (int)perm[j];
    double xq[7];
#pragma unroll
    for (int d = 0; d < 7; d++) xq[d] = x[(size_t)d * M + i];
    HjiCell c = hji_locate(H, xq);
    double out[8];
    if (c.inside) hji_interp(H, c, out);
    else {
#pragma unroll
        for (int k = 0; k < 7; k++) out[k] = 0.0;
        out[7] = INFINITY;
    }
    V[i] = out[7];
#pragma unroll
    for (int k = 0; k < 7; k++) gV[(size_t)k * M + i] = out[k];
}

// ---- very large query sets: one TMA-staged tile of corners per block of cells -----------------------------------------------------------------
// In cell order the gather is bound by L2 -> SM traffic: every query still pulls its own 4 KB of corners.  The queries of a BLOCK of cells —
// all cells along dimension 1, TMA_G consecutive cells along dimension 2, one cell in dimensions 3..7 — are one contiguous range of the
// sorted order and share one set of corner records: (TMA_G + 1) x 2^5 dimension-1 rows of the table.  The table is described to the TMA unit
// as a 5-D tensor (8 n1 floats | n2 | n3 | n4 | n5 n6 n7 — the three slow dimensions nest contiguously), one CTA stages the block's tile with
// four box loads of (8 n1) x (TMA_G + 1) x 2 x 2 x 2 floats (cp.async.bulk.tensor.5d, completion on an mbarrier; rows past the grid are
// zero-filled and never addressed), and its threads interpolate their queries out of shared memory with the arithmetic of hji_interp in the
// same order: bit-identical results.  For the 13 x 13 x 9^5 grid: 98 304 blocks, 66.6 KB per tile, ~170 queries per block at 2^24 queries.
// Measured slower than the plain cell-ordered gather (see launch_hji_lookup): selected only on request.
#define TMA_G 4
__global__ void __launch_bounds__(128) k_hji_lookup_tma(const __grid_constant__ CUtensorMap tmap, HjiView H, int M, const uint32_t* __restrict__ offs, const uint32_t* __restrict__ perm,
                                                        const double* __restrict__ x, double* __restrict__ V, double* __restrict__ gV) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) unsigned long long mbar;
    // block -> (group along dim 2, cell coordinates of dims 3..7)
    const int nc1 = H.dims[0] - 1, nc2 = H.dims[1] - 1;
    const int ng = (nc2 + TMA_G - 1) / TMA_G;
    int b = blockIdx.x;
    const int g2 = b % ng; b /= ng;
    int c[7];
    c[1] = g2 * TMA_G;
#pragma unroll
    for (int d = 2; d < 7; d++) { c[d] = b % (H.dims[d] - 1); b /= (H.dims[d] - 1); }
    long long cell_lo = 0, cs = 1;
#pragma unroll
    for (int d = 1; d < 7; d++) { cs *= H.dims[d - 1] - 1; cell_lo += (long long)c[d] * cs; }        // c1 = 0
    const int rows = min(TMA_G, nc2 - c[1]);
    const long long cell_hi = cell_lo + (long long)rows * nc1;
    const uint32_t q0 = cell_lo ? offs[cell_lo - 1] : 0u, q1 = offs[cell_hi - 1];       // after the scatter offs[c] is the END of cell c's queries
    if (q0 == q1) return;                                                                 // no query in this block (uniform)
    const int row_floats = 8 * H.dims[0], box_floats = row_floats * (TMA_G + 1) * 8;
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(tma_smem), mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(4 * box_floats * 4)) : "memory");
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            const int c4 = c[4] + H.dims[4] * ((c[5] + (bb & 1)) + H.dims[5] * (c[6] + (bb >> 1)));
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(sm + (uint32_t)(bb * box_floats * 4)),
                         "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(0), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c4), "r"(mb)
                         : "memory");
        }
    }
    __syncthreads();                       // the barrier is initialised for everyone
    {
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"(0) : "memory");
    }
    const float4* tile = reinterpret_cast<const float4*>(tma_smem);
    for (uint32_t j = q0 + threadIdx.x; j < q1; j += blockDim.x) {
        const int i = (int)perm[j];
        double xq[7];
#pragma unroll
        for (int d = 0; d < 7; d++) xq[d] = x[(size_t)d * M + i];
        const HjiCell cl = hji_locate(H, xq);
        double acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.0;
        const int r2 = cl.idx[1] - c[1];                              // row of the query's cell inside the tile
        for (int cnr = 0; cnr < 64; cnr++) {                           // the loop of hji_interp, corners read from the tile
            double wgt = 1.0;
#pragma unroll
            for (int d = 1; d < 7; d++) {
                const int bit = (cnr >> (d - 1)) & 1;
                wgt *= bit ? cl.w[d] : (1.0 - cl.w[d]);
            }
            const int b2 = cnr & 1, b3 = (cnr >> 1) & 1, b4 = (cnr >> 2) & 1, b5 = (cnr >> 3) & 1, bb = (cnr >> 4) & 3;
            const int frow = ((((bb * 2 + b5) * 2 + b4) * 2 + b3) * (TMA_G + 1) + r2 + b2);        // tile row: [box][i5][i4][i3][i2]
            const float4* rec = tile + (size_t)frow * (row_floats / 4) + 2 * cl.idx[0];
            const float4 a0 = rec[0], a1 = rec[1], b0 = rec[2], b1 = rec[3];
            const double w0 = wgt * (1.0 - cl.w[0]), w1 = wgt * cl.w[0];
            acc[0] += w0 * a0.x + w1 * b0.x; acc[1] += w0 * a0.y + w1 * b0.y; acc[2] += w0 * a0.z + w1 * b0.z; acc[3] += w0 * a0.w + w1 * b0.w;
            acc[4] += w0 * a1.x + w1 * b1.x; acc[5] += w0 * a1.y + w1 * b1.y; acc[6] += w0 * a1.z + w1 * b1.z; acc[7] += w0 * a1.w + w1 * b1.w;
        }
        V[i] = acc[7];
#pragma unroll
        for (int k = 0; k < 7; k++) gV[(size_t)k * M + i] = acc[k];
    }
}
// the queries outside the grid (the bucket behind the last cell): cache[x] = (Inf, 0)
__global__ void __launch_bounds__(256) k_hji_outside(int M, const uint32_t* __restrict__ offs, long long ncell, const uint32_t* __restrict__ perm, double* __restrict__ V, double* __restrict__ gV) {
    const uint32_t j = offs[ncell - 1] + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (uint32_t)M) return;
    const int i = (int)perm[j];
    V[i] = INFINITY;
#pragma unroll
    for (int k = 0; k < 7; k++) gV[(size_t)k * M + i] = 0.0;
}

// the table as a 5-D tensor for the TMA unit; returns false when the grid does not fit the scheme (then the lookups never take the TMA path)
bool hji_make_tensor_map(pgn_handle* h) {
    h->hji_tma_valid = 0;
    const HjiView& H = h->hji;
    if (!H.valid || 8 * H.dims[0] > 256 || H.dims[1] < 2) return false;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
    const cuuint64_t n1 = H.dims[0], n2 = H.dims[1], n3 = H.dims[2], n4 = H.dims[3], n567 = (cuuint64_t)H.dims[4] * H.dims[5] * H.dims[6];
    const cuuint64_t gdim[5] = {8 * n1, n2, n3, n4, n567};
    const cuuint64_t gstr[4] = {32 * n1, 32 * n1 * n2, 32 * n1 * n2 * n3, 32 * n1 * n2 * n3 * n4};      // bytes, dimensions 1..4
    const cuuint32_t box[5] = {(cuuint32_t)(8 * n1), TMA_G + 1, 2, 2, 2};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    static_assert(sizeof(CUtensorMap) <= sizeof(((pgn_handle*)0)->hji_tmap), "tensor map storage too small");
    CUresult r = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap*>(h->hji_tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)H.gV, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    const int tile_bytes = (int)(4 * 8 * n1 * (TMA_G + 1) * 8 * 4);
    if (tile_bytes > 200 * 1024) return false;
    if (cudaFuncSetAttribute(k_hji_lookup_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    h->hji_tma_tile_bytes = tile_bytes;
    long long nb = ((long long)(n2 - 1) + TMA_G - 1) / TMA_G;
    for (int d = 2; d < 7; d++) nb *= H.dims[d] - 1;
    if (nb >= (1ll << 31)) return false;
    h->hji_tma_blocks = nb;
    h->hji_tma_valid = 1;
    return true;
}

void launch_hji_lookup(pgn_handle* h, int M, const double* d_x, double* d_V, double* d_gV) {
    long long ncell = 1;
    for (int d = 0; d < 7; d++) ncell *= h->hji.dims[d] - 1;
    const bool sorted = h->hji_sort != 0 && (h->hji_sort > 0 || M >= (1 << 19)) && ncell + 1 < (1ll << 31);
    if (sorted) {
        // work space: counters per cell (+ the outside bucket), block totals of the scan, keys and permutation per query
        const long long nc = ncell + 1;
        const int nb = (int)((nc + 2047) / 2048);
        const size_t need = (size_t)(nc + nb + 8) * 4 + (size_t)M * 8;
        if (h->hji_ws_bytes < need) {
            if (h->d_hji_ws) { cudaStreamSynchronize(h->stream); cudaFree(h->d_hji_ws); h->d_hji_ws = nullptr; h->hji_ws_bytes = 0; }
            if (cudaMalloc(&h->d_hji_ws, need) == cudaSuccess) h->hji_ws_bytes = need;
        }
        if (h->d_hji_ws) {
            uint32_t* cnt = (uint32_t*)h->d_hji_ws; uint32_t* totals = cnt + nc; uint32_t* keys = totals + nb + 8; uint32_t* perm = keys + M;
            cudaMemsetAsync(cnt, 0, (size_t)nc * 4, h->stream);
            k_hji_keys<<<(M + 255) / 256, 256, 0, h->stream>>>(h->hji, M, ncell, d_x, keys, cnt);
            k_scan_blocks<<<nb, 256, 0, h->stream>>>(cnt, nc, totals);
            k_scan_totals<<<1, 1024, 0, h->stream>>>(totals, nb);
            k_scan_add<<<nb, 256, 0, h->stream>>>(cnt, nc, totals);
            k_hji_scatter<<<(M + 255) / 256, 256, 0, h->stream>>>(M, keys, cnt, perm);
            // The TMA-staged tiles are built, parity-green and MEASURED, and they lose: 2^24 queries 13.4 ms against 9.3 ms for the plain
            // cell-ordered gather (2^22: 4.6 / 2.2 ms; 2^20: 3.8 / 0.77 ms).  66.6 KB of tile per CTA leave 3 CTAs = 12 warps per SM, and a query
            // is a chain of 64 dependent FP64 accumulations: too few warps to hide it, while the L2 gather runs 48+ warps per SM; the tile loads
            // themselves (98 304 x 66.6 KB, not overlapped with the interpolation) cost 3.5 ms.  Kept behind pgn_set_hji_lookup_order(2).
            const bool tma = h->hji_tma_valid && h->hji_sort == 2;
            if (tma) {
                k_hji_lookup_tma<<<(unsigned)h->hji_tma_blocks, 128, h->hji_tma_tile_bytes, h->stream>>>(*reinterpret_cast<const CUtensorMap*>(h->hji_tmap), h->hji, M, cnt, perm, d_x, d_V, d_gV);
                k_hji_outside<<<(M + 255) / 256, 256, 0, h->stream>>>(M, cnt, nc, perm, d_V, d_gV);
                h->launches += 7;
            } else {
                k_hji_lookup_perm<<<(M + 127) / 128, 128, 0, h->stream>>>(h->hji, M, perm, d_x, d_V, d_gV);
                h->launches += 6;
            }
            return;
        }
    }
    k_hji_lookup<<<(M + 127) / 128, 128, 0, h->stream>>>(h->hji, M, d_x, d_V, d_gV);
    h->launches++;
}

}  // namespace pgn
