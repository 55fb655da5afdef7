// Dense-tail sweep with 4x4 register tiles: thread = tile (bi >= bk) of the block lower triangle, full symmetric tiles (no masks), named
// barrier over the participating warps only.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define T 512
__device__ void init_S(double* S, int Dm) {
    const int npk = Dm * (Dm + 1) / 2;
    for (int e = threadIdx.x; e < npk; e += T) S[e] = 0.01 * ((e * 7) % 13) - 0.05;
    __syncthreads();
    for (int i = threadIdx.x; i < Dm; i += T) S[i * (i + 1) / 2 + i] = (i & 1) ? -(10.0 + i) : (10.0 + i);
    __syncthreads();
}
__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__global__ void __launch_bounds__(T, 1) kC(int Dm, long long* out, double* Sout, int reps) {
    __shared__ double S[64 * 65 / 2];
    __shared__ __align__(16) double red[160];
    const int tid = threadIdx.x;
    double* col = red; double* dpiv = red + 128;
    const int nb = (Dm + 3) >> 2, ntile = nb * (nb + 1) / 2;
    const int nthr = (ntile + 31) & ~31;                  // participating threads (whole warps)
    int bi = (int)((sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
    bi += ((bi + 1) * (bi + 2) / 2 <= tid); bi -= (bi * (bi + 1) / 2 > tid);
    const int bk = tid - bi * (bi + 1) / 2;
    const bool live = tid < ntile;
    const int i0 = 4 * bi, k0 = 4 * bk;
    long long total = 0;
    for (int rep = 0; rep < reps; rep++) {
        init_S(S, Dm);
        double v[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int i = i0 + r, k = k0 + c, hi = max(i, k), lo = min(i, k);
                v[r][c] = (live && hi < Dm) ? S[hi * (hi + 1) / 2 + lo] : 0.0;
            }
        __syncthreads();
        long long t0 = clock64();
        if (tid < 64) col[tid] = tid < Dm ? S[tid * (tid + 1) / 2] : 0.0;
        if (tid == 0) dpiv[0] = 1.0 / S[0];
        __syncthreads();
        if (tid < nthr) {
            for (int p = 0; p < Dm; p++) {
                const double* cc = col + (p & 1) * 64;
                double* cn = col + ((p + 1) & 1) * 64;
                if (live) {
                    const double dinv = dpiv[p & 1];
                    const double2 ra = *reinterpret_cast<const double2*>(cc + i0), rb = *reinterpret_cast<const double2*>(cc + i0 + 2);
                    const double2 ca = *reinterpret_cast<const double2*>(cc + k0), cb = *reinterpret_cast<const double2*>(cc + k0 + 2);
                    const double ci[4] = {ra.x, ra.y, rb.x, rb.y}, ck[4] = {ca.x, ca.y, cb.x, cb.y};
                    double t[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) t[r] = ci[r] * dinv;
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) v[r][c] = fma(-t[r], ck[c], v[r][c]);
                    const int pb = p >> 2, pj = p & 3;
                    if (bi == pb) {                       // pivot row inside this tile: S_pk <- S_pk / d
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (r == pj) {
#pragma unroll
                                for (int c = 0; c < 4; c++) v[r][c] = ck[c] * dinv;
                            }
                    }
                    if (bk == pb) {                       // pivot column inside this tile: S_ip <- S_ip / d, S_pp <- -1 / d
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            if (c == pj) {
#pragma unroll
                                for (int r = 0; r < 4; r++) v[r][c] = (bi == pb && r == pj) ? -dinv : t[r];
                            }
                    }
                    const int nbk = (p + 1) >> 2, nj = (p + 1) & 3;
                    if (bk == nbk) {                      // next pivot column, rows of this tile
                        double w[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) w[r] = nj == 0 ? v[r][0] : (nj == 1 ? v[r][1] : (nj == 2 ? v[r][2] : v[r][3]));
                        *reinterpret_cast<double2*>(cn + i0) = make_double2(w[0], w[1]);
                        *reinterpret_cast<double2*>(cn + i0 + 2) = make_double2(w[2], w[3]);
                        if (bi == nbk) dpiv[(p + 1) & 1] = 1.0 / (nj == 0 ? w[0] : (nj == 1 ? w[1] : (nj == 2 ? w[2] : w[3])));
                    }
                    if (bi == nbk && bk < nbk) {          // next pivot row (mirror image of the column left of the diagonal tile)
                        double w[4];
#pragma unroll
                        for (int c = 0; c < 4; c++) w[c] = nj == 0 ? v[0][c] : (nj == 1 ? v[1][c] : (nj == 2 ? v[2][c] : v[3][c]));
                        *reinterpret_cast<double2*>(cn + k0) = make_double2(w[0], w[1]);
                        *reinterpret_cast<double2*>(cn + k0 + 2) = make_double2(w[2], w[3]);
                    }
                }
                bar_named(1, nthr);
            }
        }
        long long t1 = clock64();
        total += t1 - t0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int i = i0 + r, k = k0 + c;
                if (live && k <= i && i < Dm) S[i * (i + 1) / 2 + k] = v[r][c];
            }
        __syncthreads();
    }
    if (tid == 0 && blockIdx.x == 0) out[1] = total / reps;
    if (blockIdx.x == 0) for (int e = tid; e < Dm * (Dm + 1) / 2; e += T) Sout[e] = S[e];
}
__global__ void __launch_bounds__(T, 1) kRef(int Dm, double* Sout) {
    __shared__ double S[64 * 65 / 2];
    __shared__ double cc[64];
    init_S(S, Dm);
    const int npk = Dm * (Dm + 1) / 2;
    for (int p = 0; p < Dm; p++) {
        if (threadIdx.x < Dm) { int i = threadIdx.x; cc[i] = i >= p ? S[i * (i + 1) / 2 + p] : S[p * (p + 1) / 2 + i]; }
        __syncthreads();
        const double dinv = 1.0 / cc[p];
        for (int e = threadIdx.x; e < npk; e += T) {
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5); i += ((i + 1) * (i + 2) / 2 <= e); i -= (i * (i + 1) / 2 > e);
            const int k = e - i * (i + 1) / 2;
            double v;
            if (i == p && k == p) v = -dinv; else if (i == p) v = cc[k] * dinv; else if (k == p) v = cc[i] * dinv; else v = S[e] - cc[i] * cc[k] * dinv;
            S[e] = v;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0) for (int e = threadIdx.x; e < npk; e += T) Sout[e] = S[e];
}
int main() {
    long long* out; cudaMallocManaged(&out, 64); double *SB, *SR; cudaMallocManaged(&SB, 8 * 2080); cudaMallocManaged(&SR, 8 * 2080);
    for (int Dm : {59, 64, 17, 3}) {
        kC<<<148, T>>>(Dm, out, SB, 20); kRef<<<1, T>>>(Dm, SR); cudaDeviceSynchronize();
        double err = 0, mx = 0; for (int e = 0; e < Dm * (Dm + 1) / 2; e++) { err = fmax(err, fabs(SB[e] - SR[e])); mx = fmax(mx, fabs(SR[e])); }
        printf("Dm %d: 4x4 tiles %lld cycles per sweep (%lld per pivot); max |tile - ref| = %.3e (max |ref| %.3e)  %s\n", Dm, out[1], out[1] / Dm, err, mx, cudaGetErrorString(cudaGetLastError()));
    }
}
