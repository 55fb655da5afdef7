// Dense-tail sweep, two register layouts compared (cycles and results):
//   A: four contiguous columns per thread (cc[k0 + j]: lanes 32 B apart -> 4-way bank conflicts on every column load)
//   B: eight columns of stride 8 per thread, thread = (row i = tid / 8, q = tid % 8), columns q + 8 j (cc[q + 8 j]: 64 contiguous bytes per load)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define T 512
__device__ void init_S(double* S, int Dm) {
    const int npk = Dm * (Dm + 1) / 2;
    for (int e = threadIdx.x; e < npk; e += T) S[e] = 0.01 * ((e * 7) % 13) - 0.05;
    __syncthreads();
    for (int i = threadIdx.x; i < Dm; i += T) S[i * (i + 1) / 2 + i] = (i & 1) ? -(10.0 + i) : (10.0 + i);      // quasi-definite
    __syncthreads();
}
__global__ void __launch_bounds__(T, 1) kB(int Dm, long long* out, double* Sout, int reps) {
    __shared__ double S[64 * 65 / 2];
    __shared__ double red[160];
    const int tid = threadIdx.x;
    double* col = red; double* dpiv = red + 128;
    const int i = tid >> 3, q = tid & 7;
    const bool live = i < Dm && q <= i;
    long long total = 0;
    for (int rep = 0; rep < reps; rep++) {
        init_S(S, Dm);
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (live && q + 8 * j <= i) ? S[i * (i + 1) / 2 + q + 8 * j] : 0.0;
        __syncthreads();
        long long t0 = clock64();
        if (tid < Dm) col[tid] = S[tid * (tid + 1) / 2];
        if (tid == 0) dpiv[0] = 1.0 / S[0];
        __syncthreads();
        for (int p = 0; p < Dm; p++) {
            const double* cc = col + (p & 1) * 64;
            double* cn = col + ((p + 1) & 1) * 64;
            if (live) {
                const double dinv = dpiv[p & 1];
                const double ci = cc[i], t = ci * dinv;
                double c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) c[j] = cc[q + 8 * j];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fma(-t, c[j], v[j]);
                if (i == p) {
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = c[j] * dinv;
                }
                if (q == (p & 7)) {
                    const double cv = (i == p) ? -dinv : t;
                    const int pj = p >> 3;
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = (pj == j) ? cv : v[j];
                }
                if (i == p + 1) {
#pragma unroll
                    for (int j = 0; j < 8; j++) if (q + 8 * j <= i) cn[q + 8 * j] = v[j];
                }
                if (q == ((p + 1) & 7) && i >= p + 1) {
                    const int nj = (p + 1) >> 3;
                    double nv = v[0];
#pragma unroll
                    for (int j = 1; j < 8; j++) nv = (nj == j) ? v[j] : nv;
                    cn[i] = nv;
                    if (i == p + 1) dpiv[(p + 1) & 1] = 1.0 / nv;
                }
            }
            __syncthreads();
        }
        long long t1 = clock64();
        total += t1 - t0;
#pragma unroll
        for (int j = 0; j < 8; j++) if (live && q + 8 * j <= i) S[i * (i + 1) / 2 + q + 8 * j] = v[j];
        __syncthreads();
    }
    if (tid == 0 && blockIdx.x == 0) out[1] = total / reps;
    if (blockIdx.x == 0) for (int e = tid; e < Dm * (Dm + 1) / 2; e += T) Sout[e] = S[e];
}
// reference: plain in-place sweep on shared memory (two barriers per pivot), for the numerical comparison
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
    for (int Dm : {59, 64, 17}) {
        kB<<<148, T>>>(Dm, out, SB, 20); kRef<<<1, T>>>(Dm, SR); cudaDeviceSynchronize();
        double err = 0, mx = 0; for (int e = 0; e < Dm * (Dm + 1) / 2; e++) { err = fmax(err, fabs(SB[e] - SR[e])); mx = fmax(mx, fabs(SR[e])); }
        printf("Dm %d: layout B %lld cycles per sweep (%lld per pivot); max |B - ref| = %.3e (max |ref| %.3e)  %s\n", Dm, out[1], out[1] / Dm, err, mx, cudaGetErrorString(cudaGetLastError()));
    }
}
