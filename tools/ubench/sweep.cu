// Micro-benchmark of the dense-tail sweep (S <- -S^-1, packed lower, Dm = 59) with parts switched off, 512 threads, one CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define T 512
#define QPT 2
template <int MODE>   // bit0: no division (constant reciprocal), bit1: no staging stores, bit2: no update arithmetic, bit3: no barrier
__global__ void __launch_bounds__(T, 1) k(int Dm, long long* out, double* sink, int reps) {
    __shared__ double S[64 * 65 / 2];
    __shared__ double red[160];
    const int tid = threadIdx.x;
    const int npk = Dm * (Dm + 1) / 2;
    for (int e = tid; e < npk; e += T) S[e] = 0.01 * ((e * 7) % 13);
    __syncthreads();
    for (int i = tid; i < Dm; i += T) S[i * (i + 1) / 2 + i] = 10.0 + i;
    __syncthreads();
    double* col = red; double* dpiv = red + 128;
    const int nblk = (Dm + 3) >> 2, nquad = 2 * nblk * (nblk + 1);
    int qi[QPT], qk[QPT]; double v[QPT][4];
    for (int x = 0; x < QPT; x++) {
        const int g = tid + x * T;
        int b = (int)((sqrt(1.0 + 2.0 * g) - 1.0) * 0.5);
        b += (2 * (b + 1) * (b + 2) <= g); b -= (2 * b * (b + 1) > g);
        const int rem = g - 2 * b * (b + 1);
        const int i = 4 * b + rem / (b + 1), k0 = 4 * (rem % (b + 1));
        const bool live = g < nquad && i < Dm;
        qi[x] = live ? i : -8; qk[x] = k0;
        for (int j = 0; j < 4; j++) v[x][j] = (live && k0 + j <= i) ? S[i * (i + 1) / 2 + k0 + j] : 0.0;
    }
    __syncthreads();
    long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        if (tid < Dm) col[tid] = S[tid * (tid + 1) / 2];
        if (tid == 0) dpiv[0] = 1.0 / S[0];
        __syncthreads();
        for (int p = 0; p < Dm; p++) {
            const double* cc = col + (p & 1) * 64;
            double* cn = col + ((p + 1) & 1) * 64;
            const double dinv = dpiv[p & 1];
            const int pq = p & ~3, pj = p & 3, nq = (p + 1) & ~3, nj = (p + 1) & 3;
#pragma unroll
            for (int x = 0; x < QPT; x++) {
                const int i = qi[x], k0 = qk[x];
                if (i < 0) continue;
                const double ci = cc[i], t = ci * dinv;
                const double c0 = cc[k0], c1 = cc[k0 + 1], c2 = cc[k0 + 2], c3 = cc[k0 + 3];
                if (!(MODE & 4)) {
                    v[x][0] = fma(-t, c0, v[x][0]); v[x][1] = fma(-t, c1, v[x][1]); v[x][2] = fma(-t, c2, v[x][2]); v[x][3] = fma(-t, c3, v[x][3]);
                    if (i == p) { v[x][0] = c0 * dinv; v[x][1] = c1 * dinv; v[x][2] = c2 * dinv; v[x][3] = c3 * dinv; }
                    if (k0 == pq) {
                        const double cv = (i == p) ? -dinv : t;
                        v[x][0] = pj == 0 ? cv : v[x][0]; v[x][1] = pj == 1 ? cv : v[x][1]; v[x][2] = pj == 2 ? cv : v[x][2]; v[x][3] = pj == 3 ? cv : v[x][3];
                    }
                }
                if (!(MODE & 2)) {
                    if (i == p + 1) {
                        if (k0 <= i) cn[k0] = v[x][0];
                        if (k0 + 1 <= i) cn[k0 + 1] = v[x][1];
                        if (k0 + 2 <= i) cn[k0 + 2] = v[x][2];
                        if (k0 + 3 <= i) cn[k0 + 3] = v[x][3];
                    }
                    if (k0 == nq && i >= p + 1) {
                        const double nv = nj == 0 ? v[x][0] : (nj == 1 ? v[x][1] : (nj == 2 ? v[x][2] : v[x][3]));
                        cn[i] = nv;
                        if (i == p + 1) dpiv[(p + 1) & 1] = (MODE & 1) ? 0.1 : 1.0 / nv;
                    }
                }
            }
            if (!(MODE & 8)) __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[MODE] = (t1 - t0) / reps;
    double acc = 0;
    for (int x = 0; x < QPT; x++) for (int j = 0; j < 4; j++) acc += v[x][j];
    sink[tid] = acc;
}
int main() {
    long long* out; cudaMallocManaged(&out, 16 * 8); double* sink; cudaMalloc(&sink, 8 * T);
    k<0><<<148, T>>>(59, out, sink, 20); k<1><<<148, T>>>(59, out, sink, 20); k<2><<<148, T>>>(59, out, sink, 20); k<3><<<148, T>>>(59, out, sink, 20);
    k<7><<<148, T>>>(59, out, sink, 20); k<15><<<148, T>>>(59, out, sink, 20); k<4><<<148, T>>>(59, out, sink, 20);
    cudaDeviceSynchronize();
    printf("cycles per sweep of 59 pivots: full %lld | no division %lld | no staging %lld | neither %lld | no arithmetic either %lld | no barrier either %lld | staging+division only %lld   %s\n",
           out[0], out[1], out[2], out[3], out[7], out[15], out[4], cudaGetErrorString(cudaGetLastError()));
}
