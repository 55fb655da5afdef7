// What does one "barrier + a few shared loads + a few FP64 FMAs" step cost with 16 warps?  Variants isolate loads, FMAs, predicated patches.
#include <cstdio>
#include <cuda_runtime.h>
#define T 512
template <int V>
__global__ void __launch_bounds__(T, 1) k(long long* out, double* sink, int n) {
    __shared__ double cc[128];
    __shared__ double dp[2];
    const int tid = threadIdx.x;
    if (tid < 128) cc[tid] = 1.0 + 1e-3 * tid;
    if (tid < 2) dp[tid] = 0.999;
    __syncthreads();
    const int i = (tid >> 3) & 63, k0 = (tid & 7) * 4 & 63, q = tid & 7;
    double v0 = tid, v1 = tid + 1, v2 = tid + 2, v3 = tid + 3;
    long long t0 = clock64();
    for (int p = 0; p < n; p++) {
        const double* c = cc + (p & 1) * 64;
        if (V == 1) { v0 = fma(c[tid & 63], 0.5, v0); }
        if (V == 2 || V == 3) {
            const double dinv = dp[p & 1];
            const double ci = c[i], t = ci * dinv;
            const double c0 = c[k0], c1 = c[k0 + 1], c2 = c[k0 + 2], c3 = c[k0 + 3];
            v0 = fma(-t, c0, v0); v1 = fma(-t, c1, v1); v2 = fma(-t, c2, v2); v3 = fma(-t, c3, v3);
            if (V == 3) {
                const int pp = p & 63;
                if (i == pp) { v0 = c0 * dinv; v1 = c1 * dinv; v2 = c2 * dinv; v3 = c3 * dinv; }
                if (k0 == (pp & ~3)) { const double cv = (i == pp) ? -dinv : t; const int pj = pp & 3; v0 = pj == 0 ? cv : v0; v1 = pj == 1 ? cv : v1; v2 = pj == 2 ? cv : v2; v3 = pj == 3 ? cv : v3; }
            }
        }
        if (V == 4) {
            const double dinv = dp[p & 1];
            const double ci = c[i], t = ci * dinv;
            const double c0 = c[q], c1 = c[q + 8], c2 = c[q + 16], c3 = c[q + 24];
            v0 = fma(-t, c0, v0); v1 = fma(-t, c1, v1); v2 = fma(-t, c2, v2); v3 = fma(-t, c3, v3);
        }
        if (V == 5) {   // loads only, folded by integer adds of the low words (no FP64)
            const double ci = c[i]; const double c0 = c[k0], c1 = c[k0 + 1], c2 = c[k0 + 2], c3 = c[k0 + 3];
            v0 = __longlong_as_double(__double_as_longlong(v0) + __double_as_longlong(ci) + __double_as_longlong(c0) + __double_as_longlong(c1) + __double_as_longlong(c2) + __double_as_longlong(c3));
        }
        if (V == 6) {   // FMAs only (no loads)
            const double t = v3 * 1e-9;
            v0 = fma(-t, 1.0000001, v0); v1 = fma(-t, 1.0000002, v1); v2 = fma(-t, 1.0000003, v2); v3 = fma(-t, 1.0000004, v3);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[V] = (t1 - t0) / n;
    sink[tid] = v0 + v1 + v2 + v3;
}
int main() {
    long long* out; cudaMallocManaged(&out, 128); double* sink; cudaMalloc(&sink, 8 * T);
    k<0><<<148, T>>>(out, sink, 4000); k<1><<<148, T>>>(out, sink, 4000); k<2><<<148, T>>>(out, sink, 4000); k<3><<<148, T>>>(out, sink, 4000);
    k<4><<<148, T>>>(out, sink, 4000); k<5><<<148, T>>>(out, sink, 4000); k<6><<<148, T>>>(out, sink, 4000);
    cudaDeviceSynchronize();
    printf("cycles per step: barrier %lld | 1 LDS + 1 DFMA %lld | 6 LDS + DMUL + 4 DFMA (32 B lane stride) %lld | + pivot patches %lld | same loads, 8 B lane stride %lld | loads only %lld | FMAs only %lld  %s\n",
           out[0], out[1], out[2], out[3], out[4], out[5], out[6], cudaGetErrorString(cudaGetLastError()));
}
