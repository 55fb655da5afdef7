// Micro-benchmark: cost of special-register reads (%ctaid.x), indexed kernel-parameter reads and clock reads inside a loop, 512 threads/CTA.
#include <cstdio>
#include <cuda_runtime.h>
struct Args { unsigned short ph[64]; long long* out; int n; unsigned long long* cyc; };
__global__ void k(const Args a, int* sink) {
    const int tid = threadIdx.x;
    long long t0, t1;
    int acc = 0;
    t0 = clock64();
    for (int i = 0; i < a.n; i++) { unsigned v; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(v)); acc += v + i; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) a.out[0] = (t1 - t0) / a.n;
    t0 = clock64();
    for (int i = 0; i < a.n; i++) { acc += a.ph[(acc + i) & 63]; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) a.out[1] = (t1 - t0) / a.n;
    t0 = clock64();
    for (int i = 0; i < a.n; i++) { acc += (int)clock64(); }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) a.out[2] = (t1 - t0) / a.n;
    // the pattern of the profiling macro: predicate (pointer != 0 && tid == 0 && blockIdx.x == 0) evaluated every trip, body never taken
    t0 = clock64();
    for (int i = 0; i < a.n; i++) {
        if (a.cyc && threadIdx.x == 0 && blockIdx.x == 1000) { a.cyc[i & 7] += clock64(); }
        acc = acc * 3 + i;
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) a.out[3] = (t1 - t0) / a.n;
    t0 = clock64();
    for (int i = 0; i < a.n; i++) { acc = acc * 3 + i; __syncthreads(); }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) a.out[4] = (t1 - t0) / a.n;
    sink[tid] = acc;
}
int main() {
    Args a; for (int i = 0; i < 64; i++) a.ph[i] = i * 7 % 64;
    cudaMallocManaged(&a.out, 128); a.n = 2000; cudaMalloc(&a.cyc, 64);
    int* sink; cudaMalloc(&sink, 4096);
    k<<<148, 512>>>(a, sink);
    cudaDeviceSynchronize();
    printf("ctaid read %lld   indexed param LDC (dependent) %lld   clock64 %lld   barrier+untaken profiling predicate %lld   barrier only %lld   %s\n", a.out[0], a.out[1], a.out[2], a.out[3], a.out[4], cudaGetErrorString(cudaGetLastError()));
}
