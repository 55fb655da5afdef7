// Latency of special-register reads that feed a shared-memory address (the compiler re-materialises shared-window bases with
// S2UR SR_CgaCtaId inside loops): pointer chase through shared memory with the register read in the address chain.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int n, int* sink) {
    __shared__ int sm[1024];
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += blockDim.x) sm[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    long long t0, t1;
    int idx = tid;
    t0 = clock64();
    for (int i = 0; i < n; i++) idx = sm[idx & 1023];
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; i++) { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); idx = sm[(idx + r) & 1023]; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[1] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; i++) { unsigned r; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(r)); idx = sm[(idx + r) & 1023]; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[2] = (t1 - t0) / n;
    // register read made dependent on the previous load (true serial latency)
    t0 = clock64();
    for (int i = 0; i < n; i++) { unsigned r; asm volatile("{ .reg .pred p; setp.ge.s32 p, %1, 0; @p mov.u32 %0, %%cluster_ctarank; @!p mov.u32 %0, 0; }" : "=r"(r) : "r"(idx)); idx = sm[(idx + r) & 1023]; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[3] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; i++) { unsigned r; asm volatile("{ .reg .pred p; setp.ge.s32 p, %1, 0; @p mov.u32 %0, %%ctaid.x; @!p mov.u32 %0, 0; }" : "=r"(r) : "r"(idx)); idx = sm[(idx + r) & 1023]; }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[4] = (t1 - t0) / n;
    sink[tid] = idx;
}
int main() {
    long long* out; cudaMallocManaged(&out, 128); int* sink; cudaMalloc(&sink, 4096);
    for (int th : {32, 512}) {
        k<<<148, th>>>(out, 2000, sink); cudaDeviceSynchronize();
        printf("threads %d: LDS chase %lld | + cluster_ctarank (independent) %lld | + ctaid (independent) %lld | cluster_ctarank serial %lld | ctaid serial %lld  %s\n", th, out[0], out[1], out[2], out[3], out[4], cudaGetErrorString(cudaGetLastError()));
    }
}
