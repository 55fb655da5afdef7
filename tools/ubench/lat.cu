// Micro-benchmarks of the latencies that bound the ADMM kernel (one CTA per SM, 512 threads): barrier, dependent LDS, DFMA chain, SHFL chain,
// FP64 reciprocal / sqrt.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int threads_active, long long* out, double* sink, int n) {
    __shared__ double sm[2048];
    __shared__ int idx[2048];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; idx[i] = (i * 37 + 11) & 2047; }
    __syncthreads();
    long long t0, t1;
    double acc = sink[0];
    // 1. barrier only
    t0 = clock64();
    for (int i = 0; i < n; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / n;
    // 2. dependent LDS chain (int index chase)
    int j = tid & 2047;
    t0 = clock64();
    for (int i = 0; i < n; i++) j = idx[j];
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / n;
    acc += j;
    // 3. dependent LDS.64 gather + DFMA
    t0 = clock64();
    for (int i = 0; i < n; i++) { acc = fma(sm[(j + i) & 2047], 1.0000001, acc); }
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / n;
    // 4. DFMA dependent chain
    double a = acc, b = 1.0000001;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) a = fma(a, b, 1e-9);
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / n;
    // 5. SHFL (double) + DADD chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; i++) a += __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / n;
    // 6. FP64 reciprocal chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; i++) a = 1.0 / (a + 1.5);
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / n;
    // 7. FP64 sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; i++) a = sqrt(a + 1.5);
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0) / n;
    // 8. barrier + one LDS + DFMA + STS per phase (the shape of an empty-ish phase)
    t0 = clock64();
    for (int i = 0; i < n; i++) { sm[tid] = fma(sm[(tid + 33) & 2047], b, a); __syncthreads(); }
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0) / n;
    // 9. DADD dependent chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) a = a + b;
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0) / n;
    // 10. int ALU chain (IMAD)
    int q = j;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) q = q * 3 + i;
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0) / n;
    sink[tid] = a + acc + q;
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 128); cudaMalloc(&sink, 8 * 1024); cudaMemset(sink, 0, 8 * 1024);
    const char* names[] = {"__syncthreads", "LDS chase", "LDS.64+DFMA dep", "DFMA chain", "SHFL.f64+DADD", "1/x f64", "sqrt f64", "LDS+DFMA+STS+bar", "DADD chain", "IMAD chain"};
    for (int threads : {32, 128, 256, 512, 1024}) {
        k<<<148, threads>>>(threads, out, sink, 2000);
        cudaDeviceSynchronize();
        printf("threads %4d:", threads);
        for (int i = 0; i < 10; i++) printf("  %s %lld", names[i], out[i]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
