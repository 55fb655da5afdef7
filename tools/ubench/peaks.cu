// Measured peaks the ADMM roofline divides by (SURVEY.md 8d asks for them; MEASURED_PEAKS.json only holds HBM and bf16):
//   FP64 FMA throughput (DFMA, 8 independent chains per thread, 148 x 2 CTAs x 1024 threads)        -> TFLOP/s
//   shared-memory read throughput (conflict-free LDS.64 and LDS.128, 1024 threads per SM)           -> bytes / clock / SM and TB/s
// Prints one JSON object; bench.py runs this binary (when present) and uses the numbers instead of nominal ones.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peaks peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_dfma(double* sink, int n) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}
template <int W>
__global__ void __launch_bounds__(1024) k_lds(double* sink, int n, long long* cyc) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 1024) sm[i] = i;
    __syncthreads();
    double acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (W == 8) acc += sm[(threadIdx.x + 1024 * u + i) & 4095];
            else { const double2 v = reinterpret_cast<const double2*>(sm)[(threadIdx.x + 512 * u + i) & 2047]; acc += v.x + v.y; }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 123.456) sink[0] = acc;
}

int main() {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 64); cudaMallocManaged(&cyc, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // FP64
    const int n = 4000;
    k_dfma<<<2 * sms, 1024>>>(sink, 100);
    double best = 0;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k_dfma<<<2 * sms, 1024>>>(sink, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 64.0 * n * 1024.0 * 2 * sms / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    // shared memory
    double b8 = 0, b16 = 0, tb8 = 0, tb16 = 0;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k_lds<8><<<sms, 1024>>>(sink, 2000, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double v = 8.0 * 8 * 2000 * 1024 / (double)cyc[0], t = 8.0 * 8 * 2000 * 1024 * sms / (ms * 1e-3) / 1e12;
        if (v > b8) b8 = v; if (t > tb8) tb8 = t;
        cudaEventRecord(e0); k_lds<16><<<sms, 1024>>>(sink, 2000, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        v = 16.0 * 8 * 2000 * 1024 / (double)cyc[0]; t = 16.0 * 8 * 2000 * 1024 * sms / (ms * 1e-3) / 1e12;
        if (v > b16) b16 = v; if (t > tb16) tb16 = t;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"smem_lds64_bytes_per_clk_sm\": %.1f, \"smem_lds128_bytes_per_clk_sm\": %.1f, "
           "\"smem_lds64_tbs\": %.2f, \"smem_lds128_tbs\": %.2f, \"status\": \"%s\"}\n", p.name, sms, best, b8, b16, tb8, tb16, cudaGetErrorString(e));
    return e != cudaSuccess;
}
