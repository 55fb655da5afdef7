// Probe: largest dynamic shared memory per CTA that still lets TWO 256-thread CTAs share an SM (B200), with and without the MaxShared carve-out.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o occ occ.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(double* out) { extern __shared__ double sm[]; sm[threadIdx.x] = threadIdx.x; __syncthreads(); if (out) out[threadIdx.x] = sm[255 - threadIdx.x]; }
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int resv = 0; cudaDeviceGetAttribute(&resv, cudaDevAttrReservedSharedMemoryPerBlock, 0);
    printf("sharedMemPerMultiprocessor %zu, sharedMemPerBlockOptin %zu, reserved per block %d, regsPerMultiprocessor %d\n", p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, resv, p.regsPerMultiprocessor);
    for (int carve = 0; carve < 2; carve++) {
        int last2 = 0;
        for (int kb = 90; kb <= 116; kb++) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
            if (carve) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, kb * 1024);
            if (nb >= 2) last2 = kb;
        }
        printf("carve-out %s: 2 CTAs/SM up to %d KB of dynamic shared memory per CTA\n", carve ? "MaxShared" : "default", last2);
    }
    // finer probe around the limit
    for (int b = 112 * 1024; b <= 116 * 1024; b += 256) { int nb = 0; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, b); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, b); if (nb < 2) { printf("first size with 1 CTA/SM: %d bytes\n", b); break; } }
    return 0;
}
