// Probe: do TWO CTAs that each allocate 256 tensor-memory columns (and ~110 KB of shared memory) really share an SM?  The occupancy API answers
// 1 for kernels that contain tcgen05.alloc; this measures the actual co-residency with a per-SM counter.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o occ_tmem occ_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int USE_TMEM>
__global__ void __launch_bounds__(256, 2) k(int* per_sm, int* max_seen, long long spin) {
    extern __shared__ double sm[];
    __shared__ uint32_t tb;
    if (USE_TMEM) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tb)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) { const int c = atomicAdd(per_sm + smid, 1) + 1; atomicMax(max_seen, c); }
    sm[threadIdx.x] = threadIdx.x;
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    __syncthreads();
    if (threadIdx.x == 0) atomicSub(per_sm + smid, 1);
    if (USE_TMEM) { __syncthreads(); if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tb) : "memory"); }
}
int main() {
    int *per_sm, *mx; cudaMallocManaged(&per_sm, 1024 * 4); cudaMallocManaged(&mx, 4);
    for (int use = 0; use < 2; use++) for (int kb : {100, 108, 112}) {
        for (int i = 0; i < 1024; i++) per_sm[i] = 0; *mx = 0;
        const int smem = kb * 1024; int nb = 0;
        if (use) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<1>, 256, smem); k<1><<<296, 256, smem>>>(per_sm, mx, 2000000); }
        else { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<0>, 256, smem); k<0><<<296, 256, smem>>>(per_sm, mx, 2000000); }
        cudaError_t e = cudaDeviceSynchronize();
        printf("tmem %d, %d KB smem: occupancy API %d CTAs/SM, measured max co-resident CTAs per SM %d (%s)\n", use, kb, nb, *mx, cudaGetErrorString(e));
    }
    return 0;
}
