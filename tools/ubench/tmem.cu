// Can Tensor Memory hold the L factor of the ADMM kernel?  TMEM (256 KB per SM, 128 lanes x 512 columns x 32 bit) is reachable from ordinary
// threads with tcgen05.st / tcgen05.ld: warp w owns the 32 lanes of quadrant w % 4, and the shape 32x32b gives lane i of the warp the
// columns [c, c + n) of TMEM lane 32 (w % 4) + i — exactly the slot-major layout of the warp programs (slot k of lane l at ebase + 32 k + l).
// This benchmark checks, with TWO 256-thread CTAs per SM allocating 256 columns each (110 KB of dynamic shared memory per CTA to pin the occupancy):
//   1. a double round trip through tcgen05.st / tcgen05.ld is bit exact, also for the second warp of a quadrant (warps w and w + 4);
//   2. the latency of a dependent chain  LDTM.x8 (four doubles) -> wait::ld -> 4 DFMA  against  4 x LDS.64 -> 4 DFMA;
//   3. the read throughput of LDTM.x8 with all eight warps streaming.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem tmem.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st2(uint32_t taddr, double v) {
    const unsigned long long b = __double_as_longlong(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double& a, double& b, double& c, double& d) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __longlong_as_double((long long)(((unsigned long long)r1 << 32) | r0)); b = __longlong_as_double((long long)(((unsigned long long)r3 << 32) | r2));
    c = __longlong_as_double((long long)(((unsigned long long)r5 << 32) | r4)); d = __longlong_as_double((long long)(((unsigned long long)r7 << 32) | r6));
}
// the same without the wait: the caller waits once for several loads in flight
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(256, 2) k(long long* out, double* sink, int n, int* bad) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ uint32_t tbase_s;
    double* sm = reinterpret_cast<double*>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // warp w: lanes of quadrant w % 4, columns [128 (w / 4), +128) of the CTA's 256 = 64 doubles per thread
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (warp >> 2);
    for (int i = tid; i < 4096; i += 256) sm[i] = 1.0 + 1e-9 * i;
    for (int i = tid; i < 4096; i += 256) reinterpret_cast<unsigned short*>(raw + 65536)[i] = (unsigned short)((i * 37 + 11) & 2047);
    for (int k2 = 0; k2 < 64; k2++) tmem_st2(my + 2 * k2, 1.0 + 1e-3 * (warp * 64 + k2) + 1e-7 * lane);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    // 1. round trip
    int nbad = 0;
    for (int k4 = 0; k4 < 16; k4++) {
        double a, b, c, d;
        tmem_ld8(my + 8 * k4, a, b, c, d);
        const double e0 = 1.0 + 1e-3 * (warp * 64 + 4 * k4) + 1e-7 * lane;
        nbad += (a != e0) + (b != 1.0 + 1e-3 * (warp * 64 + 4 * k4 + 1) + 1e-7 * lane) + (c != 1.0 + 1e-3 * (warp * 64 + 4 * k4 + 2) + 1e-7 * lane) +
                (d != 1.0 + 1e-3 * (warp * 64 + 4 * k4 + 3) + 1e-7 * lane);
    }
    if (nbad) atomicAdd(bad, nbad);
    long long t0, t1;
    double acc = sink[0];
    // 2a. dependent chain through TMEM: the column of the next load depends on the previous result
    uint32_t col = 0;
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        double a, b, c, d;
        tmem_ld8(my + col, a, b, c, d);
        acc = fma(a, 1.0000001, acc); acc = fma(b, 1.0000001, acc); acc = fma(c, 1.0000001, acc); acc = fma(d, 1.0000001, acc);
        col = (uint32_t)(__double2int_rd(acc) & 0) + ((col + 8) & 127);
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / n;
    // 2b. the same through shared memory (4 x LDS.64 at lane stride, then 4 DFMA)
    int j = lane;
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        const double a = sm[j], b = sm[j + 32], c = sm[j + 64], d = sm[j + 96];
        acc = fma(a, 1.0000001, acc); acc = fma(b, 1.0000001, acc); acc = fma(c, 1.0000001, acc); acc = fma(d, 1.0000001, acc);
        j = (__double2int_rd(acc) & 0) + ((j + 128) & 2047);
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[1] = (t1 - t0) / n;
    // 2c. TMEM load issued together with a dependent-index LDS gather (the shape of a solve batch): idx (LDS.u16 x4) -> x (LDS.64 x4), L from TMEM
    const unsigned short* ix = reinterpret_cast<const unsigned short*>(raw + 65536);
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        uint32_t r[8];
        tmem_ld8_nowait(my + col, r);
        const int i0 = ix[j] & 2047, i1 = ix[j + 32] & 2047, i2 = ix[j + 64] & 2047, i3 = ix[j + 96] & 2047;
        const double x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const double l0 = __longlong_as_double((long long)(((unsigned long long)r[1] << 32) | r[0])), l1 = __longlong_as_double((long long)(((unsigned long long)r[3] << 32) | r[2]));
        const double l2 = __longlong_as_double((long long)(((unsigned long long)r[5] << 32) | r[4])), l3 = __longlong_as_double((long long)(((unsigned long long)r[7] << 32) | r[6]));
        acc = fma(l0, x0, acc); acc = fma(l1, x1, acc); acc = fma(l2, x2, acc); acc = fma(l3, x3, acc);
        col = (uint32_t)(__double2int_rd(acc) & 0) + ((col + 8) & 127);
        j = (j + 128) & 2047;
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[2] = (t1 - t0) / n;
    // 2d. the same batch entirely from shared memory (what the kernel does today)
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        const int i0 = ix[j] & 2047, i1 = ix[j + 32] & 2047, i2 = ix[j + 64] & 2047, i3 = ix[j + 96] & 2047;
        const double l0 = sm[2048 + j], l1 = sm[2048 + j + 32], l2 = sm[2048 + j + 64], l3 = sm[2048 + j + 96];
        const double x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
        acc = fma(l0, x0, acc); acc = fma(l1, x1, acc); acc = fma(l2, x2, acc); acc = fma(l3, x3, acc);
        j = (__double2int_rd(acc) & 0) + ((j + 128) & 1023);
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[3] = (t1 - t0) / n;
    // 3. throughput: independent loads, four in flight per warp
    __syncthreads();
    t0 = clock64();
    double s0 = 0, s1 = 0;
    for (int i = 0; i < n; i++) {
        uint32_t r[8], q[8];
        tmem_ld8_nowait(my + ((8 * i) & 127), r);
        tmem_ld8_nowait(my + ((8 * i + 64) & 127), q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        s0 += __longlong_as_double((long long)(((unsigned long long)r[1] << 32) | r[0])); s1 += __longlong_as_double((long long)(((unsigned long long)q[7] << 32) | q[6]));
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[4] = (t1 - t0);      // cycles for n x 2 x 8 warps x 1 KB
    acc += s0 + s1;
    if (acc == 123.456) sink[1] = acc + j + col;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
}

int main() {
    long long* out; double* sink; int* bad;
    cudaMallocManaged(&out, 64 * 8); cudaMallocManaged(&sink, 64); cudaMallocManaged(&bad, 4);
    sink[0] = 0.5; *bad = 0;
    const int smem = 110 * 1024, n = 2000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 256, smem);
    k<<<2 * 148, 256, smem>>>(out, sink, n, bad);
    cudaError_t e = cudaDeviceSynchronize();
    printf("tmem ubench: %s; CTAs/SM by occupancy API %d; round-trip mismatches %d\n", cudaGetErrorString(e), nb, *bad);
    printf("  dependent batch of 4 doubles + 4 DFMA:   TMEM (LDTM.x8 + wait) %lld cycles,  shared (4 x LDS.64) %lld cycles\n", out[0], out[1]);
    printf("  solve-shaped batch (idx -> x gather + L): L from TMEM %lld cycles,  all shared %lld cycles\n", out[2], out[3]);
    printf("  streaming: %d x 2 LDTM.x8 per warp, 8 warps, 2 CTAs/SM: %lld cycles => %.1f B/cycle/CTA\n", n, out[4], (double)n * 2 * 8 * 1024 / (double)out[4]);
    return e != cudaSuccess || *bad;
}
