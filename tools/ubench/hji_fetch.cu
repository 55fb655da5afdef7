// Why does the HJI gather move 1.6x its algorithmic bytes through DRAM?  Each query reads 64 PAIRS of adjacent 32-byte records (the two dim-1
// corners of a cell), i.e. 64 contiguous 64-byte segments at 32-byte granularity.  This benchmark replays that access pattern on a synthetic
// 320 MB table (>> L2) in four variants and reports the time per query; run it under `ncu --metrics dram__bytes_read.sum` for the DRAM bytes:
//   A  pairs at arbitrary 32-byte offsets (the engine's layout: half of them straddle a 64-byte boundary)
//   B  the same number of 64-byte segments, every one 64-byte ALIGNED (what a layout that stores each dim-1 cell as one aligned block would read)
//   C  A, with the queries visited in table order (sorted by cell): neighbouring queries share sectors through L2
//   D  single 32-byte records at random (the sector-granularity floor)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hji_fetch hji_fetch.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void __launch_bounds__(128) k(const float4* __restrict__ tab, size_t nrec, int M, float* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= M) return;
    // a query's 64 pairs sit at pseudo-random "rows" of the table (the 6 slow dimensions), the same dim-1 offset in each
    const uint32_t h0 = MODE == 2 ? (uint32_t)(((unsigned long long)qi * (nrec / 16)) / M) * 16u : mix(qi * 2654435761u + 17u);
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < 64; c++) {
        size_t r;
        if (MODE == 2) r = ((size_t)h0 + (size_t)c * 13u * ((c & 1) ? 169u : 1u) * ((c & 2) ? 9u : 1u)) % (nrec - 2);      // a cell's corners: strides of the slow dims
        else r = (size_t)(mix(h0 + 0x9e3779b9u * c) % (uint32_t)(nrec - 2));
        if (MODE == 1) r &= ~(size_t)1;                      // 64-byte aligned pair
        const float4 a0 = __ldg(tab + 2 * r), a1 = __ldg(tab + 2 * r + 1);
        acc += a0.x + a1.w;
        if (MODE != 3) { const float4 b0 = __ldg(tab + 2 * r + 2), b1 = __ldg(tab + 2 * r + 3); acc += b0.y + b1.z; }
    }
    out[qi] = acc;
}
int main() {
    const size_t nrec = 10 * 1000 * 1000;            // 32-byte records: 320 MB
    const int M = 1 << 22;
    float4* tab; float* out;
    cudaMalloc(&tab, nrec * 32 + 64); cudaMalloc(&out, (size_t)M * 4);
    cudaMemset(tab, 0, nrec * 32 + 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"A unaligned pairs (engine layout)", "B 64-byte aligned pairs", "C unaligned pairs, queries in table order", "D single records"};
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<(M + 127) / 128, 128>>>(tab, nrec, M, out);
            if (mode == 1) k<1><<<(M + 127) / 128, 128>>>(tab, nrec, M, out);
            if (mode == 2) k<2><<<(M + 127) / 128, 128>>>(tab, nrec, M, out);
            if (mode == 3) k<3><<<(M + 127) / 128, 128>>>(tab, nrec, M, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double bytes = (double)M * 64 * (mode == 3 ? 32 : 64);
        printf("%-44s %8.3f ms  %7.1f M queries/s  %7.1f GB/s algorithmic\n", names[mode], best, M / best / 1e3, bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
