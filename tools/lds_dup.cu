// lds_dup.cu -- what does a 128-bit shared-memory load cost on B200 when lanes of different quarter-warps read the SAME
// address?  (Question behind pairing two ciphertexts in one warp: lanes l and l+16 multiply different ciphertexts'
// digits with the same bootstrapping-key element; if the duplicate addresses are merged into one wavefront, the key
// reads from the TMA ring cost half the shared-memory data-pipe cycles.)
// 12 warps per SM, one CTA per SM, LDS.128 only; patterns (element index read by a lane, 16-byte elements):
//   0: lane           -- 32 distinct elements, 512 B  (4 wavefronts expected)
//   1: lane & 15      -- lanes l and l+16 read the same element, 256 B unique
//   2: lane & 7       -- four-fold duplication, 128 B unique
//   3: 0              -- all lanes one element (broadcast)
//   4: (lane & 7) | ((lane & 16) >> 1)   -- duplicates INSIDE each quarter-warp pair arrangement: lanes {0..7} and {8..15} the same
//                        elements, {16..23} and {24..31} the next 8: 256 B unique with the duplicates adjacent
// Also the 64-bit variants (LDS.64) of patterns 0 and 1.  Prints cycles per warp-level load per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT, int WIDTH>
__global__ void __launch_bounds__(384, 1) k(uint32_t *sink, int iters, long long *cycles, int slot) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 12 * 8 * 64; i += 384) sm[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    int idx = PAT == 0 ? lane : PAT == 1 ? (lane & 15) : PAT == 2 ? (lane & 7) : PAT == 3 ? 0 : ((lane & 7) | ((lane & 16) >> 1));
    const uint4 *base = sm + warp * 8 * 64 + idx;
    uint4 a = make_uint4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            // asm volatile: the loads must stay in the loop (the addresses repeat, nothing is stored in between)
            if (WIDTH == 16) {
                uint4 r;
                const uint32_t ad = (uint32_t)__cvta_generic_to_shared(base + i * 64 + ((it & 1) << 5));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(ad));
                a.x += r.x; a.y ^= r.y; a.z += r.z; a.w ^= r.w;
            } else {
                uint2 r;
                const uint32_t ad = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<const uint2 *>(base) + i * 128 + ((it & 1) << 6));
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(ad));
                a.x += r.x; a.y ^= r.y;
            }
        }
    }
    long long t1 = clock64();
    if (a.x + a.y + a.z + a.w == 0x12345678u) sink[0] = a.x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[slot] = t1 - t0;
}

int main() {
    uint32_t *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 16 * sizeof(long long));
    const int iters = 20000, smem = 12 * 8 * 64 * 16;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define RUN(P, W, S) { cudaFuncSetAttribute(k<P, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<P, W><<<sms, 384, smem>>>(sink, iters, cyc, S); }
    for (int rep = 0; rep < 2; rep++) {
        RUN(0, 16, 0) RUN(1, 16, 1) RUN(2, 16, 2) RUN(3, 16, 3) RUN(4, 16, 4) RUN(0, 8, 5) RUN(1, 8, 6)
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    // per iteration per SM: 12 warps x 8 loads
    const char *names[7] = {"lds128_distinct", "lds128_dup_halfwarps", "lds128_dup_4x", "lds128_broadcast", "lds128_dup_adjacent_quarters", "lds64_distinct", "lds64_dup_halfwarps"};
    printf("{");
    for (int i = 0; i < 7; i++) printf("\"%s_cycles_per_warp_load\": %.2f%s", names[i], (double)cyc[i] / iters / 96.0, i < 6 ? ", " : "}\n");
    return 0;
}
