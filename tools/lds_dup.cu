// lds_dup.cu -- what do 128-bit shared-memory accesses cost on B200, by address pattern?
// Questions behind it: (1) if lanes of different quarter-warps read the SAME address (two ciphertexts paired in one warp
// multiply different digits with the same bootstrapping-key element), are the duplicates merged?  (2) does the padded
// exchange layout phys(j) = j + (j >> 3) of the FFT kernel cost more than a contiguous 512-byte access?  (3) do stores cost
// what loads cost?
// 12 warps per SM, one CTA per SM; every warp works on its own 16 KB region; patterns (16-byte element read/written by a lane):
//   0: lane                      contiguous 512 B
//   1: lane & 15                 lanes l and l+16 the same element (256 B unique)
//   2: lane & 7                  four-fold duplication (128 B unique)
//   3: 0                         broadcast
//   4: lane + (lane >> 3)        the kernel's padded layout: four runs of 8 elements, 16 B gaps
//   5: lane * 9                  layout C of the kernel: stride 9 elements (conflict-free per quarter-warp)
//   6: (lane & 15) + ((lane & 15) >> 3)   padded AND duplicated across half-warps
// OP 0 = LDS.128, 1 = STS.128, 2 = alternating STS.128 / LDS.128 (no dependence between them), 3 = STS.64, 4 = LDS.64
// (8-byte elements at 8-byte stride: lane -> element lane, 256 B per instruction).
// Prints cycles per warp-level instruction per SM, timed over ALL warps of the CTA.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT, int OP>
__global__ void __launch_bounds__(384, 1) k(uint32_t *sink, int iters, long long *cycles, int slot) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 12 * 1024; i += 384) sm[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    const int idx = PAT == 0 ? lane : PAT == 1 ? (lane & 15) : PAT == 2 ? (lane & 7) : PAT == 3 ? 0 : PAT == 4 ? lane + (lane >> 3)
                  : PAT == 5 ? lane * 9 : (lane & 15) + ((lane & 15) >> 3);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + warp * 1024 + idx);
    uint4 a = make_uint4(lane, 1, 2, 3);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t ad = base + (uint32_t)(i * 64 + ((it & 1) << 5)) * 16u;   // 8 rows of 64 elements, two column halves
            if (OP == 3 || OP == 4) {
                const uint32_t ad8 = (uint32_t)__cvta_generic_to_shared(sm + warp * 1024) + (uint32_t)(lane + i * 128 + ((it & 1) << 6)) * 8u;
                if (OP == 4) {
                    uint2 r;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(ad8));
                    a.x += r.x; a.y ^= r.y;
                } else {
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ad8), "r"(a.x), "r"(a.y) : "memory");
                }
            } else if (OP == 0 || (OP == 2 && (i & 1))) {
                uint4 r;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(ad));
                a.x += r.x; a.y ^= r.y; a.z += r.z; a.w ^= r.w;
            } else {
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
            }
        }
    }
    __syncthreads();   // time ALL warps of the CTA
    long long t1 = clock64();
    if (a.x + a.y + a.z + a.w == 0x12345678u) sink[0] = a.x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[slot] = t1 - t0;
}

int main() {
    uint32_t *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 32 * sizeof(long long));
    const int iters = 20000, smem = 12 * 1024 * 16;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int slot = 0;
    const char *names[32];
#define RUN(P, O, NAME) { cudaFuncSetAttribute(k<P, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<P, O><<<sms, 384, smem>>>(sink, iters, cyc, slot); names[slot++] = NAME; }
    for (int rep = 0; rep < 2; rep++) {
        slot = 0;
        RUN(0, 0, "lds128_contiguous") RUN(1, 0, "lds128_dup_halfwarps") RUN(2, 0, "lds128_dup_4x") RUN(3, 0, "lds128_broadcast")
        RUN(4, 0, "lds128_padded") RUN(5, 0, "lds128_stride9") RUN(6, 0, "lds128_padded_dup_halfwarps")
        RUN(0, 1, "sts128_contiguous") RUN(4, 1, "sts128_padded") RUN(5, 1, "sts128_stride9")
        RUN(0, 2, "mix_contiguous") RUN(4, 2, "mix_padded") RUN(0, 3, "sts64_contiguous") RUN(0, 4, "lds64_contiguous")
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    printf("{\"unit\": \"cycles per warp-level instruction per SM (12 warps)\"");
    for (int i = 0; i < slot; i++) printf(", \"%s\": %.2f", names[i], (double)cyc[i] / iters / 96.0);
    printf("}\n");
    return 0;
}
