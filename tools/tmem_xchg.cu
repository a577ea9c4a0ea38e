// tmem_xchg.cu -- can tensor memory serve as a second exchange path for the FFT passes of the blind rotation?
// (1) What permutation does "store 32x32b, load 16x256b" realise inside a warp?  Every thread stores 8 words (lane l, columns 0..7,
//     value = l * 16 + column) with tcgen05.st.32x32b.x8 and reads 4 + 4 words back with tcgen05.ld.16x256b.x1 at lane offsets 0 and 16:
//     the program prints which (lane, column) each thread's registers received.
// (2) Does TMEM traffic run beside shared-memory traffic or on the same pipe?  12 warps per SM; loops of
//       mode 0: 8 STS.128 + 8 LDS.128 per thread (a shared-memory exchange of 8 complex doubles),
//       mode 1: tcgen05.st 32 columns + tcgen05.ld 32 columns (the same 128 bytes per thread through tensor memory),
//       mode 2: both per iteration.
//     Prints cycles per iteration per SM (timed over all warps of the CTA).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1) map_kernel(uint32_t *out) {
    __shared__ uint32_t tmem_base_s;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s, lane = threadIdx.x;
    uint32_t v[8];
    for (int c = 0; c < 8; c++) v[c] = lane * 16 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base + (16u << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) out[lane * 8 + i] = r[i];
    // the reverse direction: store with 16x256b, load with 32x32b
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(base), "r"(lane * 16 + 0), "r"(lane * 16 + 1), "r"(lane * 16 + 2), "r"(lane * 16 + 3) : "memory");
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (16u << 16)), "r"(lane * 16 + 4), "r"(lane * 16 + 5), "r"(lane * 16 + 6), "r"(lane * 16 + 7) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) out[256 + lane * 8 + i] = r[i];
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(384, 1) bw_kernel(uint32_t *sink, int iters, long long *cycles) {
    extern __shared__ uint4 sm[];
    __shared__ uint32_t tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w owns TMEM lanes 32 (w % 4) .. +31 and columns 32 (w / 4) .. +31
    const uint32_t taddr = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 32u;
    uint4 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = make_uint4(threadIdx.x, i, 2 * i, 3);
    const uint32_t sbase = smem_u32(sm + warp * 8 * 36 + lane + (lane >> 3));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)i * 36u * 16u), "r"(w[i].x), "r"(w[i].y), "r"(w[i].z), "r"(w[i].w) : "memory");
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint4 r;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(sbase + (uint32_t)((i + it) & 7) * 36u * 16u) : "memory");
                w[i].x += r.y; w[i].y ^= r.z; w[i].z += r.w; w[i].w ^= r.x;
            }
        }
        if (MODE == 1 || MODE == 2) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
                "%24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
                "r"(w[0].x), "r"(w[0].y), "r"(w[0].z), "r"(w[0].w), "r"(w[1].x), "r"(w[1].y), "r"(w[1].z), "r"(w[1].w), "r"(w[2].x), "r"(w[2].y), "r"(w[2].z), "r"(w[2].w),
                "r"(w[3].x), "r"(w[3].y), "r"(w[3].z), "r"(w[3].w), "r"(w[4].x), "r"(w[4].y), "r"(w[4].z), "r"(w[4].w), "r"(w[5].x), "r"(w[5].y), "r"(w[5].z), "r"(w[5].w),
                "r"(w[6].x), "r"(w[6].y), "r"(w[6].z), "r"(w[6].w), "r"(w[7].x), "r"(w[7].y), "r"(w[7].z), "r"(w[7].w)
                : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            // read back with the transposing shape: 2 lane halves x 4 column groups of 8 x 4 registers = 32 words
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    uint4 &d = w[h * 4 + g];
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(taddr + ((uint32_t)h * 16u << 16) + (uint32_t)g * 8u) : "memory");
                }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; i++) w[i].x += (uint32_t)it;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += w[i].x + w[i].y + w[i].z + w[i].w;
    if (s == 0x12345678u) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base_s) : "memory");
}

int main() {
    uint32_t *out; long long *cyc; uint32_t *sink;
    cudaMallocManaged(&out, 512 * 4); cudaMallocManaged(&cyc, 3 * sizeof(long long)); cudaMalloc(&sink, 4);
    map_kernel<<<1, 32>>>(out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"map_kernel: %s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    printf("store 32x32b.x8 (value = lane*16 + column), load 16x256b.x1 at lane offsets 0 and 16: thread t received (lane,column) in r0..r7:\n");
    for (int t = 0; t < 32; t++) {
        printf("  t%2d:", t);
        for (int i = 0; i < 8; i++) printf(" (%2u,%u)", out[t * 8 + i] / 16, out[t * 8 + i] % 16);
        printf("\n");
    }
    printf("store 16x256b.x1 at lane offsets 0 / 16 (thread t stores t*16 + {0..3} / {4..7}), load 32x32b.x8: lane l, columns 0..7 hold (thread,register):\n");
    for (int t = 0; t < 32; t++) {
        printf("  l%2d:", t);
        for (int i = 0; i < 8; i++) printf(" (%2u,%u)", out[256 + t * 8 + i] / 16, out[256 + t * 8 + i] % 16);
        printf("\n");
    }
    const int iters = 20000, smem = 12 * 8 * 36 * 16 + 1024;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(bw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) {
        bw_kernel<0><<<sms, 384, smem>>>(sink, iters, cyc);
        bw_kernel<1><<<sms, 384, smem>>>(sink, iters, cyc);
        bw_kernel<2><<<sms, 384, smem>>>(sink, iters, cyc);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"bw_kernel: %s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    const double c0 = (double)cyc[0] / iters, c1 = (double)cyc[1] / iters, c2 = (double)cyc[2] / iters;
    printf("{\"smem_exchange_cycles\": %.1f, \"tmem_exchange_cycles\": %.1f, \"both_cycles\": %.1f, \"sum\": %.1f, \"max\": %.1f, "
           "\"note\": \"12 warps per SM, per iteration and warp: 128 bytes per thread stored and loaded (8 STS.128 + 8 LDS.128 / tcgen05.st x32 + 8 tcgen05.ld 16x256b.x1)\"}\n",
           c0, c1, c2, c0 + c1, c0 > c1 ? c0 : c1);
    return 0;
}
