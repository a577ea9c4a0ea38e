// tmem_cp.cu -- can the key of the multiply-accumulate phase reach the threads through tensor memory instead of ld.shared?
// tcgen05.cp.32x128b.warpx4 copies 32 rows x 16 bytes of shared memory into 4 tensor-memory columns of ALL FOUR lane quarters
// (the same data for the four ciphertexts of a CTA); tcgen05.ld.32x32b then gives every thread its lane's columns.
// (1) layout: shared memory holds the word index; 8 copies of 512 contiguous bytes (descriptor: no swizzle, 8-row groups 128
//     bytes apart) go to columns 4e..4e+3; every warp prints what its lanes read.
// (2) cost: cycles per 24 KB (one key row of the N = 512 set: 48 copies) for the copy stream alone, for the shared-memory loads
//     it would replace (12 warps x 16 LDS.128 per thread per row ... here per 6 polynomials: 12 warps x 32 lanes x 2 x 8 x 16 B),
//     for the tensor-memory loads that replace them, and for the combinations.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t cp_desc(uint32_t addr) {   // K-major, no swizzle: start >> 4, LBO 0, SBO 128 >> 4, version 1
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void cp_32x128b_warpx4(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    for (long long i = 0; !ok && i < (1ll << 26); i++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#define LD32(v, addr)                                                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, " \
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),  \
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),    \
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),    \
                   "=r"(v[31])                                                                                                                       \
                 : "r"(addr)                                                                                                                         \
                 : "memory")

__global__ void __launch_bounds__(128, 1) map_kernel(uint32_t *out) {
    __shared__ __align__(1024) uint32_t src[1024];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 128) src[i] = i;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of src -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s;
    if (threadIdx.x == 0) {
        for (uint32_t e = 0; e < 8; e++) cp_32x128b_warpx4(base + 4 * e, cp_desc(smem_u32(src) + 512 * e));
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    LD32(v, base + ((warp * 32u) << 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out[(warp * 32 + lane) * 32 + i] = v[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

// MODE bit 0: thread 0 streams copies (48 per "row", commit + wait per row); bit 1: all 12 warps load 2 x 32 columns per row from
// tensor memory (what a thread needs of a row: 2 limbs x 8 points x 16 B); bit 2: all 12 warps load the same bytes with LDS.128
template <int MODE>
__global__ void __launch_bounds__(384, 1) bw_kernel(uint32_t *sink, int rows, long long *cycles) {
    extern __shared__ __align__(1024) uint4 sm[];   // 24 KB key row + padding
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1536; i += 384) sm[i] = make_uint4(i, i, i, i);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s, sub = warp >> 2;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rows; r++) {
        if ((MODE & 1) && threadIdx.x == 0) {
            for (uint32_t q = 0; q < 48; q++) cp_32x128b_warpx4(base + 4 * q, cp_desc(smem_u32(sm) + 512 * q));
            commit(&bar);
            mbar_wait(&bar, r & 1);
        }
        if (MODE & 2) {
            uint32_t v[32];
            for (uint32_t l = 0; l < 2; l++) {
                LD32(v, base + (((warp & 3u) * 32u) << 16) + 64 * sub + 32 * l);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; i++) acc ^= v[i];
            }
        }
        if (MODE & 4) {
            for (uint32_t l = 0; l < 2; l++)
#pragma unroll
                for (uint32_t e = 0; e < 8; e++) {
                    uint4 x;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(smem_u32(sm) + ((sub * 2 + l) * 256 + e * 32 + lane) * 16));
                    acc ^= x.x ^ x.y ^ x.z ^ x.w;
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 384 + threadIdx.x] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
}

template <int MODE>
double run(uint32_t *sink, long long *cyc, int rows) {
    cudaFuncSetAttribute(bw_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    bw_kernel<MODE><<<148, 384, 32768>>>(sink, rows, cyc);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"bw_kernel<%d>: %s\"}\n", MODE, cudaGetErrorString(cudaGetLastError())); return -1; }
    bw_kernel<MODE><<<148, 384, 32768>>>(sink, rows, cyc);
    cudaDeviceSynchronize();
    double s = 0;
    for (int i = 0; i < 148; i++) s += (double)cyc[i];
    return s / 148 / rows;
}

int main() {
    uint32_t *out, *sink;
    long long *cyc;
    cudaMallocManaged(&out, 128 * 32 * 4);
    cudaMallocManaged(&sink, 148 * 384 * 4);
    cudaMallocManaged(&cyc, 148 * 8);
    map_kernel<<<1, 128>>>(out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"map_kernel: %s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    int bad = 0;
    for (int w = 0; w < 4; w++)
        for (int l = 0; l < 32; l++)
            for (int c = 0; c < 32; c++) bad += out[(w * 32 + l) * 32 + c] != (uint32_t)((c / 4) * 128 + l * 4 + (c % 4));
    printf("layout: lane l of every quarter, column 4e + k  <-  shared word e * 128 + 4 l + k : %s (%d mismatches)\n", bad ? "NO" : "yes", bad);
    if (bad)
        for (int w = 0; w < 4; w += 3)
            for (int l = 0; l < 4; l++) {
                printf("  q%d l%d:", w, l);
                for (int c = 0; c < 12; c++) printf(" %4u", out[(w * 32 + l) * 32 + c]);
                printf("\n");
            }
    const int rows = 2000;
    const double cp = run<1>(sink, cyc, rows), ldtm = run<2>(sink, cyc, rows), lds = run<4>(sink, cyc, rows), cp_ldtm = run<3>(sink, cyc, rows), cp_lds = run<5>(sink, cyc, rows),
                 ldtm_lds = run<6>(sink, cyc, rows), all = run<7>(sink, cyc, rows);
    printf("{\"unit\": \"cycles per key row (24 KB copied once / 12 warps x 256 B per thread loaded)\", \"cp\": %.1f, \"ldtm\": %.1f, \"lds\": %.1f, \"cp+ldtm\": %.1f, \"cp+lds\": %.1f, "
           "\"ldtm+lds\": %.1f, \"cp+ldtm+lds\": %.1f}\n",
           cp, ldtm, lds, cp_ldtm, cp_lds, ldtm_lds, all);
    return 0;
}
