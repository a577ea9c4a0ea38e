// shfl_peak.cu -- does the warp shuffle share the shared-memory data pipe on B200?
// Three loops, 12 warps per SM (the blind-rotation kernel's occupancy), one CTA per SM:
//   mode 0: SHFL.BFLY only;  mode 1: LDS.128 + STS.128 only;  mode 2: both interleaved (same counts per iteration).
// Prints cycles per iteration per SM for each; if mode 2 ~ mode 0 + mode 1 the two share a pipe, if ~ max they do not.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(uint32_t *sink, int iters, long long *cycles) {
    extern __shared__ uint4 sm[];
    uint32_t v[8];
    uint4 w[4];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 7 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = make_uint4(threadIdx.x, i, 2, 3);
    uint4 *mine = sm + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __shfl_xor_sync(0xFFFFFFFFu, v[i], 1 + (i & 15)) + 1;
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) mine[i * 384] = w[i];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 r = mine[i * 384 + ((it + i) & 1)];   // neighbour's slot half of the time: still conflict-free
                w[i].x += r.y; w[i].y ^= r.z; w[i].z += r.w; w[i].w ^= r.x;
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += w[i].x + w[i].y + w[i].z + w[i].w;
    if (s == 0x12345678u) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = t1 - t0;
}

int main() {
    uint32_t *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 3 * sizeof(long long));
    const int iters = 20000, smem = 4 * 384 * 16 + 64;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    k<0><<<sms, 384, smem>>>(sink, iters, cyc); k<1><<<sms, 384, smem>>>(sink, iters, cyc); k<2><<<sms, 384, smem>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    k<0><<<sms, 384, smem>>>(sink, iters, cyc); k<1><<<sms, 384, smem>>>(sink, iters, cyc); k<2><<<sms, 384, smem>>>(sink, iters, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    // per iteration per SM: 12 warps x 8 SHFL (32 lanes x 4 B = 128 B each); 12 warps x (4 STS.128 + 4 LDS.128) x 512 B
    double c0 = (double)cyc[0] / iters, c1 = (double)cyc[1] / iters, c2 = (double)cyc[2] / iters;
    printf("{\"shfl_only_cycles\": %.1f, \"shfl_warp_instr_per_clk_per_sm\": %.3f, \"smem_only_cycles\": %.1f, \"smem_bytes_per_clk_per_sm\": %.1f, "
           "\"both_cycles\": %.1f, \"sum\": %.1f, \"max\": %.1f}\n",
           c0, 96.0 / c0, c1, 12.0 * 8 * 512 / c1, c2, c0 + c1, c0 > c1 ? c0 : c1);
    return 0;
}
