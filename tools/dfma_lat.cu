// dfma_lat.cu -- dependent-issue latency of DFMA / DADD / DMUL on B200 (one warp per SM sub-partition, nothing else running):
// a chain of N dependent operations, cycles per operation.  Context: a lone ciphertext's CMUX step is a chain of ~27 dependent FP64
// operations per transform (3 per butterfly stage, 9 stages) -- profiles/r02_latency_phases.txt.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double *out, double a, double b, long long *cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int j = 0; j < 64; j++) x = KIND == 0 ? __fma_rn(x, b, a) : KIND == 1 ? __dadd_rn(x, a) : __dmul_rn(x, b);
    }
    long long t1 = clock64();
    if (x == 1.2345) out[0] = x;
    if (threadIdx.x == 0) cyc[KIND] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 3 * sizeof(long long));
    for (int rep = 0; rep < 2; rep++) { k<0><<<1, 32>>>(out, 1.0000001, 0.9999999, cyc); k<1><<<1, 32>>>(out, 1e-9, 1.0, cyc); k<2><<<1, 32>>>(out, 1.0, 1.0000001, cyc); cudaDeviceSynchronize(); }
    printf("{\"dfma_dependent_cycles\": %.1f, \"dadd_dependent_cycles\": %.1f, \"dmul_dependent_cycles\": %.1f}\n", cyc[0] / 4096.0, cyc[1] / 4096.0, cyc[2] / 4096.0);
    return 0;
}
