// fp64_peak.cu -- stand-alone B200 microbenchmark: FP64 pipe (DFMA/DADD/DMUL) rate and latency, alone and
// co-issued with IMAD, plus a register-resident complex-butterfly stream.  Decides whether an exact
// FP64-FFT external product (limb-split key) can beat the 2-prime integer NTT on this chip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int KIND>
__global__ void __launch_bounds__(256) k_rate(double *sink, double a, double b, uint32_t ia, uint32_t ib, int iters) {
    double x[8];
    uint32_t y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = a + threadIdx.x + i; y[i] = ia + threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == 0) x[i] = fma(x[i], a, b);                    // DFMA
                if (KIND == 1) x[i] = x[i] + b;                           // DADD
                if (KIND == 2) x[i] = x[i] * a;                           // DMUL
                if (KIND == 3) y[i] = y[i] * ia + ib;                     // IMAD
                if (KIND == 4) { x[i] = fma(x[i], a, b); y[i] = y[i] * ia + ib; }   // DFMA + IMAD co-issue
                if (KIND == 5) { x[i] = fma(x[i], a, b); y[i] = __umulhi(y[i], ia) + ib; }  // DFMA + IMAD.HI
                if (KIND == 6) { x[i] = fma(x[i], a, b); y[i] = (y[i] + ib) ^ ia; }  // DFMA + ALU
            }
        }
    }
    double s = 0; uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i]; t += y[i]; }
    if (s == 1.2345 || t == 77) sink[0] = s + t;
}

// DFMA with three distinct register operands (as in the FFT butterflies / multiply-accumulate): does operand
// collection sustain the full rate?  KIND 0: x = fma(x, y, z) (3 register pairs); 1: x = fma(x, y, c) (2 + constant);
// 2: like 0 with 16 independent chains per thread
template <int KIND>
__global__ void __launch_bounds__(256) k_rate3(double *sink, double a, double b, int iters) {
    constexpr int NCH = KIND == 2 ? 16 : 8;
    double x[NCH], y[8], z[8];
#pragma unroll
    for (int i = 0; i < NCH; i++) x[i] = a + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; i++) { y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (threadIdx.x * 3 + i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < (KIND == 2 ? 2 : 4); u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (KIND == 1) x[i] = fma(x[i], y[(i + u) & 7], b);
                else x[i] = fma(x[i], y[(i + u) & 7], z[(i + 2 * u + 1) & 7]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += x[i];
    if (s == 1.2345) sink[0] = s;
}

// dependent chain latency: one warp per SM, one chain
template <int KIND>
__global__ void k_lat(double *sink, double a, double b, int iters, long long *cycles) {
    double x = a + threadIdx.x;
    uint32_t y = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (KIND == 0) x = fma(x, a, b);
            if (KIND == 1) x = x + b;
            if (KIND == 2) y = y * 0x9E3779B1u + 12345u;
            if (KIND == 3) y = __umulhi(y, 0x9E3779B1u) + 12345u;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (x == 1.2345 || y == 77) sink[0] = x + y;
}

// register-resident radix-2 complex butterfly stream, 8 complex points per thread, 3 stages per round
// (X' = X + w*Y, Y' = X - w*Y written with explicit FMAs: 8 DP ops per butterfly)
__global__ void __launch_bounds__(128) k_bfly(double *sink, double wr, double wi, int iters) {
    double xr[8], xi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { xr[i] = 1e-3 * (threadIdx.x + i); xi[i] = 1e-3 * (threadIdx.x - i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 3; s++) {
            const int bit = 4 >> s;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                if (e & bit) continue;
                const double yr = xr[e + bit], yi = xi[e + bit];
                const double tr = fma(-yi, wi, yr * wr), ti = fma(yi, wr, yr * wi);
                xr[e + bit] = xr[e] - tr; xi[e + bit] = xi[e] - ti;
                xr[e] = xr[e] + tr;       xi[e] = xi[e] + ti;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += xr[i] + xi[i];
    if (s == 1.2345) sink[0] = s;
}

// One FFT-pass-like loop per warp: 8 x LDS.128, 36 forward butterflies (twiddles in registers: 4 of 6 DFMAs have three
// register operands), 8 x STS.128, conflict-free addresses, NO barriers; `threads` per CTA, one CTA per SM.  Reports
// cycles per iteration per SM next to what the FP64 pipe alone and the shared-memory pipe alone would need: do the two
// overlap across warps?
template <int MODE>   // 0: loads + math + stores, 1: math only, 2: loads + stores only, 3: like 0 + a 64-thread named barrier per pass,
                      // 4: like 3 + per-pass twiddle loads from global memory (4 x LDG.128 per thread)
__global__ void k_pass(double *sink, int iters, long long *cycles, const double2 *gtw) {
    extern __shared__ double2 sm[];
    double2 x[8], w[4];
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = make_double2(1e-3 * (t + i), 1e-3 * (t - i));
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = make_double2(0.7 + 1e-3 * (t + i), 0.7 - 1e-3 * i);
    for (int i = 0; i < 8; i++) sm[i * blockDim.x + t] = x[i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE >= 3) asm volatile("bar.sync %0, 64;" ::"r"(1 + t / 64) : "memory");
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 4; i++) w[i] = __ldg(gtw + i * blockDim.x + t);
        }
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = sm[i * blockDim.x + (t ^ 1)];   // a neighbour's data: cannot be forwarded from registers
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i].x += 1.0;
        }
        if (MODE != 2) {
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const int bit = 4 >> s;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    if (e & bit) continue;
                    const double2 ww = w[(e + s) & 3];
                    double2 &X = x[e], &Y = x[e + bit];
                    const double xr = __fma_rn(-Y.y, ww.y, __fma_rn(Y.x, ww.x, X.x));
                    const double xi = __fma_rn(Y.y, ww.x, __fma_rn(Y.x, ww.y, X.y));
                    Y.x = __fma_rn(2.0, X.x, -xr); Y.y = __fma_rn(2.0, X.y, -xi);
                    X.x = xr; X.y = xi;
                }
            }
        }
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) sm[i * blockDim.x + t] = x[i];
        }
    }
    const long long t1 = clock64();
    if (t == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    if (s == 1.2345) sink[0] = s;
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d", pr.name, pr.multiProcessorCount, clk_khz);
    double *sink; long long *cyc;
    CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&cyc, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = pr.multiProcessorCount * 8, iters = 4096;
    const char *names[7] = {"dfma", "dadd", "dmul", "imad", "dfma+imad", "dfma+imadhi", "dfma+alu"};
    for (int kind = 0; kind < 7; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            switch (kind) {
                case 0: k_rate<0><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 1: k_rate<1><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 2: k_rate<2><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 3: k_rate<3><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 4: k_rate<4><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 5: k_rate<5><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
                case 6: k_rate<6><<<blocks, 256>>>(sink, 1.0000001, 1e-9, 3, 5, iters); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double ops = (double)blocks * 256 * iters * 32.0;  // per pipe: 32 ops per thread per iteration
        printf(", \"%s_Tops\": %.3f", names[kind], ops / (best * 1e-3) / 1e12);
    }
    const char *n3[3] = {"dfma_3reg", "dfma_2reg_const", "dfma_3reg_16chains"};
    for (int kind = 0; kind < 3; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (kind == 0) k_rate3<0><<<blocks, 256>>>(sink, 1.0000001, 1e-9, iters);
            else if (kind == 1) k_rate3<1><<<blocks, 256>>>(sink, 1.0000001, 1e-9, iters);
            else k_rate3<2><<<blocks, 256>>>(sink, 1.0000001, 1e-9, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double ops = (double)blocks * 256 * iters * 32.0;
        printf(", \"%s_Tops\": %.3f", n3[kind], ops / (best * 1e-3) / 1e12);
    }
    for (int w = 1; w <= 4; w++) {   // rate vs resident warps per scheduler (3-register DFMA, 8 chains): 128-thread CTAs, w per SM
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            k_rate3<0><<<pr.multiProcessorCount * w, 128>>>(sink, 1.0000001, 1e-9, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf(", \"dfma_3reg_%dwarps_per_sched_Tops\": %.3f", w, (double)pr.multiProcessorCount * w * 128 * iters * 32.0 / (best * 1e-3) / 1e12);
    }
    {
        float best = 1e30f;
        const int bi = 2048;
        const int bb = pr.multiProcessorCount * 4;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k_bfly<<<bb, 128>>>(sink, 0.7071, 0.7071, bi);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double bf = (double)bb * 128 * bi * 12.0;
        printf(", \"cplx_bfly_T_per_s\": %.3f, \"cplx_bfly_warps_per_sm\": 16", bf / (best * 1e-3) / 1e12);
    }
    const char *ln[4] = {"dfma_lat", "dadd_lat", "imad_lat", "imadhi_lat"};
    for (int kind = 0; kind < 4; kind++) {
        const int li = 1024;
        for (int rep = 0; rep < 2; rep++) {
            switch (kind) {
                case 0: k_lat<0><<<1, 32>>>(sink, 1.0000001, 1e-9, li, cyc); break;
                case 1: k_lat<1><<<1, 32>>>(sink, 1.0000001, 1e-9, li, cyc); break;
                case 2: k_lat<2><<<1, 32>>>(sink, 1.0000001, 1e-9, li, cyc); break;
                case 3: k_lat<3><<<1, 32>>>(sink, 1.0000001, 1e-9, li, cyc); break;
            }
            CK(cudaDeviceSynchronize());
        }
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf(", \"%s_cycles\": %.2f", ln[kind], (double)c / (li * 16.0));
    }
    for (int threads = 128; threads <= 512; threads += 128) {
        const int it3 = 2000;
        const size_t smem = (size_t)threads * 8 * 16;
        long long c[5];
        double2 *gtw; CK(cudaMalloc(&gtw, 512 * 4 * 16)); CK(cudaMemset(gtw, 0, 512 * 4 * 16));
        for (int mode = 0; mode < 5; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) { cudaFuncSetAttribute(k_pass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<0><<<pr.multiProcessorCount, threads, smem>>>(sink, it3, cyc, gtw); }
                if (mode == 1) { cudaFuncSetAttribute(k_pass<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<1><<<pr.multiProcessorCount, threads, smem>>>(sink, it3, cyc, gtw); }
                if (mode == 2) { cudaFuncSetAttribute(k_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<2><<<pr.multiProcessorCount, threads, smem>>>(sink, it3, cyc, gtw); }
                if (mode == 3) { cudaFuncSetAttribute(k_pass<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<3><<<pr.multiProcessorCount, threads, smem>>>(sink, it3, cyc, gtw); }
                if (mode == 4) { cudaFuncSetAttribute(k_pass<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<4><<<pr.multiProcessorCount, threads, smem>>>(sink, it3, cyc, gtw); }
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(&c[mode], cyc, 8, cudaMemcpyDeviceToHost));
        }
        printf(", \"pass_loop_%dthr_cycles\": {\"both\": %.0f, \"math_only\": %.0f, \"lds_sts_only\": %.0f, \"both_pair_barrier\": %.0f, \"both_pair_barrier_ldg_twiddles\": %.0f}", threads,
               (double)c[0] / it3, (double)c[1] / it3, (double)c[2] / it3, (double)c[3] / it3, (double)c[4] / it3);
        cudaFree(gtw);
    }
    printf("}\n");
    return 0;
}
