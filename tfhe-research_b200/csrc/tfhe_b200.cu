// tfhe_b200.cu -- C-ABI implementation (device side) of include/tfhe_b200.h: context, key upload and
// the batched PBS entry points.  No CPU fallback: every entry point here needs a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "api_internal.hpp"
#include "host_tables.hpp"
#include "host_tables_fft.hpp"
#include "kernels.cuh"
#include "kernels_fft.cuh"
#include "kernels_fft_latency.cuh"
#include "kernels_fft_cluster.cuh"
#include "kernels_ks_tcgen05.cuh"

using namespace tfhe;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct tfhe_ctx {
    tfhe_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int pbs_id = -1, ks_id = -1;
    int path = TFHE_PATH_NTT;               // arithmetic path of the external product (tfhe_ctx_set_pbs_path)
    bool fft_check = false;                 // FFT path: run the kernel variant that records the rounding margin
    int ks_path = TFHE_KS_TCGEN05;          // arithmetic of the key-switching product: TFHE_KS_IMAD / TFHE_KS_MMA (mma.sync) / TFHE_KS_TCGEN05 (tcgen05.mma + TMEM + TMA)
    int latency_cfg = 4;                    // FFT path, small batches: 0 throughput kernel, 1 one team + deep key ring, 2 all teams of a CTA on one ciphertext, 3 a cluster of L CTAs per ciphertext, 4 the same with every CTA split by key limb (then 2, then 0 as the batch grows)
    fft::cplx *d_ftw[4] = {};                // FFT pass-B / pass-C twiddle tables, zeta^m table (BMMP), per-lane table of the tensor-memory exchanges
    bool fft_tmem = true;                    // FFT path, N = 512: exchange the register passes of the transforms through tensor memory (fft_tmem.cuh)
    fft::TwTablesF ftw;
    unsigned long long *d_margin = nullptr;  // FFT path: largest distance to an integer seen before rounding
    std::string err;
    uint32_t *d_tw[2][4] = {};
    TwTables tw[2];
    PrimeTab prime[2];
    DevBuf in0, in1, in2, out, glwe, digits, body, luts, lutidx, misc;
    uint32_t *d_err = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    cudaEvent_t ev[5] = {};
    double last_ms[3] = {0, 0, 0};
    std::vector<tfhe_bk *> keys;             // keys uploaded through this ctx and not yet freed (orphaned by tfhe_ctx_destroy)
    size_t N() const { return (size_t)1 << p.glwe_poly_degree; }
    size_t k() const { return p.glwe_dimension; }
    size_t n() const { return p.lwe_dimension; }
    size_t glwe_words() const { return (k() + 1) * N(); }
    size_t ggsw_words() const { return (k() + 1) * p.pbs_levels * glwe_words(); }
    size_t kd() const { return k() * N() * p.ks_levels; }
};

struct tfhe_bk {
    tfhe_ctx *ctx = nullptr;        // null once the context is gone: the key can then only be freed
    int device = 0;                 // where the allocations below live (tfhe_bk_free does not touch ctx)
    int path = TFHE_PATH_NTT;
    bool bmmp = false;              // key triples of the unrolled-by-two blind rotation (FFT path only)
    uint32_t *d_bsk_ntt = nullptr;  // [n][2][ROWS][P][N]                      (TFHE_PATH_NTT)
    fft::cplx *d_bsk_fft = nullptr; // [n][ROWS][2 limbs][P][N/2], scaled 2/N  (TFHE_PATH_FFT)
    fft::cplx *d_bsk_fft_x = nullptr; // the same key, every polynomial in the spectral order of the tensor-memory-exchange kernel (N = 512)
    uint32_t *d_ksk = nullptr;      // [kN*l_ks][ksk_stride], ksk_stride = n+1 rounded up to 128 words, zero padded
    uint8_t *d_ksk_t = nullptr;     // [ksk_stride*4][kN*l_ks] byte planes, k contiguous (ks_mma_kernel, ks_tcgen05_kernel); null if not applicable
    CUtensorMap map_kskt;           // TMA descriptor of d_ksk_t (ks_tcgen05_kernel)
    bool tc5 = false;               // map_kskt is valid
    size_t ksk_stride = 0;
};

namespace {

int fail(tfhe_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? TFHE_E_OOM : TFHE_E_CUDA, b_);              \
        }                                                                                                  \
    } while (0)

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// device view of an input: device pointers pass through, host pointers are copied into `stage`
int stage_in(tfhe_ctx *ctx, const void *src, size_t bytes, DevBuf &stage, const void **dev) {
    if (is_device_ptr(src)) { *dev = src; return TFHE_OK; }
    CU(stage.ensure(bytes));
    CU(cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = stage.p;
    return TFHE_OK;
}
// device destination for an output (the user's buffer if it is a device pointer)
int stage_out(tfhe_ctx *ctx, void *dst, size_t bytes, DevBuf &stage, void **dev) {
    if (is_device_ptr(dst)) { *dev = dst; return TFHE_OK; }
    CU(stage.ensure(bytes));
    *dev = stage.p;
    return TFHE_OK;
}
int finish_out(tfhe_ctx *ctx, void *dst, size_t bytes, const void *dev) {
    if (dev != dst) CU(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return TFHE_OK;
}

#ifndef TFHE_DEFAULT_PATH
#define TFHE_DEFAULT_PATH TFHE_PATH_FFT   // path chosen for parameter sets that have both instantiations (measured faster)
#endif
// ---- kernel configurations (must match api_internal.hpp) ----
#ifndef TFHE_STAGE_P0
#define TFHE_STAGE_P0 1
#endif
#ifndef TFHE_STAGE_P1
#define TFHE_STAGE_P1 1
#endif
#ifndef TFHE_TWCG_P0
#define TFHE_TWCG_P0 1
#endif
using K0 = PbsCfg<9, 3, 2, 6, 4, TFHE_STAGE_P0 != 0, TFHE_TWCG_P0 != 0>;
#ifndef TFHE_TWCG_P1
#define TFHE_TWCG_P1 0
#endif
using K1 = PbsCfg<10, 4, 1, 3, 8, TFHE_STAGE_P1 != 0, TFHE_TWCG_P1 != 0>;
using K2 = PbsCfg<11, 4, 1, 3, 8, /*STAGE_G=*/false, /*TWC_GLOBAL=*/true>;  // 32 KB of row staging would halve its occupancy;
                                                                        // at 128 registers register-resident twiddles spill
template <class K> struct MinBlocks;
#ifndef TFHE_MINB_P0
#define TFHE_MINB_P0 4
#endif
#ifndef TFHE_MINB_P1
#define TFHE_MINB_P1 3
#endif
#ifndef TFHE_MINB_P2
#define TFHE_MINB_P2 2
#endif
template <> struct MinBlocks<K0> { static constexpr int v = TFHE_MINB_P0; };
template <> struct MinBlocks<K1> { static constexpr int v = TFHE_MINB_P1; };
template <> struct MinBlocks<K2> { static constexpr int v = TFHE_MINB_P2; };

template <class K>
int launch_pbs_t(tfhe_ctx *ctx, const PbsArgs &a) {
    const size_t smem = (size_t)K::SM_AT + (((size_t)a.n + 1) * 2 + 15) / 16 * 16;
    auto kern = pbs_kernel<K, MinBlocks<K>::v>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.batch, K::THREADS, smem, ctx->stream>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
    return TFHE_OK;
}
// ---- FP64-FFT path (kernels_fft.cuh): instantiated parameter sets
#ifndef TFHE_FFT_CTS_P1
#define TFHE_FFT_CTS_P1 3
#endif
#ifndef TFHE_FFT_P1_LOGE
#define TFHE_FFT_P1_LOGE 3
#endif
#ifndef TFHE_FFT_P1_SINGLE
#define TFHE_FFT_P1_SINGLE 0
#endif
#ifndef TFHE_FFT_P1_TWT
#define TFHE_FFT_P1_TWT 0   // 2: derived pass twiddles of P1 wait in tensor memory (bit-identical, measured neutral: 64.80 vs 64.93 ms -- P1 is bound by the shared-memory pipe, not by FP64 issue); 0: re-derived per pass
#endif
using KF1 = fft::FftPbsCfg<10, TFHE_FFT_P1_LOGE, 1, 3, 8, TFHE_FFT_CTS_P1, false, TFHE_FFT_P1_SINGLE != 0, 1, TFHE_FFT_NSLOT, TFHE_FFT_P1_TWT>;   // production: a-priori exactness bound only
using KF1C = fft::FftPbsCfg<10, TFHE_FFT_P1_LOGE, 1, 3, 8, TFHE_FFT_CTS_P1, true, TFHE_FFT_P1_SINGLE != 0, 1, TFHE_FFT_NSLOT, TFHE_FFT_P1_TWT>;
// P1 with the last three stages of every transform in tensor memory (one shared-memory exchange instead of two; fft_tmem.cuh tail9)
#ifndef TFHE_FFT_P1T_NSLOT
#define TFHE_FFT_P1T_NSLOT TFHE_FFT_NSLOT
#endif
using KF1T = fft::FftPbsCfg<10, 3, 1, 3, 8, TFHE_FFT_CTS_P1, false, false, 1, TFHE_FFT_P1T_NSLOT, 3>;
using KF1TC = fft::FftPbsCfg<10, 3, 1, 3, 8, TFHE_FFT_CTS_P1, true, false, 1, TFHE_FFT_P1T_NSLOT, 3>;   // + records the rounding margin (tests, validation)
#ifndef TFHE_FFT_CTS_P0
#define TFHE_FFT_CTS_P0 4
#endif
#ifndef TFHE_FFT_TMEM_NSLOT
#define TFHE_FFT_TMEM_NSLOT 6   // the shared memory the exchange buffers no longer need holds a deeper key ring
#endif
using KF0 = fft::FftPbsCfg<9, 3, 2, 6, 4, TFHE_FFT_CTS_P0, false>;    // reference defaults (lib.rs:101-123)
using KF0C = fft::FftPbsCfg<9, 3, 2, 6, 4, TFHE_FFT_CTS_P0, true>;
using KF0T = fft::FftPbsCfg<9, 3, 2, 6, 4, 4, false, false, 1, TFHE_FFT_TMEM_NSLOT, 1>;   // register passes exchanged through tensor memory
using KF0TC = fft::FftPbsCfg<9, 3, 2, 6, 4, 4, true, false, 1, TFHE_FFT_TMEM_NSLOT, 1>;
#ifndef TFHE_FFT_CTS_P2
#define TFHE_FFT_CTS_P2 2
#endif
// N = 2048: 16 points per thread, one exchange buffer per sub-team, half-row key slots (fft_team.cuh FftPbsCfg)
using KF2 = fft::FftPbsCfg<11, 4, 1, 3, 8, TFHE_FFT_CTS_P2, false, true, 2>;
using KF2C = fft::FftPbsCfg<11, 4, 1, 3, 8, TFHE_FFT_CTS_P2, true, true, 2>;
// N = 2048 with one shared-memory exchange per transform and the last two stages on a tensor-memory swap (fft_tmem.cuh tail16)
using KF2T = fft::FftPbsCfg<11, 4, 1, 3, 8, TFHE_FFT_CTS_P2, false, true, 2, TFHE_FFT_NSLOT, 3>;
using KF2TC = fft::FftPbsCfg<11, 4, 1, 3, 8, TFHE_FFT_CTS_P2, true, true, 2, TFHE_FFT_NSLOT, 3>;
static_assert(fft::key_slot_layout_ok<KF0>() && fft::key_slot_layout_ok<KF1>() && fft::key_slot_layout_ok<KF2>(), "diagonal-major key layout");
// the spectral layouts of the tensor-memory kernels are permutations of a polynomial's points (the key re-ordering kernels rely on it)
constexpr bool tmem_layouts_are_permutations() {
    bool seen8[256] = {}, seen9[512] = {}, seen10[1024] = {};
    for (uint32_t j = 0; j < 256; j++) { const uint32_t s = fft::tmem_slot_of_index(j); if (s >= 256 || seen8[s]) return false; seen8[s] = true; }
    for (uint32_t j = 0; j < 512; j++) { const uint32_t s = fft::tail9_slot_of_index(j); if (s >= 512 || seen9[s]) return false; seen9[s] = true; }
    for (uint32_t j = 0; j < 1024; j++) {
        const uint32_t s = fft::tail10_reg_of_index(j) * 64u + fft::tail10_thread_of_index(j);
        if (s >= 1024 || seen10[s]) return false;
        seen10[s] = true;
    }
    // the swizzles of the one shared-memory exchange are permutations too, and layout B'' covers every point exactly once
    bool b9[512] = {}, b10[1024] = {};
    for (uint32_t t = 0; t < 64; t++)
        for (uint32_t e = 0; e < 8; e++) { const uint32_t a = fft::swz9(fft::jbase_Bsw(t) | (e << 3)); if (a >= 512 || b9[a]) return false; b9[a] = true; }
    for (uint32_t t = 0; t < 64; t++)
        for (uint32_t e = 0; e < 16; e++) { const uint32_t a = fft::swz10(fft::jbase_Bsw16(t) | (e << 2)); if (a >= 1024 || b10[a]) return false; b10[a] = true; }
    return true;
}
static_assert(tmem_layouts_are_permutations(), "tensor-memory spectral layouts");
// latency configurations: ONE ciphertext per CTA, the rest of the shared memory is a deep key ring (batches of at most one
// ciphertext per SM; the production two-slot ring would leave such a CTA waiting for the round trip of every refill)
using KF0L = fft::FftPbsCfg<9, 3, 2, 6, 4, 1, false, false, 1, 7>;
using KF1L = fft::FftPbsCfg<10, TFHE_FFT_P1_LOGE, 1, 3, 8, 1, false, TFHE_FFT_P1_SINGLE != 0, 1, 5>;
using KF2L = fft::FftPbsCfg<11, 4, 1, 3, 8, 1, false, true, 2, 5>;
// all-teams-on-one-ciphertext configurations (kernels_fft_latency.cuh): the ring holds one GGSW row per team
using KF0H = fft::FftPbsCfg<9, 3, 2, 6, 4, 4, false, false, 1, 4>;
using KF1H = fft::FftPbsCfg<10, 3, 1, 3, 8, 3, false, false, 1, 3>;
// one-level-per-CTA cluster configurations (kernels_fft_cluster.cuh): cluster of L CTAs per ciphertext
using KF0X = fft::FftPbsCfg<9, 3, 2, 6, 4, 1, false, false, 1, 4>;
using KF1X = fft::FftPbsCfg<10, 3, 1, 3, 8, 1, false, false, 1, 4>;
// ... with each CTA's work split by key limb over twice the warps (two more exchange buffers per CTA: a three-row ring)
using KF0S = fft::FftPbsCfg<9, 3, 2, 6, 4, 1, false, false, 1, 3>;
using KF1S = fft::FftPbsCfg<10, 3, 1, 3, 8, 1, false, false, 1, 3>;
template <class K>
constexpr size_t fft_smem_bytes(size_t n) { return (size_t)K::CTS * K::team_bytes((int)n) + (size_t)K::NSLOT * K::SLOT_BYTES + 2 * K::NSLOT * 8 + 4 * K::NSLOT + 16; }
// the FFT path is instantiated for P0 and P1 shapes; its shared-memory layout holds the mod-switched mask of every
// resident ciphertext, which bounds the LWE dimension (P1 shape: n <= 1151, P0 shape: n <= 2111; above that the NTT path serves)
bool fft_available(int pbs_id, size_t n = 0) {
    if (pbs_id == 0) return fft_smem_bytes<KF0>(n) <= 227 * 1024;
    if (pbs_id == 1) return fft_smem_bytes<KF1>(n) <= 227 * 1024;
    if (pbs_id == 2) return fft_smem_bytes<KF2>(n) <= 227 * 1024;
    return false;
}

template <class K, bool BMMP = false>
int launch_pbs_fft_t(tfhe_ctx *ctx, const PbsArgs &a, const fft::cplx *key) {
    fft::FftArgs f = {};
    f.tw = ctx->ftw;
    f.bsk_fft = key;
    f.lwe_in = a.lwe_in; f.luts = a.luts; f.lut_idx = a.lut_idx;
    f.in0 = a.in0; f.in1 = a.in1; f.ggsw_index = a.ggsw_index;
    f.glwe_out = a.glwe_out; f.err_flag = a.err_flag; f.margin = ctx->d_margin;
    f.n = a.n; f.batch = a.batch; f.mode = a.mode; f.log_p = a.log_p; f.enc_shift = a.enc_shift; f.n_luts = a.n_luts;
    const size_t smem = fft_smem_bytes<K>(a.n);
    if (smem > 227 * 1024) return fail(ctx, TFHE_E_PARAM, "lwe_dimension too large for the FFT path's shared-memory layout");
    auto kern = fft::pbs_fft_kernel<K, BMMP>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)a.batch;
    if (a.mode == 0) {
        // whole waves of one CTA per SM; the kernel spreads the batch evenly over them (<= CTS ciphertexts per CTA)
        const unsigned ctas = (unsigned)((a.batch + K::CTS - 1) / K::CTS), sms = (unsigned)ctx->sm_count;
        // less than one wave: one CTA per SM with as few ciphertexts each as possible (shortest latency)
        grid = ctas <= sms ? (a.batch < sms ? (unsigned)a.batch : sms) : ((ctas + sms - 1) / sms) * sms;
        if (grid > a.batch) grid = (unsigned)a.batch;
    }
    kern<<<grid, K::THREADS, smem, ctx->stream>>>(f);
    CU(cudaGetLastError());
    ctx->launches++;
    return TFHE_OK;
}
// returns TFHE_OK, or -100 when the cluster shape cannot be launched on this device (the caller falls back)
template <class K, bool SPLIT = false>
int launch_pbs_fft_cluster_t(tfhe_ctx *ctx, const PbsArgs &a, const fft::cplx *key) {
    fft::FftArgs f = {};
    f.tw = ctx->ftw;
    f.bsk_fft = key;
    f.lwe_in = a.lwe_in; f.luts = a.luts; f.lut_idx = a.lut_idx;
    f.glwe_out = a.glwe_out; f.err_flag = a.err_flag; f.margin = ctx->d_margin;
    f.n = a.n; f.batch = a.batch; f.mode = 0; f.log_p = a.log_p; f.enc_shift = a.enc_shift; f.n_luts = a.n_luts;
    using Layout = typename std::conditional<SPLIT, fft::ClusterSplitLayout<K>, fft::ClusterLayout<K>>::type;
    const size_t smem = Layout::smem_bytes(a.n);
    if (smem > 227 * 1024) return -100;
    void (*kern)(fft::FftArgs);
    if constexpr (SPLIT) kern = fft::pbs_fft_cluster_split_kernel<K>;
    else kern = fft::pbs_fft_cluster_kernel<K>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return -100;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.batch * K::L), 1, 1);
    cfg.blockDim = dim3(SPLIT ? fft::ClusterSplitLayout<K>::THREADS : K::TEAM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = K::L;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    // all clusters of the batch must be co-resident (a second wave of clusters would double the latency: the all-teams kernel is
    // faster then); clusters are placed inside GPCs, so this is fewer than sm_count / L (measured: 24 clusters of 6 do not fit)
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1 || (int)a.batch > max_clusters) {
        cudaGetLastError();
        return -100;
    }
    unsigned long long *d_prof = nullptr;
    if (getenv("TFHE_B200_LAT_PROF")) {   // measurement only: phase cycles of the first CTA (kernels_fft_cluster.cuh)
        CU(cudaMalloc(&d_prof, 16 * 8));
        CU(cudaMemsetAsync(d_prof, 0, 16 * 8, ctx->stream));
        f.prof = d_prof;
    }
    if (cudaLaunchKernelEx(&cfg, kern, f) != cudaSuccess) {
        cudaGetLastError();
        if (d_prof) cudaFree(d_prof);
        return -100;
    }
    ctx->launches++;
    if (d_prof) {
        unsigned long long h[16];
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(d_prof);
        const char *lab[9] = {"digits + pass A", "forward rest", "wait own row", "mac own row", "team barrier", "wait + mac peer rows", "inverse + rounding + push",
                              "cluster barrier", "accumulate"};
        const char *labs[9] = {"digits + pass A", "forward rest", "barrier (rows published)", "mac (one limb, all rows)", "barrier (rows read)", "inverse (one limb) + rounding",
                               "combine limbs + push", "cluster barrier", "accumulate"};
        fprintf(stderr, "%s (L = %d CTAs per ciphertext), first CTA, cycles per step (n = %u):", SPLIT ? "cluster kernel, CTAs split by key limb" : "cluster kernel", K::L, a.n);
        for (int k = 0; k < 9; k++) fprintf(stderr, " %s %.0f,", SPLIT ? labs[k] : lab[k], (double)h[k] / a.n);
        fprintf(stderr, " total %.0f\n", (double)h[9] / a.n);
    }
    return TFHE_OK;
}
template <class K>
int launch_pbs_fft_latency_t(tfhe_ctx *ctx, const PbsArgs &a, const fft::cplx *key) {
    fft::FftArgs f = {};
    f.tw = ctx->ftw;
    f.bsk_fft = key;
    f.lwe_in = a.lwe_in; f.luts = a.luts; f.lut_idx = a.lut_idx;
    f.glwe_out = a.glwe_out; f.err_flag = a.err_flag; f.margin = ctx->d_margin;
    f.n = a.n; f.batch = a.batch; f.mode = 0; f.log_p = a.log_p; f.enc_shift = a.enc_shift; f.n_luts = a.n_luts;
    const size_t smem = fft::LatencyLayout<K>::smem_bytes(a.n);
    auto kern = fft::pbs_fft_latency_kernel<K>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long *d_prof = nullptr;
    if (getenv("TFHE_B200_LAT_PROF")) {   // measurement only: phase cycles of CTA 0 (kernels_fft_latency.cuh)
        CU(cudaMalloc(&d_prof, 32 * 8));
        CU(cudaMemsetAsync(d_prof, 0, 32 * 8, ctx->stream));
        f.prof = d_prof;
    }
    kern<<<(unsigned)a.batch, K::THREADS, smem, ctx->stream>>>(f);
    CU(cudaGetLastError());
    ctx->launches++;
    if (d_prof) {
        unsigned long long h[32];
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(d_prof);
        fprintf(stderr, "latency kernel, CTA 0, cycles per step (n = %u): owner: decompose %.0f, wait B1 %.0f, levels %.0f, wait B2 %.0f, add partials %.0f, inverse+update %.0f, total %.0f | "
                        "team 1: wait B1 %.0f (incl. idle since B2), levels %.0f, wait B2 %.0f\n", a.n,
                (double)h[0] / a.n, (double)h[1] / a.n, (double)h[2] / a.n, (double)h[3] / a.n, (double)h[4] / a.n, (double)h[5] / a.n, (double)h[6] / a.n,
                ((double)h[16] + h[17]) / a.n, (double)h[18] / a.n, (double)h[19] / a.n);
        for (int tm_ = 0; tm_ < 2; tm_++)
            fprintf(stderr, "   team %d levels split: forward %.0f, wait own row %.0f, mac own %.0f, team barrier %.0f, wait peer rows %.0f, mac peer rows %.0f\n", tm_,
                    (double)h[tm_ * 16 + 8] / a.n, (double)h[tm_ * 16 + 9] / a.n, (double)h[tm_ * 16 + 10] / a.n, (double)h[tm_ * 16 + 11] / a.n, (double)h[tm_ * 16 + 12] / a.n,
                    (double)h[tm_ * 16 + 13] / a.n);
    }
    return TFHE_OK;
}
template <class K>
int launch_fft_transform_t(tfhe_ctx *ctx, const uint32_t *raw, fft::cplx *out, size_t n, uint32_t keys_per_step = 1) {
    fft::FftTransformArgs ta;
    ta.tw = ctx->ftw; ta.raw = raw; ta.out = out; ta.keys_per_step = keys_per_step;
    const size_t smem = (size_t)2 * 2 * K::F::MPAD * 16;
    auto kern = fft::bsk_fft_transform_kernel<K>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(n * K::ROWS * K::P), 2 * K::T, smem, ctx->stream>>>(ta);
    CU(cudaGetLastError());
    ctx->launches++;
    return TFHE_OK;
}
int launch_fft_transform(tfhe_ctx *ctx, const uint32_t *raw, fft::cplx *out, size_t n_ggsw, uint32_t keys_per_step) {
    switch (ctx->pbs_id) {
    case 0: return launch_fft_transform_t<KF0>(ctx, raw, out, n_ggsw, keys_per_step);
    case 1: return launch_fft_transform_t<KF1>(ctx, raw, out, n_ggsw, keys_per_step);
    case 2: return launch_fft_transform_t<KF2>(ctx, raw, out, n_ggsw, keys_per_step);
    }
    return fail(ctx, TFHE_E_PARAM, "no FFT-path instantiation for this parameter set");
}
size_t fft_key_bytes(const tfhe_ctx *ctx) {
    return ctx->n() * (ctx->k() + 1) * ctx->p.pbs_levels * 2 * (ctx->k() + 1) * (ctx->N() / 2) * sizeof(fft::cplx);
}

int launch_pbs(tfhe_ctx *ctx, const PbsArgs &a, const tfhe_bk *bk) {
    if (bk->bmmp) {
        if (a.mode != 0) return fail(ctx, TFHE_E_PARAM, "a BMMP key serves blind rotations only (its GGSWs encrypt products of key bits)");
        switch (ctx->pbs_id) {
        case 0: return ctx->fft_check ? launch_pbs_fft_t<KF0C, true>(ctx, a, bk->d_bsk_fft) : launch_pbs_fft_t<KF0, true>(ctx, a, bk->d_bsk_fft);
        case 1: return ctx->fft_check ? launch_pbs_fft_t<KF1C, true>(ctx, a, bk->d_bsk_fft) : launch_pbs_fft_t<KF1, true>(ctx, a, bk->d_bsk_fft);
        }
        return fail(ctx, TFHE_E_PARAM, "no BMMP instantiation for this parameter set");
    }
    if (bk->path == TFHE_PATH_FFT) {
#ifndef TFHE_FFT_LATENCY_CFG
#define TFHE_FFT_LATENCY_CFG 1
#endif
        if (TFHE_FFT_LATENCY_CFG && ctx->latency_cfg && a.mode == 0 && !ctx->fft_check && a.batch <= (uint32_t)ctx->sm_count) {
            if (ctx->latency_cfg == 4) {   // one CTA per gadget level, a cluster per ciphertext, every CTA split by key limb
                int rc = -100;
                if (ctx->pbs_id == 0 && a.batch * KF0S::L <= (uint32_t)ctx->sm_count) rc = launch_pbs_fft_cluster_t<KF0S, true>(ctx, a, bk->d_bsk_fft);
                if (ctx->pbs_id == 1 && a.batch * KF1S::L <= (uint32_t)ctx->sm_count) rc = launch_pbs_fft_cluster_t<KF1S, true>(ctx, a, bk->d_bsk_fft);
                if (rc != -100) return rc;
            }
            if (ctx->latency_cfg >= 3) {   // one CTA per gadget level, a cluster per ciphertext
                int rc = -100;
                if (ctx->pbs_id == 0 && a.batch * KF0X::L <= (uint32_t)ctx->sm_count && fft::ClusterLayout<KF0X>::smem_bytes(a.n) <= 227 * 1024)
                    rc = launch_pbs_fft_cluster_t<KF0X>(ctx, a, bk->d_bsk_fft);
                if (ctx->pbs_id == 1 && a.batch * KF1X::L <= (uint32_t)ctx->sm_count && fft::ClusterLayout<KF1X>::smem_bytes(a.n) <= 227 * 1024)
                    rc = launch_pbs_fft_cluster_t<KF1X>(ctx, a, bk->d_bsk_fft);
                if (rc != -100) return rc;
            }
            if (ctx->latency_cfg >= 2) {
                if (ctx->pbs_id == 0 && fft::LatencyLayout<KF0H>::smem_bytes(a.n) <= 227 * 1024) return launch_pbs_fft_latency_t<KF0H>(ctx, a, bk->d_bsk_fft);
                if (ctx->pbs_id == 1 && fft::LatencyLayout<KF1H>::smem_bytes(a.n) <= 227 * 1024) return launch_pbs_fft_latency_t<KF1H>(ctx, a, bk->d_bsk_fft);
            }
            switch (ctx->pbs_id) {
            case 0: if (fft_smem_bytes<KF0L>(a.n) <= 227 * 1024) return launch_pbs_fft_t<KF0L>(ctx, a, bk->d_bsk_fft); break;
            case 1: if (fft_smem_bytes<KF1L>(a.n) <= 227 * 1024) return launch_pbs_fft_t<KF1L>(ctx, a, bk->d_bsk_fft); break;
            case 2: if (fft_smem_bytes<KF2L>(a.n) <= 227 * 1024) return launch_pbs_fft_t<KF2L>(ctx, a, bk->d_bsk_fft); break;
            }
        }
        switch (ctx->pbs_id) {
        case 0:
            if (a.mode == 0 && bk->d_bsk_fft_x && ctx->fft_tmem && fft_smem_bytes<KF0T>(a.n) <= 227 * 1024)
                return ctx->fft_check ? launch_pbs_fft_t<KF0TC>(ctx, a, bk->d_bsk_fft_x) : launch_pbs_fft_t<KF0T>(ctx, a, bk->d_bsk_fft_x);
            return ctx->fft_check ? launch_pbs_fft_t<KF0C>(ctx, a, bk->d_bsk_fft) : launch_pbs_fft_t<KF0>(ctx, a, bk->d_bsk_fft);
        case 1:
            if (a.mode == 0 && bk->d_bsk_fft_x && ctx->fft_tmem && fft_smem_bytes<KF1T>(a.n) <= 227 * 1024)
                return ctx->fft_check ? launch_pbs_fft_t<KF1TC>(ctx, a, bk->d_bsk_fft_x) : launch_pbs_fft_t<KF1T>(ctx, a, bk->d_bsk_fft_x);
            return ctx->fft_check ? launch_pbs_fft_t<KF1C>(ctx, a, bk->d_bsk_fft) : launch_pbs_fft_t<KF1>(ctx, a, bk->d_bsk_fft);
        case 2:
            if (a.mode == 0 && bk->d_bsk_fft_x && ctx->fft_tmem && fft_smem_bytes<KF2T>(a.n) <= 227 * 1024)
                return ctx->fft_check ? launch_pbs_fft_t<KF2TC>(ctx, a, bk->d_bsk_fft_x) : launch_pbs_fft_t<KF2T>(ctx, a, bk->d_bsk_fft_x);
            return ctx->fft_check ? launch_pbs_fft_t<KF2C>(ctx, a, bk->d_bsk_fft) : launch_pbs_fft_t<KF2>(ctx, a, bk->d_bsk_fft);
        }
        return fail(ctx, TFHE_E_PARAM, "no FFT-path instantiation for this parameter set");
    }
    switch (ctx->pbs_id) {
    case 0: return launch_pbs_t<K0>(ctx, a);
    case 1: return launch_pbs_t<K1>(ctx, a);
    case 2: return launch_pbs_t<K2>(ctx, a);
    }
    return fail(ctx, TFHE_E_PARAM, "no kernel instantiation for this parameter set");
}
template <class K>
int launch_transform_t(tfhe_ctx *ctx, const uint32_t *raw, uint32_t *out, size_t n) {
    const size_t polys = n * K::ROWS * K::P;
    TransformArgs ta;
    ta.prime[0] = ctx->prime[0]; ta.prime[1] = ctx->prime[1];
    ta.tw[0] = ctx->tw[0]; ta.tw[1] = ctx->tw[1];
    ta.raw = raw; ta.out = out;
    bsk_transform_kernel<K><<<(unsigned)polys, K::THREADS, 0, ctx->stream>>>(ta);
    CU(cudaGetLastError());
    ctx->launches++;
    return TFHE_OK;
}
int launch_transform(tfhe_ctx *ctx, const uint32_t *raw, uint32_t *out, size_t n) {
    switch (ctx->pbs_id) {
    case 0: return launch_transform_t<K0>(ctx, raw, out, n);
    case 1: return launch_transform_t<K1>(ctx, raw, out, n);
    case 2: return launch_transform_t<K2>(ctx, raw, out, n);
    }
    return fail(ctx, TFHE_E_PARAM, "no kernel instantiation for this parameter set");
}

// TMA descriptor of a K-major byte matrix [rows][kd] with a [box_rows][128-byte] box and the 128-byte swizzle the tcgen05
// shared-memory descriptors of ks_tcgen05_kernel expect.  cuTensorMapEncodeTiled is a driver entry point: resolved at run time.
bool make_byte_tensor_map(CUtensorMap *map, const void *base, size_t rows, size_t kd, uint32_t box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        encode = (encode_fn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)kd, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kd};
    const cuuint32_t box[2] = {(cuuint32_t)KT_BK, box_rows}, estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// key switch of `batch` ciphertexts; src is GLWE accumulators [B][P][N] (from_lwe=0) or extracted LWEs
// [B][kN+1] (from_lwe=1), all device pointers.
int run_key_switch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *src, int from_lwe, size_t batch, uint32_t *d_out) {
    const size_t KD = ctx->kd(), kN = ctx->k() * ctx->N();
    CU(ctx->digits.ensure(batch * KD));
    CU(ctx->body.ensure(batch * sizeof(uint32_t)));
    const size_t total = batch * kN;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    int8_t *dg = (int8_t *)ctx->digits.p;
    uint32_t *bd = (uint32_t *)ctx->body.p;
    if (ctx->ks_id == 0)
        ks_digits_kernel<4, 5><<<blocks, 256, 0, ctx->stream>>>(src, dg, bd, (uint32_t)ctx->k(), ctx->p.glwe_poly_degree, (uint32_t)batch, from_lwe);
    else
        ks_digits_kernel<2, 8><<<blocks, 256, 0, ctx->stream>>>(src, dg, bd, (uint32_t)ctx->k(), ctx->p.glwe_poly_degree, (uint32_t)batch, from_lwe);
    CU(cudaGetLastError());
    if (batch <= (size_t)KSV_MAXB) {
        // latency path: rows of the KSK split over ~2 CTAs per SM, partial sums combined with u32 atomics
        const size_t len = batch * (ctx->n() + 1);
        ks_init_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(bd, d_out, (uint32_t)ctx->n(), (uint32_t)batch);
        CU(cudaGetLastError());
        const uint32_t ctas = (uint32_t)ctx->sm_count * 2;
        const uint32_t rows_per_cta = (uint32_t)((KD + ctas - 1) / ctas);
        ks_gemv_kernel<<<(unsigned)((KD + rows_per_cta - 1) / rows_per_cta), KSV_THREADS, 0, ctx->stream>>>(
            dg, bk->d_ksk, d_out, (uint32_t)KD, (uint32_t)ctx->n(), (uint32_t)batch, rows_per_cta, (uint32_t)bk->ksk_stride);
        CU(cudaGetLastError());
        ctx->launches += 3;
        return TFHE_OK;
    }
    if (bk->tc5 && ctx->ks_path == TFHE_KS_TCGEN05) {
        // 5th-generation tensor cores: tcgen05.mma kind::i8, accumulator in TMEM, operands by TMA (kernels_ks_tcgen05.cuh)
        CUtensorMap map_a;
        if (!make_byte_tensor_map(&map_a, dg, batch, KD, KT_BM)) return fail(ctx, TFHE_E_CUDA, "cuTensorMapEncodeTiled failed for the digit matrix");
        dim3 grid((unsigned)(bk->ksk_stride * 4 / KT_BN), (unsigned)((batch + KT_BM - 1) / KT_BM));
        ks_tcgen05_kernel<<<grid, KT_THREADS, KT_SMEM, ctx->stream>>>(map_a, bk->map_kskt, bd, d_out, (uint32_t)KD, (uint32_t)ctx->n(), (uint32_t)batch, ctx->d_err);
        CU(cudaGetLastError());
        ctx->launches += 2;
        return TFHE_OK;
    }
    if (bk->d_ksk_t && ctx->ks_path != TFHE_KS_IMAD) {   // TFHE_KS_MMA, or TFHE_KS_TCGEN05 where that form does not apply
        // integer tensor cores, exact through byte planes (kernels.cuh K4-MMA)
        dim3 grid((unsigned)(bk->ksk_stride * 4 / KM_BN), (unsigned)((batch + KM_BM - 1) / KM_BM));
        ks_mma_kernel<<<grid, KM_THREADS, KM_SMEM, ctx->stream>>>(dg, bk->d_ksk_t, bd, d_out, (uint32_t)KD, (uint32_t)ctx->n(), (uint32_t)batch);
        CU(cudaGetLastError());
        ctx->launches += 2;
        return TFHE_OK;
    }
    dim3 grid((unsigned)((ctx->n() + 1 + KS_BN - 1) / KS_BN), (unsigned)((batch + KS_BM - 1) / KS_BM));
    if (KD % KS_BK) return fail(ctx, TFHE_E_PARAM, "k*N*ks_levels must be a multiple of 32");
    ks_gemm_kernel<<<grid, KS_THREADS, 0, ctx->stream>>>(dg, bk->d_ksk, bd, d_out, (uint32_t)KD, (uint32_t)ctx->n(), (uint32_t)batch, (uint32_t)bk->ksk_stride);
    CU(cudaGetLastError());
    ctx->launches += 2;
    return TFHE_OK;
}

// device KSK: rows padded with zeros to a multiple of 128 words (aligned 128-bit tile loads in ks_gemm_kernel)
cudaError_t alloc_ksk(tfhe_ctx *ctx, tfhe_bk *bk) {
    bk->ksk_stride = (ctx->n() + 1 + 127) / 128 * 128;
    cudaError_t e = cudaMalloc(&bk->d_ksk, ctx->kd() * bk->ksk_stride * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(bk->d_ksk, 0, ctx->kd() * bk->ksk_stride * 4, ctx->stream);
    return e;
}
cudaError_t copy_ksk(tfhe_ctx *ctx, tfhe_bk *bk, const uint32_t *ksk) {
    const size_t row = (ctx->n() + 1) * 4, KD = ctx->kd();
    cudaError_t e = cudaMemcpy2DAsync(bk->d_ksk, bk->ksk_stride * 4, ksk, row, row, KD,
                                      is_device_ptr(ksk) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return e;
    // byte-plane copy for the tensor-core key switch: needs whole 64-digit stages and s8 x u8 sums that fit 32 bits
    const double bound = (double)KD * (double)(1u << ctx->p.ks_log_base) * 255.0;
    if (KD % KM_BK == 0 && ctx->p.ks_log_base <= 6 && bound < 2147483648.0) {
        if ((e = cudaMalloc(&bk->d_ksk_t, bk->ksk_stride * 4 * KD)) != cudaSuccess) return e;
        dim3 grid((unsigned)((bk->ksk_stride + 127) / 128), (unsigned)(KD / 4));
        ksk_byte_transpose_kernel<<<grid, 128, 0, ctx->stream>>>(bk->d_ksk, bk->d_ksk_t, (uint32_t)KD, (uint32_t)bk->ksk_stride);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM)) != cudaSuccess) return e;
        // tcgen05 form: whole 128-byte K slabs and 256-column tiles (ksk_stride is a multiple of 128 words = 512 byte columns)
        if (KD % KT_BK == 0 && make_byte_tensor_map(&bk->map_kskt, bk->d_ksk_t, bk->ksk_stride * 4, KD, KT_BN) &&
            cudaFuncSetAttribute(ks_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM) == cudaSuccess)
            bk->tc5 = true;
        else
            cudaGetLastError();
        ctx->launches++;
    }
    return cudaSuccess;
}

int check_bk(tfhe_ctx *ctx, const tfhe_bk *bk) {
    if (!ctx) return TFHE_E_PARAM;
    if (!bk || bk->ctx != ctx) return fail(ctx, TFHE_E_PARAM, "bootstrapping key does not belong to this context (or its context was destroyed)");
    return TFHE_OK;
}

// core of bootstrap: d_in [B][n+1] device, luts/lut_idx device (lut_idx may be null) -> d_out [B][n+1]
int run_bootstrap(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *d_in, const uint32_t *d_luts, size_t n_luts, const uint32_t *d_lut_idx,
                  size_t batch, uint32_t *d_out) {
    CU(ctx->glwe.ensure(batch * ctx->glwe_words() * 4));
    CU(cudaMemsetAsync(ctx->d_err, 0, 4, ctx->stream));
    PbsArgs a = {};
    a.bsk_ntt = bk->d_bsk_ntt;
    a.tw[0] = ctx->tw[0]; a.tw[1] = ctx->tw[1];
    a.prime[0] = ctx->prime[0]; a.prime[1] = ctx->prime[1];
    a.lwe_in = d_in; a.luts = d_luts; a.lut_idx = d_lut_idx;
    a.glwe_out = (uint32_t *)ctx->glwe.p;
    a.err_flag = ctx->d_err;
    a.n = (uint32_t)ctx->n(); a.batch = (uint32_t)batch; a.mode = 0;
    a.log_p = ctx->p.log_p;
    a.enc_shift = ctx->p.log_q - (ctx->p.log_p + ctx->p.padding_bits);
    a.n_luts = (uint32_t)n_luts;
    if (const char *e = getenv("TFHE_B200_SKEW_NS")) {
        a.skew_ns = (uint32_t)atoi(e);
        a.skew_div = (uint32_t)ctx->sm_count;
        a.skew_mod = 3;
        if (const char *m = getenv("TFHE_B200_SKEW_MOD")) a.skew_mod = (uint32_t)atoi(m);
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    int rc = launch_pbs(ctx, a, bk);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    rc = run_key_switch(ctx, bk, (const uint32_t *)ctx->glwe.p, 0, batch, d_out);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    return TFHE_OK;
}
// err_flag written by the blind-rotation kernels (kernels.cuh PbsArgs::err_flag)
int decode_err_flag(tfhe_ctx *ctx, uint32_t flag) {
    if (flag & 2u) return fail(ctx, TFHE_E_CUDA, "TMA bulk copy wait timed out in the blind-rotation kernel (results are invalid)");
    if (flag & 4u) return fail(ctx, TFHE_E_PARAM, "lut_idx entry >= n_luts");
    if (flag & 1u) return fail(ctx, TFHE_E_ASSERT, "test vector entry >= 2^log_p (reference assert! glwe.rs:144)");
    return TFHE_OK;
}
int finish_timed(tfhe_ctx *ctx) {
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t = 0;
    CU(cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2])); ctx->last_ms[0] = t;
    CU(cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3])); ctx->last_ms[1] = t;
    CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[4])); ctx->last_ms[2] = t;
    uint32_t flag = 0;
    CU(cudaMemcpy(&flag, ctx->d_err, 4, cudaMemcpyDeviceToHost));
    return decode_err_flag(ctx, flag);
}

template <class K>
int launch_polymul_t(tfhe_ctx *ctx, const int32_t *a, const uint32_t *g, uint32_t *out, size_t batch) {
    PolyMulArgs pa;
    pa.prime[0] = ctx->prime[0]; pa.prime[1] = ctx->prime[1];
    pa.tw[0] = ctx->tw[0]; pa.tw[1] = ctx->tw[1];
    pa.a = a; pa.g = g; pa.out = out;
    polymul_kernel<K><<<(unsigned)batch, K::THREADS, 0, ctx->stream>>>(pa);
    CU(cudaGetLastError());
    ctx->launches++;
    return TFHE_OK;
}

}  // namespace

extern "C" {

int tfhe_ctx_create(const tfhe_params *p, int device, tfhe_ctx **out) {
    if (!p || !out) return TFHE_E_PARAM;
    if (tfhe_params_validate(p) != TFHE_OK) return TFHE_E_PARAM;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return TFHE_E_CUDA;  // no CPU fallback by design
    }
    tfhe_ctx *ctx = new tfhe_ctx();
    ctx->p = *p;
    ctx->device = device;
    ctx->pbs_id = tfhe_host::pbs_config_id(*p);
    ctx->ks_id = tfhe_host::ks_config_id(*p);
    auto bail = [&](const char *what) {
        fprintf(stderr, "tfhe_ctx_create: %s failed: %s\n", what, cudaGetErrorString(cudaGetLastError()));
        tfhe_ctx_destroy(ctx);   // releases whatever was created so far (stream, events, device tables)
        return (int)TFHE_E_CUDA;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate");
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return bail("cudaEventCreate");
    if (cudaMalloc(&ctx->d_err, 4) != cudaSuccess) return bail("cudaMalloc");
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    // per-thread twiddle tables
    const int loge = (p->glwe_poly_degree == 9) ? 3 : 4;
    HostTw tw;
    build_tw_tables((int)p->glwe_poly_degree, loge, tw);
    fill_prime_tab(0, (int)p->glwe_poly_degree, loge, ctx->prime[0]);
    fill_prime_tab(1, (int)p->glwe_poly_degree, loge, ctx->prime[1]);
    for (int pr = 0; pr < 2; pr++) {
        const std::vector<uint32_t> *src[4] = {&tw.fwdB[pr], &tw.fwdC[pr], &tw.invB[pr], &tw.invC[pr]};
        for (int i = 0; i < 4; i++) {
            if (cudaMalloc(&ctx->d_tw[pr][i], src[i]->size() * 4) != cudaSuccess) return bail("cudaMalloc");
            if (cudaMemcpy(ctx->d_tw[pr][i], src[i]->data(), src[i]->size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
                return bail("cudaMemcpy");
        }
        ctx->tw[pr].fwdB = (const uint2 *)ctx->d_tw[pr][0];
        ctx->tw[pr].fwdC = (const uint2 *)ctx->d_tw[pr][1];
        ctx->tw[pr].invB = (const uint2 *)ctx->d_tw[pr][2];
        ctx->tw[pr].invC = (const uint2 *)ctx->d_tw[pr][3];
    }
    if (fft_available(ctx->pbs_id, ctx->n())) {
        const int logm = (int)p->glwe_poly_degree - 1, floge = ctx->pbs_id == 2 ? 4 : ctx->pbs_id == 1 ? TFHE_FFT_P1_LOGE : 3;   // = FftPbsCfg::F::LOGE of the instantiation
        fft::HostFftTw ft;
        fft::build_fft_tables(logm, floge, ft);
        for (size_t i = 0; i < ft.A.size(); i++) ctx->ftw.twA[i] = ft.A[i];
        std::vector<fft::cplx> fx;
        if (logm == 8) fft::build_fft_tmem_table(fx);
        if (logm == 9 && floge == 3) fft::build_fft_tail_table(fx);
        if (logm == 10 && floge == 4) fft::build_fft_tail16_table(fx);
        const std::vector<fft::cplx> *fsrc[4] = {&ft.B, &ft.C, &ft.Z, &fx};
        for (int i = 0; i < 4; i++) {
            if (fsrc[i]->empty()) continue;
            if (cudaMalloc(&ctx->d_ftw[i], fsrc[i]->size() * sizeof(fft::cplx)) != cudaSuccess) return bail("cudaMalloc");
            if (cudaMemcpy(ctx->d_ftw[i], fsrc[i]->data(), fsrc[i]->size() * sizeof(fft::cplx), cudaMemcpyHostToDevice) != cudaSuccess)
                return bail("cudaMemcpy");
        }
        ctx->ftw.twB = ctx->d_ftw[0];
        ctx->ftw.twC = ctx->d_ftw[1];
        ctx->ftw.ztab = ctx->d_ftw[2];
        ctx->ftw.twX = ctx->d_ftw[3];
        if (cudaMalloc(&ctx->d_margin, 8) != cudaSuccess || cudaMemset(ctx->d_margin, 0, 8) != cudaSuccess) return bail("cudaMalloc");
        ctx->path = TFHE_DEFAULT_PATH;
    }
    if (const char *e = getenv("TFHE_B200_FFT_CHECK")) ctx->fft_check = atoi(e) != 0;
    if (const char *e = getenv("TFHE_B200_FFT_TMEM")) ctx->fft_tmem = atoi(e) != 0;
    if (const char *e = getenv("TFHE_B200_KS")) ctx->ks_path = !strcmp(e, "imad") ? TFHE_KS_IMAD : !strcmp(e, "tcgen05") ? TFHE_KS_TCGEN05 : TFHE_KS_MMA;
    if (const char *e = getenv("TFHE_B200_LATENCY_CFG")) { const int v = atoi(e); if (v >= 0 && v <= 4) ctx->latency_cfg = v; }
    if (const char *e = getenv("TFHE_B200_PBS_PATH")) {
        if (!strcmp(e, "fft") && fft_available(ctx->pbs_id, ctx->n())) ctx->path = TFHE_PATH_FFT;
        if (!strcmp(e, "ntt")) ctx->path = TFHE_PATH_NTT;
    }
    *out = ctx;
    return TFHE_OK;
}

void tfhe_ctx_destroy(tfhe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (tfhe_bk *k : ctx->keys) k->ctx = nullptr;   // keys outlive the ctx as plain allocations: tfhe_bk_free still works
    ctx->keys.clear();
    for (DevBuf *b : {&ctx->in0, &ctx->in1, &ctx->in2, &ctx->out, &ctx->glwe, &ctx->digits, &ctx->body, &ctx->luts, &ctx->lutidx, &ctx->misc})
        b->release();
    for (int pr = 0; pr < 2; pr++)
        for (int i = 0; i < 4; i++)
            if (ctx->d_tw[pr][i]) cudaFree(ctx->d_tw[pr][i]);
    if (ctx->d_err) cudaFree(ctx->d_err);
    for (int i = 0; i < 4; i++)
        if (ctx->d_ftw[i]) cudaFree(ctx->d_ftw[i]);
    if (ctx->d_margin) cudaFree(ctx->d_margin);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *tfhe_last_error(const tfhe_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int tfhe_ctx_set_stream(tfhe_ctx *ctx, void *s) {
    if (!ctx) return TFHE_E_PARAM;
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (s) { ctx->stream = (cudaStream_t)s; ctx->own_stream = false; }
    else {
        ctx->own_stream = true;
        CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    }
    return TFHE_OK;
}
uint64_t tfhe_ctx_launch_count(const tfhe_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *tfhe_ctx_get_stream(const tfhe_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int tfhe_ctx_set_pbs_path(tfhe_ctx *ctx, int path) {
    if (!ctx) return TFHE_E_PARAM;
    if (path == TFHE_PATH_NTT) { ctx->path = path; return TFHE_OK; }
    if (path == TFHE_PATH_FFT && fft_available(ctx->pbs_id, ctx->n())) { ctx->path = path; return TFHE_OK; }
    return fail(ctx, TFHE_E_PARAM, "arithmetic path not instantiated for this parameter set");
}
int tfhe_ctx_get_pbs_path(const tfhe_ctx *ctx) { return ctx ? ctx->path : TFHE_E_PARAM; }
int tfhe_ctx_set_ks_path(tfhe_ctx *ctx, int path) {
    if (!ctx) return TFHE_E_PARAM;
    if (path != TFHE_KS_IMAD && path != TFHE_KS_MMA && path != TFHE_KS_TCGEN05) return fail(ctx, TFHE_E_PARAM, "unknown key-switch path");
    ctx->ks_path = path;
    return TFHE_OK;
}
int tfhe_ctx_set_latency_config(tfhe_ctx *ctx, int on) {
    if (!ctx) return TFHE_E_PARAM;
    if (on < 0 || on > 4) return fail(ctx, TFHE_E_PARAM, "latency configuration: 0 .. 4");
    ctx->latency_cfg = on;
    return TFHE_OK;
}
int tfhe_ctx_set_fft_exchange(tfhe_ctx *ctx, int tensor_memory) {
    if (!ctx) return TFHE_E_PARAM;
    ctx->fft_tmem = tensor_memory != 0;
    return TFHE_OK;
}
int tfhe_ctx_set_fft_check(tfhe_ctx *ctx, int on) {
    if (!ctx) return TFHE_E_PARAM;
    ctx->fft_check = on != 0;
    return TFHE_OK;
}
int tfhe_fft_rounding_margin(tfhe_ctx *ctx, double *out) {
    if (!ctx || !out) return TFHE_E_PARAM;
    *out = 0.0;
    if (!ctx->d_margin) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned long long bits = 0;
    CU(cudaMemcpy(&bits, ctx->d_margin, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemset(ctx->d_margin, 0, 8));
    memcpy(out, &bits, 8);
    return TFHE_OK;
}

int tfhe_bk_upload(tfhe_ctx *ctx, const uint32_t *bsk, const uint32_t *ksk, tfhe_bk **out) {
    if (!ctx || !bsk || !ksk || !out) return TFHE_E_PARAM;
    CU(cudaSetDevice(ctx->device));
    const size_t bsk_words = ctx->n() * ctx->ggsw_words();
    const size_t ksk_words = ctx->kd() * (ctx->n() + 1);
    tfhe_bk *bk = new tfhe_bk();
    bk->ctx = ctx;
    bk->device = ctx->device;
    bk->path = ctx->path;
    auto cleanup = [&]() { tfhe_bk_free(bk); };
    cudaError_t e;
    e = bk->path == TFHE_PATH_FFT ? cudaMalloc(&bk->d_bsk_fft, fft_key_bytes(ctx)) : cudaMalloc(&bk->d_bsk_ntt, bsk_words * 2 * 4);
    if (e != cudaSuccess || (e = alloc_ksk(ctx, bk)) != cudaSuccess) {
        cleanup();
        return fail(ctx, TFHE_E_OOM, cudaGetErrorString(e));
    }
    uint32_t *d_raw = nullptr;
    const uint32_t *raw_dev = bsk;
    if (!is_device_ptr(bsk)) {
        if ((e = cudaMalloc(&d_raw, bsk_words * 4)) != cudaSuccess) { cleanup(); return fail(ctx, TFHE_E_OOM, cudaGetErrorString(e)); }
        if ((e = cudaMemcpyAsync(d_raw, bsk, bsk_words * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
            cudaFree(d_raw); cleanup();
            return fail(ctx, TFHE_E_CUDA, cudaGetErrorString(e));
        }
        raw_dev = d_raw;
    }
    e = copy_ksk(ctx, bk, ksk);
    int rc;
    if (e != cudaSuccess) rc = fail(ctx, TFHE_E_CUDA, cudaGetErrorString(e));
    else if (bk->path == TFHE_PATH_FFT) {
        rc = launch_fft_transform(ctx, raw_dev, bk->d_bsk_fft, ctx->n(), 1);
        if (rc == TFHE_OK && (ctx->pbs_id == 0 || (ctx->pbs_id == 1 && TFHE_FFT_P1_LOGE == 3) || ctx->pbs_id == 2) && ctx->fft_tmem) {   // second copy in the order of the tensor-memory kernels
            if ((e = cudaMalloc(&bk->d_bsk_fft_x, fft_key_bytes(ctx))) != cudaSuccess) rc = fail(ctx, TFHE_E_OOM, cudaGetErrorString(e));
            else {
                const size_t M = ctx->N() / 2, polys = fft_key_bytes(ctx) / (M * sizeof(fft::cplx));
                if (ctx->pbs_id == 0) fft::bsk_fft_reslot_kernel<<<(unsigned)polys, 256, 0, ctx->stream>>>(bk->d_bsk_fft, bk->d_bsk_fft_x, polys);
                else if (ctx->pbs_id == 2) {
                    const unsigned rows = (unsigned)(ctx->n() * KF2T::ROWS), pps = 2u * KF2T::P;
                    fft::bsk_fft_reslot10_kernel<<<dim3(rows, pps), 1024, 0, ctx->stream>>>(bk->d_bsk_fft, bk->d_bsk_fft_x, pps);
                } else fft::bsk_fft_reslot9_kernel<<<(unsigned)polys, 512, 0, ctx->stream>>>(bk->d_bsk_fft, bk->d_bsk_fft_x, polys);
                ctx->launches++;
                if ((e = cudaGetLastError()) != cudaSuccess) rc = fail(ctx, TFHE_E_CUDA, cudaGetErrorString(e));
            }
        }
    } else rc = launch_transform(ctx, raw_dev, bk->d_bsk_ntt, ctx->n());
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (d_raw) cudaFree(d_raw);
    if (rc == TFHE_OK && es != cudaSuccess) rc = fail(ctx, TFHE_E_CUDA, cudaGetErrorString(es));
    if (rc != TFHE_OK) { cleanup(); return rc; }
    ctx->keys.push_back(bk);
    *out = bk;
    return TFHE_OK;
}

int tfhe_bk_upload_bmmp(tfhe_ctx *ctx, const uint32_t *bsk3, const uint32_t *ksk, tfhe_bk **out) {
    if (!ctx || !bsk3 || !ksk || !out) return TFHE_E_PARAM;
    if (!fft_available(ctx->pbs_id, ctx->n()) || ctx->pbs_id == 2)
        return fail(ctx, TFHE_E_PARAM, "the BMMP variant is instantiated for the P0 and P1 shapes only");
    if (ctx->n() & 1) return fail(ctx, TFHE_E_PARAM, "the BMMP variant needs an even lwe_dimension");
    CU(cudaSetDevice(ctx->device));
    const size_t n_ggsw = 3 * (ctx->n() / 2);
    const size_t bsk_words = n_ggsw * ctx->ggsw_words(), ksk_words = ctx->kd() * (ctx->n() + 1);
    tfhe_bk *bk = new tfhe_bk();
    bk->ctx = ctx;
    bk->device = ctx->device;
    bk->path = TFHE_PATH_FFT;
    bk->bmmp = true;
    auto cleanup = [&]() { tfhe_bk_free(bk); };
    cudaError_t e;
    if ((e = cudaMalloc(&bk->d_bsk_fft, fft_key_bytes(ctx) / ctx->n() * n_ggsw)) != cudaSuccess || (e = alloc_ksk(ctx, bk)) != cudaSuccess) {
        cleanup();
        return fail(ctx, TFHE_E_OOM, cudaGetErrorString(e));
    }
    uint32_t *d_raw = nullptr;
    const uint32_t *raw_dev = bsk3;
    if (!is_device_ptr(bsk3)) {
        if ((e = cudaMalloc(&d_raw, bsk_words * 4)) != cudaSuccess) { cleanup(); return fail(ctx, TFHE_E_OOM, cudaGetErrorString(e)); }
        if ((e = cudaMemcpyAsync(d_raw, bsk3, bsk_words * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
            cudaFree(d_raw); cleanup();
            return fail(ctx, TFHE_E_CUDA, cudaGetErrorString(e));
        }
        raw_dev = d_raw;
    }
    e = copy_ksk(ctx, bk, ksk);
    int rc = e != cudaSuccess ? fail(ctx, TFHE_E_CUDA, cudaGetErrorString(e)) : launch_fft_transform(ctx, raw_dev, bk->d_bsk_fft, n_ggsw, 3);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (d_raw) cudaFree(d_raw);
    if (rc == TFHE_OK && es != cudaSuccess) rc = fail(ctx, TFHE_E_CUDA, cudaGetErrorString(es));
    if (rc != TFHE_OK) { cleanup(); return rc; }
    ctx->keys.push_back(bk);
    *out = bk;
    return TFHE_OK;
}

void tfhe_bk_free(tfhe_bk *bk) {
    if (!bk) return;
    if (bk->ctx) {
        auto &ks = bk->ctx->keys;
        for (size_t i = 0; i < ks.size(); i++)
            if (ks[i] == bk) { ks.erase(ks.begin() + i); break; }
    }
    cudaSetDevice(bk->device);
    if (bk->d_bsk_ntt) cudaFree(bk->d_bsk_ntt);
    if (bk->d_bsk_fft) cudaFree(bk->d_bsk_fft);
    if (bk->d_bsk_fft_x) cudaFree(bk->d_bsk_fft_x);
    if (bk->d_ksk) cudaFree(bk->d_ksk);
    if (bk->d_ksk_t) cudaFree(bk->d_ksk_t);
    delete bk;
}

size_t tfhe_bk_transformed_bytes(const tfhe_bk *bk) {
    if (!bk || !bk->ctx) return 0;
    if (bk->bmmp) return fft_key_bytes(bk->ctx) / bk->ctx->n() * (3 * (bk->ctx->n() / 2));
    return bk->path == TFHE_PATH_FFT ? fft_key_bytes(bk->ctx) : bk->ctx->n() * bk->ctx->ggsw_words() * 2 * 4;
}
int tfhe_bk_get_path(const tfhe_bk *bk) { return bk ? bk->path : TFHE_E_PARAM; }
int tfhe_bk_read_transformed(const tfhe_bk *bk, void *out, size_t bytes) {
    if (!bk || !bk->ctx || !out || bytes != tfhe_bk_transformed_bytes(bk)) return TFHE_E_PARAM;
    tfhe_ctx *ctx = bk->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, bk->path == TFHE_PATH_FFT ? (const void *)bk->d_bsk_fft : (const void *)bk->d_bsk_ntt, bytes, cudaMemcpyDeviceToHost));
    return TFHE_OK;
}

int tfhe_bootstrap_batch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts,
                         const uint32_t *lut_idx, size_t batch, uint32_t *lwe_out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!lwe_in || !luts || !lwe_out || n_luts == 0) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    const size_t io_bytes = batch * (ctx->n() + 1) * 4;
    const void *d_in, *d_luts, *d_idx = nullptr;
    void *d_out;
    if ((rc = stage_in(ctx, lwe_in, io_bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_in(ctx, luts, n_luts * ctx->N() * 4, ctx->luts, &d_luts))) return rc;
    if (lut_idx && (rc = stage_in(ctx, lut_idx, batch * 4, ctx->lutidx, &d_idx))) return rc;
    if ((rc = stage_out(ctx, lwe_out, io_bytes, ctx->out, &d_out))) return rc;
    if ((rc = run_bootstrap(ctx, bk, (const uint32_t *)d_in, (const uint32_t *)d_luts, n_luts, (const uint32_t *)d_idx, batch, (uint32_t *)d_out))) return rc;
    if ((rc = finish_out(ctx, lwe_out, io_bytes, d_out))) return rc;
    return finish_timed(ctx);
}

int tfhe_gates_batch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint8_t *gates, const uint32_t *ct0, const uint32_t *ct1, size_t batch,
                     uint32_t *out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!gates || !ct0 || !ct1 || !out) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    bool any_neg = false;
    std::vector<uint32_t> idx(batch);
    for (size_t b = 0; b < batch; b++) {
        if (gates[b] > TFHE_XNOR) return fail(ctx, TFHE_E_PARAM, "unknown gate opcode");
        idx[b] = gates[b] % 3;
        any_neg |= gates[b] >= 3;
    }
    const size_t N = ctx->N();
    std::vector<uint32_t> tvs(3 * N);
    for (int g = 0; g < 3; g++)
        if ((rc = tfhe_test_vector_boolean(&ctx->p, g, tvs.data() + g * N))) return fail(ctx, rc, "test vector construction failed");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    const size_t io_bytes = batch * (ctx->n() + 1) * 4;
    const void *d0, *d1;
    void *d_out;
    if ((rc = stage_in(ctx, ct0, io_bytes, ctx->in0, &d0))) return rc;
    if ((rc = stage_in(ctx, ct1, io_bytes, ctx->in1, &d1))) return rc;
    if ((rc = stage_out(ctx, out, io_bytes, ctx->out, &d_out))) return rc;
    CU(ctx->in2.ensure(io_bytes));
    CU(ctx->luts.ensure(3 * N * 4));
    CU(ctx->lutidx.ensure(batch * 4));
    CU(ctx->misc.ensure(batch));
    CU(cudaMemcpyAsync(ctx->luts.p, tvs.data(), 3 * N * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lutidx.p, idx.data(), batch * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->misc.p, gates, batch, cudaMemcpyHostToDevice, ctx->stream));
    const size_t len = batch * (ctx->n() + 1);
    gate_linear_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d0, (const uint32_t *)d1, (uint32_t *)ctx->in2.p, len);
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = run_bootstrap(ctx, bk, (const uint32_t *)ctx->in2.p, (const uint32_t *)ctx->luts.p, 3, (const uint32_t *)ctx->lutidx.p, batch, (uint32_t *)d_out))) return rc;
    if (any_neg) {
        uint32_t one = 1u << (ctx->p.log_q - (ctx->p.log_p + ctx->p.padding_bits));
        gate_negate_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((uint32_t *)d_out, (const uint8_t *)ctx->misc.p, 0, (uint32_t)ctx->n(), (uint32_t)batch, one);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    // the pageable staging vectors (tvs, idx) must outlive the async copies: finish_timed synchronises
    if ((rc = finish_out(ctx, out, io_bytes, d_out))) return rc;
    return finish_timed(ctx);
}

int tfhe_gate_batch(tfhe_ctx *ctx, const tfhe_bk *bk, int gate, const uint32_t *ct0, const uint32_t *ct1, size_t batch, uint32_t *out) {
    if (gate < TFHE_AND || gate > TFHE_XNOR) return fail(ctx, TFHE_E_PARAM, "unknown gate opcode");
    std::vector<uint8_t> g(batch, (uint8_t)gate);
    return tfhe_gates_batch(ctx, bk, g.data(), ct0, ct1, batch, out);
}

// ------------------------------------------------------------------ compositions
int tfhe_bootstrap_batch_ks_first(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts,
                                  const uint32_t *lut_idx, size_t batch, uint32_t *lwe_out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!lwe_in || !luts || !lwe_out || n_luts == 0) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    const size_t kN1 = ctx->k() * ctx->N() + 1, io_bytes = batch * kN1 * 4;
    const void *d_in, *d_luts, *d_idx = nullptr;
    void *d_out;
    if ((rc = stage_in(ctx, lwe_in, io_bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_in(ctx, luts, n_luts * ctx->N() * 4, ctx->luts, &d_luts))) return rc;
    if (lut_idx && (rc = stage_in(ctx, lut_idx, batch * 4, ctx->lutidx, &d_idx))) return rc;
    if ((rc = stage_out(ctx, lwe_out, io_bytes, ctx->out, &d_out))) return rc;
    // key_switching.rs:63-103 first: [B][kN+1] -> [B][n+1]
    CU(ctx->in2.ensure(batch * (ctx->n() + 1) * 4));
    CU(ctx->glwe.ensure(batch * ctx->glwe_words() * 4));
    CU(cudaMemsetAsync(ctx->d_err, 0, 4, ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if ((rc = run_key_switch(ctx, bk, (const uint32_t *)d_in, 1, batch, (uint32_t *)ctx->in2.p))) return rc;
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    // bootstrapping.rs:67-105 blind rotation, then :122-156 sample extraction
    PbsArgs a = {};
    a.bsk_ntt = bk->d_bsk_ntt;
    a.tw[0] = ctx->tw[0]; a.tw[1] = ctx->tw[1];
    a.prime[0] = ctx->prime[0]; a.prime[1] = ctx->prime[1];
    a.lwe_in = (const uint32_t *)ctx->in2.p; a.luts = (const uint32_t *)d_luts; a.lut_idx = (const uint32_t *)d_idx;
    a.glwe_out = (uint32_t *)ctx->glwe.p;
    a.err_flag = ctx->d_err;
    a.n = (uint32_t)ctx->n(); a.batch = (uint32_t)batch; a.mode = 0;
    a.log_p = ctx->p.log_p;
    a.enc_shift = ctx->p.log_q - (ctx->p.log_p + ctx->p.padding_bits);
    a.n_luts = (uint32_t)n_luts;
    if ((rc = launch_pbs(ctx, a, bk))) return rc;
    const size_t total = batch * kN1;
    sample_extract_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)ctx->glwe.p, (uint32_t *)d_out, (uint32_t)ctx->k(),
                                                                                     (int)ctx->p.glwe_poly_degree, (uint32_t)batch);
    CU(cudaGetLastError());
    ctx->launches++;
    // timing slots: [0] = blind rotation + extraction, [1] = key switch (ev[1] closes the rotation interval)
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if ((rc = finish_out(ctx, lwe_out, io_bytes, d_out))) return rc;
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float t = 0;
    CU(cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[1])); ctx->last_ms[0] = t;
    CU(cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3])); ctx->last_ms[1] = t;
    CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[4])); ctx->last_ms[2] = t;
    uint32_t flag = 0;
    CU(cudaMemcpy(&flag, ctx->d_err, 4, cudaMemcpyDeviceToHost));
    return decode_err_flag(ctx, flag);
}

int tfhe_gate_k_batch(tfhe_ctx *ctx, const tfhe_bk *bk, uint32_t k_inputs, uint32_t truth_table, const uint32_t *const *cts, size_t batch,
                      uint32_t *out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!cts || !out) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (k_inputs < 2 || k_inputs > ctx->p.log_p) return fail(ctx, TFHE_E_PARAM, "k-input gate needs 2 <= k <= log_p");
    for (uint32_t i = 0; i < k_inputs; i++)
        if (!cts[i]) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    const uint32_t pm = 1u << ctx->p.log_p, nf = 1u << k_inputs;
    const bool negate = truth_table & 1u;   // f(0) = 1: bootstrap 1 - f, then trivial(1) - result
    std::vector<uint32_t> lut(pm, 0u);
    for (uint32_t j = 0; j < nf; j++) lut[j] = (((truth_table >> j) & 1u) ^ (negate ? 1u : 0u));
    const size_t N = ctx->N();
    std::vector<uint32_t> tv(N);
    if ((rc = tfhe_test_vector_from_lut(&ctx->p, lut.data(), lut.size(), tv.data()))) return fail(ctx, rc, "test vector construction failed");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    const size_t len = batch * (ctx->n() + 1), io_bytes = len * 4;
    void *d_out;
    if ((rc = stage_out(ctx, out, io_bytes, ctx->out, &d_out))) return rc;
    CU(ctx->in2.ensure(io_bytes));
    CU(ctx->luts.ensure(N * 4));
    CU(cudaMemcpyAsync(ctx->luts.p, tv.data(), N * 4, cudaMemcpyHostToDevice, ctx->stream));
    // Horner: acc = c_{k-1}; acc = 2*acc + c_i  (each step is boolean.rs:18's `2*ct1 + ct0`)
    const void *d_acc;
    if ((rc = stage_in(ctx, cts[k_inputs - 1], io_bytes, ctx->in1, &d_acc))) return rc;
    for (int i = (int)k_inputs - 2; i >= 0; i--) {
        const void *d_ci;
        if ((rc = stage_in(ctx, cts[i], io_bytes, ctx->in0, &d_ci))) return rc;
        gate_linear_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d_ci, (const uint32_t *)d_acc, (uint32_t *)ctx->in2.p, len);
        CU(cudaGetLastError());
        ctx->launches++;
        d_acc = ctx->in2.p;   // in place from the second step on: element-wise, each thread reads before it writes
    }
    if ((rc = run_bootstrap(ctx, bk, (const uint32_t *)d_acc, (const uint32_t *)ctx->luts.p, 1, nullptr, batch, (uint32_t *)d_out))) return rc;
    if (negate) {
        const uint32_t one = 1u << (ctx->p.log_q - (ctx->p.log_p + ctx->p.padding_bits));
        gate_negate_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((uint32_t *)d_out, nullptr, 1, (uint32_t)ctx->n(), (uint32_t)batch, one);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if ((rc = finish_out(ctx, out, io_bytes, d_out))) return rc;
    return finish_timed(ctx);   // synchronises: the pageable tv vector outlives its async copy
}

// ------------------------------------------------------------------ sub-operations
int tfhe_switch_modulus(tfhe_ctx *ctx, const uint32_t *values, size_t len, uint32_t *out) {
    if (!ctx || !values || !out) return TFHE_E_PARAM;
    if (len == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const void *d_in; void *d_out; int rc;
    if ((rc = stage_in(ctx, values, len * 4, ctx->in0, &d_in))) return rc;
    if ((rc = stage_out(ctx, out, len * 4, ctx->out, &d_out))) return rc;
    switch_modulus_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d_in, (uint32_t *)d_out, len, (int)ctx->p.glwe_poly_degree);
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = finish_out(ctx, out, len * 4, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_decompose(tfhe_ctx *ctx, int which, const uint32_t *values, size_t len, uint32_t *out) {
    if (!ctx || !values || !out || which < 0 || which > 1) return TFHE_E_PARAM;
    if (len == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const uint32_t lb = which ? ctx->p.ks_log_base : ctx->p.pbs_log_base, lv = which ? ctx->p.ks_levels : ctx->p.pbs_levels;
    const void *d_in; void *d_out; int rc;
    if ((rc = stage_in(ctx, values, len * 4, ctx->in0, &d_in))) return rc;
    if ((rc = stage_out(ctx, out, len * lv * 4, ctx->out, &d_out))) return rc;
    const unsigned blocks = (unsigned)((len + 255) / 256);
    const uint32_t *i = (const uint32_t *)d_in; uint32_t *o = (uint32_t *)d_out;
    if (lb == 4 && lv == 6) decompose_kernel<4, 6><<<blocks, 256, 0, ctx->stream>>>(i, o, len);
    else if (lb == 4 && lv == 5) decompose_kernel<4, 5><<<blocks, 256, 0, ctx->stream>>>(i, o, len);
    else if (lb == 8 && lv == 3) decompose_kernel<8, 3><<<blocks, 256, 0, ctx->stream>>>(i, o, len);
    else if (lb == 2 && lv == 8) decompose_kernel<2, 8><<<blocks, 256, 0, ctx->stream>>>(i, o, len);
    else return fail(ctx, TFHE_E_PARAM, "decomposer not instantiated");
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = finish_out(ctx, out, len * lv * 4, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_glwe_mul_monomial(tfhe_ctx *ctx, const uint32_t *glwe, const int64_t *index, size_t batch, uint32_t *out) {
    if (!ctx || !glwe || !index || !out) return TFHE_E_PARAM;
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = batch * ctx->glwe_words() * 4;
    std::vector<uint32_t> rot(batch);
    for (size_t b = 0; b < batch; b++) rot[b] = (uint32_t)(((uint64_t)index[b]) % (uint64_t)(2 * ctx->N()));  // utils.rs:186
    const void *d_in; void *d_out; int rc;
    if ((rc = stage_in(ctx, glwe, bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_out(ctx, out, bytes, ctx->out, &d_out))) return rc;
    CU(ctx->lutidx.ensure(batch * 4));
    CU(cudaMemcpyAsync(ctx->lutidx.p, rot.data(), batch * 4, cudaMemcpyHostToDevice, ctx->stream));
    const size_t total = batch * ctx->glwe_words();
    mul_monomial_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d_in, (const uint32_t *)ctx->lutidx.p, (uint32_t *)d_out,
                                                                                   (uint32_t)(ctx->k() + 1), (int)ctx->p.glwe_poly_degree, total);
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = finish_out(ctx, out, bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

static int ext_or_cmux(tfhe_ctx *ctx, const tfhe_bk *bk, int mode, const uint32_t *ggsw_index, const uint32_t *in0, const uint32_t *in1,
                       size_t batch, uint32_t *out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!ggsw_index || !in0 || (mode == 2 && !in1) || !out) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    for (size_t b = 0; b < batch; b++)
        if (ggsw_index[b] >= ctx->n()) return fail(ctx, TFHE_E_PARAM, "ggsw_index out of range");
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = batch * ctx->glwe_words() * 4;
    const void *d0, *d1 = nullptr; void *d_out;
    if ((rc = stage_in(ctx, in0, bytes, ctx->in0, &d0))) return rc;
    if (mode == 2 && (rc = stage_in(ctx, in1, bytes, ctx->in1, &d1))) return rc;
    if ((rc = stage_out(ctx, out, bytes, ctx->out, &d_out))) return rc;
    CU(ctx->lutidx.ensure(batch * 4));
    CU(cudaMemcpyAsync(ctx->lutidx.p, ggsw_index, batch * 4, cudaMemcpyHostToDevice, ctx->stream));
    PbsArgs a = {};
    a.bsk_ntt = bk->d_bsk_ntt;
    a.tw[0] = ctx->tw[0]; a.tw[1] = ctx->tw[1];
    a.prime[0] = ctx->prime[0]; a.prime[1] = ctx->prime[1];
    a.in0 = (const uint32_t *)d0; a.in1 = (const uint32_t *)(mode == 2 ? d1 : d0);
    a.ggsw_index = (const uint32_t *)ctx->lutidx.p;
    a.glwe_out = (uint32_t *)d_out;
    a.err_flag = ctx->d_err;
    a.n = (uint32_t)ctx->n(); a.batch = (uint32_t)batch; a.mode = (uint32_t)mode;
    if ((rc = launch_pbs(ctx, a, bk))) return rc;
    if ((rc = finish_out(ctx, out, bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}
int tfhe_external_product(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *ggsw_index, const uint32_t *glwe, size_t batch, uint32_t *out) {
    return ext_or_cmux(ctx, bk, 1, ggsw_index, glwe, nullptr, batch, out);
}
int tfhe_cmux(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *ggsw_index, const uint32_t *ct0, const uint32_t *ct1, size_t batch, uint32_t *out) {
    return ext_or_cmux(ctx, bk, 2, ggsw_index, ct0, ct1, batch, out);
}

int tfhe_blind_rotate(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts, const uint32_t *lut_idx,
                      size_t batch, uint32_t *glwe_out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!lwe_in || !luts || !glwe_out || n_luts == 0) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t in_bytes = batch * (ctx->n() + 1) * 4, out_bytes = batch * ctx->glwe_words() * 4;
    const void *d_in, *d_luts, *d_idx = nullptr; void *d_out;
    if ((rc = stage_in(ctx, lwe_in, in_bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_in(ctx, luts, n_luts * ctx->N() * 4, ctx->luts, &d_luts))) return rc;
    if (lut_idx && (rc = stage_in(ctx, lut_idx, batch * 4, ctx->lutidx, &d_idx))) return rc;
    if ((rc = stage_out(ctx, glwe_out, out_bytes, ctx->out, &d_out))) return rc;
    CU(cudaMemsetAsync(ctx->d_err, 0, 4, ctx->stream));
    PbsArgs a = {};
    a.bsk_ntt = bk->d_bsk_ntt;
    a.tw[0] = ctx->tw[0]; a.tw[1] = ctx->tw[1];
    a.prime[0] = ctx->prime[0]; a.prime[1] = ctx->prime[1];
    a.lwe_in = (const uint32_t *)d_in; a.luts = (const uint32_t *)d_luts; a.lut_idx = (const uint32_t *)d_idx;
    a.glwe_out = (uint32_t *)d_out;
    a.err_flag = ctx->d_err;
    a.n = (uint32_t)ctx->n(); a.batch = (uint32_t)batch; a.mode = 0;
    a.log_p = ctx->p.log_p;
    a.enc_shift = ctx->p.log_q - (ctx->p.log_p + ctx->p.padding_bits);
    a.n_luts = (uint32_t)n_luts;
    if ((rc = launch_pbs(ctx, a, bk))) return rc;
    if ((rc = finish_out(ctx, glwe_out, out_bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    uint32_t flag = 0;
    CU(cudaMemcpy(&flag, ctx->d_err, 4, cudaMemcpyDeviceToHost));
    return decode_err_flag(ctx, flag);
}

int tfhe_sample_extract(tfhe_ctx *ctx, const uint32_t *glwe, size_t batch, uint32_t *lwe_out) {
    if (!ctx || !glwe || !lwe_out) return TFHE_E_PARAM;
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t in_bytes = batch * ctx->glwe_words() * 4, kN1 = ctx->k() * ctx->N() + 1, out_bytes = batch * kN1 * 4;
    const void *d_in; void *d_out; int rc;
    if ((rc = stage_in(ctx, glwe, in_bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_out(ctx, lwe_out, out_bytes, ctx->out, &d_out))) return rc;
    const size_t total = batch * kN1;
    sample_extract_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d_in, (uint32_t *)d_out, (uint32_t)ctx->k(),
                                                                                     (int)ctx->p.glwe_poly_degree, (uint32_t)batch);
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = finish_out(ctx, lwe_out, out_bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_key_switch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, size_t batch, uint32_t *lwe_out) {
    int rc = check_bk(ctx, bk);
    if (rc) return rc;
    if (!lwe_in || !lwe_out) return fail(ctx, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t in_bytes = batch * (ctx->k() * ctx->N() + 1) * 4, out_bytes = batch * (ctx->n() + 1) * 4;
    const void *d_in; void *d_out;
    if ((rc = stage_in(ctx, lwe_in, in_bytes, ctx->in0, &d_in))) return rc;
    if ((rc = stage_out(ctx, lwe_out, out_bytes, ctx->out, &d_out))) return rc;
    if ((rc = run_key_switch(ctx, bk, (const uint32_t *)d_in, 1, batch, (uint32_t *)d_out))) return rc;
    if ((rc = finish_out(ctx, lwe_out, out_bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_negacyclic_mul(tfhe_ctx *ctx, const int32_t *a, const uint32_t *g, size_t batch, uint32_t *out) {
    if (!ctx || !a || !g || !out) return TFHE_E_PARAM;
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = batch * ctx->N() * 4;
    // range check of the small operand (on a host copy: this is a parity-test entry point, not a hot path)
    std::vector<int32_t> ha(batch * ctx->N());
    CU(cudaMemcpy(ha.data(), a, bytes, cudaMemcpyDefault));
    bool small = true;
    for (int32_t v : ha)
        if (v > 1024 || v < -1024) { small = false; break; }
    const void *d_a, *d_g; void *d_out; int rc;
    if ((rc = stage_in(ctx, a, bytes, ctx->in0, &d_a))) return rc;
    if ((rc = stage_in(ctx, g, bytes, ctx->in1, &d_g))) return rc;
    if ((rc = stage_out(ctx, out, bytes, ctx->out, &d_out))) return rc;
    auto mul = [&](const int32_t *aa, uint32_t *oo) {
        switch (ctx->pbs_id) {
        case 0: return launch_polymul_t<K0>(ctx, aa, (const uint32_t *)d_g, oo, batch);
        case 1: return launch_polymul_t<K1>(ctx, aa, (const uint32_t *)d_g, oo, batch);
        case 2: return launch_polymul_t<K2>(ctx, aa, (const uint32_t *)d_g, oo, batch);
        }
        return fail(ctx, TFHE_E_PARAM, "no kernel instantiation");
    };
    if (small) {
        rc = mul((const int32_t *)d_a, (uint32_t *)d_out);
    } else {
        // any u32 operand (the reference's poly_mul takes u32 x u32): a = sum_i 2^(8 i) byte_i(a) as a word mod 2^32, so
        // a (*) g = sum_i (byte_i(a) (*) g) << 8 i  mod 2^32 -- four products with an operand inside the exact range
        const size_t len = batch * ctx->N();
        CU(ctx->misc.ensure(2 * bytes));
        int32_t *d_limb = (int32_t *)ctx->misc.p;
        uint32_t *d_part = (uint32_t *)((uint8_t *)ctx->misc.p + bytes);
        const unsigned blocks = (unsigned)((len + 255) / 256);
        rc = TFHE_OK;
        for (int limb = 0; limb < 4 && rc == TFHE_OK; limb++) {
            byte_limb_kernel<<<blocks, 256, 0, ctx->stream>>>((const int32_t *)d_a, d_limb, limb, len);
            CU(cudaGetLastError());
            rc = mul(d_limb, d_part);
            if (rc != TFHE_OK) break;
            shl_accumulate_kernel<<<blocks, 256, 0, ctx->stream>>>((uint32_t *)d_out, d_part, 8 * limb, limb == 0, len);
            CU(cudaGetLastError());
            ctx->launches += 2;
        }
    }
    if (rc) return rc;
    if ((rc = finish_out(ctx, out, bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_gate_linear(tfhe_ctx *ctx, const uint32_t *ct0, const uint32_t *ct1, size_t batch, uint32_t *out) {
    if (!ctx || !ct0 || !ct1 || !out) return TFHE_E_PARAM;
    if (batch == 0) return TFHE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t len = batch * (ctx->n() + 1), bytes = len * 4;
    const void *d0, *d1; void *d_out; int rc;
    if ((rc = stage_in(ctx, ct0, bytes, ctx->in0, &d0))) return rc;
    if ((rc = stage_in(ctx, ct1, bytes, ctx->in1, &d1))) return rc;
    if ((rc = stage_out(ctx, out, bytes, ctx->out, &d_out))) return rc;
    gate_linear_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)d0, (const uint32_t *)d1, (uint32_t *)d_out, len);
    CU(cudaGetLastError());
    ctx->launches++;
    if ((rc = finish_out(ctx, out, bytes, d_out))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return TFHE_OK;
}

int tfhe_measure_int_peak(tfhe_ctx *ctx, double out[8]) {
    if (!ctx || !out) return TFHE_E_PARAM;
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    uint32_t *sink = nullptr;
    CU(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int kind = 0; kind < 6; kind++) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, ctx->stream));
            if (kind == 0) int_peak_kernel<0><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            else if (kind == 1) int_peak_kernel<1><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            else if (kind == 2) int_peak_kernel<2><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            else if (kind == 3) int_peak_kernel<3><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            else if (kind == 4) int_peak_kernel<4><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            else int_peak_kernel<5><<<blocks, 256, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, iters);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double ops = (double)blocks * 256.0 * iters * (kind >= 3 ? 32.0 : 64.0);
            const double rate = ops / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;  // first repetition is warm-up
            ctx->launches++;
        }
        out[kind] = best;
    }
    // FMA-bound butterfly stream at the blind-rotation kernel's residency (128-thread CTAs, 3 and 8 per SM)
    for (int cfg = 0; cfg < 2; cfg++) {
        const int per_sm = cfg == 0 ? 3 : 8, nblocks = prop.multiProcessorCount * per_sm, it2 = 8192;
        double best = 0;
        for (int rep = 0; rep < 3; rep++) {
            CU(cudaEventRecord(e0, ctx->stream));
            bfly_stream_kernel<<<nblocks, 128, 0, ctx->stream>>>(sink, 0x9E3779B1u, 12345u, it2);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double rate = (double)nblocks * 128.0 * it2 * 32.0 / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;
            ctx->launches++;
        }
        out[6 + cfg] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return TFHE_OK;
}

int tfhe_measure_fp64_peak(tfhe_ctx *ctx, double out[4]) {
    if (!ctx || !out) return TFHE_E_PARAM;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    double *sink = nullptr;
    CU(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int kind = 0; kind < 3; kind++) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, ctx->stream));
            if (kind == 0) fft::fp64_peak_kernel<0><<<blocks, 256, 0, ctx->stream>>>(sink, 1.0000001, 1e-9, iters);
            else if (kind == 1) fft::fp64_peak_kernel<1><<<blocks, 256, 0, ctx->stream>>>(sink, 1.0000001, 1e-9, iters);
            else fft::fp64_peak_kernel<2><<<blocks, 256, 0, ctx->stream>>>(sink, 1.0000001, 1e-9, iters);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            // kinds 0/1: 32 DFMA per thread per iteration; kind 2: 16 butterflies (6 DFMA each) per thread per iteration
            const double units = (double)blocks * 256.0 * iters * (kind == 2 ? 16.0 : 32.0);
            const double rate = units / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;
            ctx->launches++;
        }
        out[kind] = best;
    }
    out[3] = 0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return TFHE_OK;
}

int tfhe_last_timing(const tfhe_ctx *ctx, double out[3]) {
    if (!ctx || !out) return TFHE_E_PARAM;
    for (int i = 0; i < 3; i++) out[i] = ctx->last_ms[i];
    return TFHE_OK;
}

}  // extern "C"
