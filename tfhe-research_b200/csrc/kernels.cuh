// kernels.cuh -- sm_100a kernels of the PBS path.  The blind rotation is modular integer arithmetic on the CUDA cores
// (IMAD / IMAD.WIDE on the FMA pipe, IADD3/LOP3/VIADDMNMX on the ALU pipe; the FP64-FFT form is in kernels_fft.cuh);
// the key-switching product -- a genuine integer matrix product -- also has an exact integer-tensor-core form.
//
//   K2 pbs_kernel           blind rotation = n fused CMUX steps per ciphertext, accumulator resident
//                           in shared memory (bootstrapping.rs:58-105, ggsw.rs:132-178); GGSW rows are
//                           streamed global -> shared with cp.async.bulk (TMA) + mbarrier; also runs a
//                           single external product / CMUX for the sub-operation entry points
//   K0 bsk_transform_kernel raw BSK -> 2-prime NTT domain, pre-scaled by N^-1 (one-off at upload)
//   K3+K4 ks_digits_kernel, then ks_mma_kernel (default; s8 x u8 IMMA over the key's byte planes, exact) /
//                           ks_gemm_kernel (IMAD) / ks_gemv_kernel (batches <= 8): sample extract + KS decomposition,
//                           then the wrapping u32 accumulation against the KSK (bootstrapping.rs:122-156,
//                           key_switching.rs:63-103)
//   K5 small element-wise kernels (gate linear part lwe.rs:9-23, NAND-style negation, sub-ops)
#pragma once
#include <cuda_runtime.h>

#ifndef TFHE_MERGE_ROWS
#define TFHE_MERGE_ROWS 0
#endif

#include "pbs_team.cuh"

namespace tfhe {

// ------------------------------------------------------------------------------------------ K2
struct PbsArgs {
    PrimeTab prime[2];         // per-prime constants + pass-A twiddles (constant bank, uniform LDC)
    const uint32_t *bsk_ntt;   // [n][2][ROWS][P][N] slot order, residues < q, pre-scaled by N^-1
    TwTables tw[2];
    // mode 0 (blind rotate)
    const uint32_t *lwe_in;    // [B][n+1]
    const uint32_t *luts;      // [T][N] unencoded
    const uint32_t *lut_idx;   // [B] or nullptr
    // mode 1 (external product: out = ExtProd(G, in0)) / mode 2 (cmux: out = ExtProd(G, in1-in0)+in0)
    const uint32_t *in0, *in1; // [B][P][N]
    const uint32_t *ggsw_index;  // [B]
    uint32_t *glwe_out;        // [B][P][N]
    uint32_t *err_flag;        // bit 0: a test-vector entry >= 2^log_p (glwe.rs:144); bit 1: TMA wait timed out; bit 2: lut_idx out of range
    uint32_t n, batch, mode, log_p, enc_shift;
    uint32_t n_luts;           // lut_idx[b] >= n_luts sets err_flag bit 2 and selects test vector 0
    uint32_t skew_ns, skew_div, skew_mod;  // start-up stagger of co-resident CTAs (see pbs_kernel)
};

__device__ __forceinline__ void team_bar(int pr, int nthreads) {
    // literal barrier ids so that ptxas reserves 3 named barriers per CTA instead of all 16
    if (pr == 0) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Bounded wait (a hung kernel would cost a GPU strike).  The bound is wall-clock (%globaltimer, 10 s), not a spin count, so a
// debugger or heavy time-slicing cannot trip it.  On expiry the waiter sets bit 1 of err_flag and RETURNS: every other
// waiter of the grid sees the bit at its next poll and returns too, so the kernel runs to completion with invalid data
// (all addresses are data independent) and the host reports TFHE_E_CUDA -- no trap, no sticky context error.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool mbar_try_once(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *err_flag) {
    if (mbar_try_once(bar, parity)) return;   // the common case
    unsigned long long t0 = 0;
    for (uint32_t spin = 1; !mbar_try_once(bar, parity); spin++) {
        if ((spin & 255u) == 0u) {
            if (*(volatile uint32_t *)err_flag & 2u) return;                      // another waiter timed out: abandon
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) { atomicOr(err_flag, 2u); return; }
        }
    }
}

template <class K>
__device__ __forceinline__ void team_cmux(TeamRegs<K> &R, uint32_t pr, uint32_t t, uint32_t jbB, const PrimeTab &pt, const TwTables &tw,
                                          const uint8_t *dig, uint32_t *buf0, uint32_t *buf1, const uint32_t *g, uint32_t *gbuf,
                                          uint64_t *bar, uint32_t &parity, uint32_t *err_flag) {
    team_zero_acc<K>(R);
#if TFHE_MERGE_ROWS
    // Software pipelining across rows: the last pass + multiply-accumulate of row r share one barrier interval
    // with the digit load + first pass of row r+1, so the scheduler has two independent instruction streams.
    phase_F1<K>(R, t, jbB, pt, tw, dig, 0, buf0);
    team_bar(pr, K::T);
#pragma unroll 1
    for (int r = 0; r < K::ROWS; r++) {
        if constexpr (K::STAGE_G) {
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, K::G_ROW_BYTES);
                bulk_g2s(gbuf, g + (size_t)r * K::P * K::N, K::G_ROW_BYTES, bar);
            }
        }
        phase_F2<K>(R, jbB, pt, buf0, buf1);
        team_bar(pr, K::T);
        phase_F3a<K>(R, t, pt, tw, buf1);
        if (r + 1 < K::ROWS) {
            uint32_t y[K::E];
            phase_F1x<K>(y, t, pt, dig, r + 1, buf0);   // buf0 was last read before the previous barrier
        }
        if constexpr (K::STAGE_G) {
            mbar_wait(bar, parity, err_flag);
            parity ^= 1u;
            phase_F3b<K, true>(R, t, gbuf);
        } else {
            phase_F3b<K, false>(R, t, g + (size_t)r * K::P * K::N);
        }
        team_bar(pr, K::T);  // buf0 (row r+1) complete; gbuf and buf1 free again
    }
#else
#pragma unroll 1
    for (int r = 0; r < K::ROWS; r++) {
        phase_F1<K>(R, t, jbB, pt, tw, dig, r, buf0);
        team_bar(pr, K::T);  // also: every thread of the team is done reading gbuf (row r-1)
        if constexpr (K::STAGE_G) {
            if (t == 0) {
                // the staging buffer was last touched through the generic proxy (LDS of row r-1, or the residue
                // stores/loads that alias it): order those before the async-proxy write of the bulk copy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, K::G_ROW_BYTES);
                bulk_g2s(gbuf, g + (size_t)r * K::P * K::N, K::G_ROW_BYTES, bar);
            }
        }
        phase_F2<K>(R, jbB, pt, buf0, buf1);
        team_bar(pr, K::T);
        phase_F3a<K>(R, t, pt, tw, buf1);
        if constexpr (K::STAGE_G) {
            mbar_wait(bar, parity, err_flag);
            parity ^= 1u;
            phase_F3b<K, true>(R, t, gbuf);
        } else {
            phase_F3b<K, false>(R, t, g + (size_t)r * K::P * K::N);
        }
    }
#endif
}
template <class K>
__device__ __forceinline__ void team_inverse(TeamRegs<K> &R, uint32_t pr, uint32_t t, uint32_t jbB, const PrimeTab &pt, const TwTables &tw,
                                             uint32_t *buf0, uint32_t *buf1, uint32_t *res_pr) {
#pragma unroll 1
    for (int c = 0; c < K::P; c++) {
        phase_I1<K>(R, t, jbB, c, pt, tw, buf0);
        team_bar(pr, K::T);
        phase_I2<K>(R, jbB, pt, buf0, buf1);
        team_bar(pr, K::T);
        phase_I3<K>(R, t, pt, buf1, res_pr + c * K::N);
    }
}

template <class K, int MINB>
__global__ void __launch_bounds__(K::THREADS, MINB) pbs_kernel(const __grid_constant__ PbsArgs a) {
    using C = typename K::Ntt;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem + K::SM_ACC);
    uint8_t *dig = smem + K::SM_DIG;
    uint32_t *res = reinterpret_cast<uint32_t *>(smem + K::SM_RES);  // aliases the G rows (or dig), see PbsCfg
    uint32_t *buf = reinterpret_cast<uint32_t *>(smem + K::SM_BUF);
    uint16_t *at = reinterpret_cast<uint16_t *>(smem + K::SM_AT);

    const uint32_t tid = threadIdx.x;
    const uint32_t pr = tid / K::T, t = tid % K::T;  // warp-uniform: T is a multiple of 32
    const uint32_t ct = blockIdx.x;
    const uint32_t jbB = jbase_B<C>(t);
    uint32_t *buf0 = buf + (pr * 2 + 0) * C::NPAD, *buf1 = buf + (pr * 2 + 1) * C::NPAD;
    uint32_t *res_pr = res + pr * K::P * K::N;
    uint32_t *gbuf = reinterpret_cast<uint32_t *>(smem + K::SM_G + pr * K::G_ROW_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + K::SM_BAR) + pr;
    uint32_t parity = 0;
    constexpr size_t GGSW_WORDS = (size_t)2 * K::ROWS * K::P * K::N;
    const PrimeTab &pt = a.prime[pr];
    const TwTables &tw = a.tw[pr];

    if constexpr (K::STAGE_G) {
        if (t == 0) mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    TeamRegs<K> R;
    team_init<K>(R, tw, t);

    // Co-resident CTAs run identical phase sequences; started together they stay phase-locked and hit the
    // FMA pipe in the same bursts.  An initial stagger of a fraction of a CMUX step de-phases them.
    if (a.skew_ns) {
        const uint32_t slot = (blockIdx.x / a.skew_div) % a.skew_mod;
        for (uint32_t s = 0; s < slot; s++) __nanosleep(a.skew_ns);
    }

    uint32_t n_steps;
    if (a.mode == 0) {
        // utils.rs:23-33 mod switch of (a_0..a_{n-1}, b) to 2N
        const uint32_t *lwe = a.lwe_in + (size_t)ct * (a.n + 1);
        for (uint32_t i = tid; i <= a.n; i += K::THREADS) at[i] = (uint16_t)mod_switch(__ldg(lwe + i), K::LOGN);
        __syncthreads();
        // acc = trivial GLWE of the encoded test vector times X^{-b~}  (bootstrapping.rs:79-86)
        const uint32_t b = at[a.n];
        uint32_t li = a.lut_idx ? __ldg(a.lut_idx + ct) : 0u;
        if (li >= a.n_luts) {   // never read a test vector out of bounds: flag it (host returns TFHE_E_PARAM) and fall back to 0
            atomicOr(a.err_flag, 4u);
            li = 0u;
        }
        const uint32_t *lut = a.luts + (size_t)li * K::N;
        for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::THREADS) {
            const uint32_t p = idx >> K::LOGN, j = idx & (K::N - 1u);
            uint32_t v = 0;
            if (p == (uint32_t)K::K) {
                const uint32_t src = (j + b) & (2u * K::N - 1u);
                const uint32_t m = __ldg(lut + (src & (K::N - 1u)));
                if (m >> a.log_p) atomicOr(a.err_flag, 1u);
                v = m << a.enc_shift;
                if (src & K::N) v = 0u - v;
            }
            acc[idx] = v;
        }
        n_steps = a.n;
    } else {
        const uint32_t *base = a.in0 + (size_t)ct * K::P * K::N;
        for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::THREADS) acc[idx] = (a.mode == 2) ? __ldg(base + idx) : 0u;
        n_steps = 1;
    }
    __syncthreads();

#pragma unroll 1
    for (uint32_t i = 0; i < n_steps; i++) {
        uint32_t gi = i;
        if (a.mode == 0) {
            const uint32_t rot = at[i];
            if (rot == 0) continue;  // diff == 0 => external product == 0 exactly (CTA-uniform)
            phase_digits<K>(tid, dig, [&](uint32_t p, uint32_t j) { return rot_coeff(acc + p * K::N, j, rot, K::LOGN) - acc[p * K::N + j]; });
        } else {
            gi = __ldg(a.ggsw_index + ct);
            const uint32_t *x0 = a.in0 + (size_t)ct * K::P * K::N, *x1 = a.in1 + (size_t)ct * K::P * K::N;
            if (a.mode == 1) phase_digits<K>(tid, dig, [&](uint32_t p, uint32_t j) { return __ldg(x0 + p * K::N + j); });
            else phase_digits<K>(tid, dig, [&](uint32_t p, uint32_t j) { return __ldg(x1 + p * K::N + j) - __ldg(x0 + p * K::N + j); });
        }
        __syncthreads();
        const uint32_t *g = a.bsk_ntt + (size_t)gi * GGSW_WORDS + (size_t)pr * (K::ROWS * K::P * K::N);
        team_cmux<K>(R, pr, t, jbB, pt, tw, dig, buf0, buf1, g, gbuf, bar, parity, a.err_flag);
        if constexpr (!K::STAGE_G) __syncthreads();  // both teams are done reading dig before res (same bytes) is written
        team_inverse<K>(R, pr, t, jbB, pt, tw, buf0, buf1, res_pr);
        __syncthreads();
        phase_crt<K>(tid, res, acc);
        __syncthreads();
    }
    uint32_t *out = a.glwe_out + (size_t)ct * K::P * K::N;
    for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::THREADS) out[idx] = acc[idx];
}

// ------------------------------------------------------------------------------------------ K0
// grid = number of polynomials (n*ROWS*P); block = 2T (one team per prime); in natural [n][ROWS][P][N],
// out [n][2][ROWS][P][N].
struct TransformArgs {
    PrimeTab prime[2];
    TwTables tw[2];
    const uint32_t *raw;
    uint32_t *out;
};
template <class K>
__global__ void __launch_bounds__(K::THREADS) bsk_transform_kernel(const __grid_constant__ TransformArgs a) {
    using C = typename K::Ntt;
    __shared__ __align__(16) uint32_t buf[2 * 2 * C::NPAD];
    const uint32_t tid = threadIdx.x, pr = tid / K::T, t = tid % K::T;
    const size_t poly = blockIdx.x;  // = (i*ROWS + r)*P + c
    const size_t i = poly / (K::ROWS * K::P), rc = poly % (K::ROWS * K::P);
    const uint32_t *g = a.raw + poly * K::N;
    uint32_t *o = a.out + ((i * 2 + pr) * (K::ROWS * K::P) + rc) * K::N;
    uint32_t *buf0 = buf + (pr * 2) * C::NPAD, *buf1 = buf0 + C::NPAD;
    const uint32_t jbB = jbase_B<C>(t);
    const PrimeTab &pt = a.prime[pr];
    const TwTables &tw = a.tw[pr];
    TeamRegs<K> R;
    team_init<K>(R, tw, t);
    phase_T1<K>(R, t, jbB, pt, tw, g, buf0);
    team_bar(pr, K::T);
    phase_F2<K>(R, jbB, pt, buf0, buf1);
    team_bar(pr, K::T);
    phase_T3<K>(R, t, pt, tw, buf1, o);
}

// ------------------------------------------------------------------------------------------ K3+K4
// Sample extraction (bootstrapping.rs:122-156, index 0) fused with the KS decomposition
// (key_switching.rs:71-78): glwe [B][P][N] -> digits int8 [B][kN*L] (+ body b' per ciphertext).
template <int LOGB, int L>
__global__ void ks_digits_kernel(const uint32_t *__restrict__ glwe, int8_t *__restrict__ digits, uint32_t *__restrict__ body,
                                 uint32_t k, uint32_t logn, uint32_t batch, int from_lwe) {
    const uint32_t N = 1u << logn, kN = k * N;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * kN) return;
    const uint32_t b = (uint32_t)(gid / kN), idx = (uint32_t)(gid % kN);
    uint32_t v;
    if (from_lwe) {  // input already is an extracted LWE [B][kN+1]
        v = glwe[(size_t)b * (kN + 1) + idx];
        if (idx == 0) body[b] = glwe[(size_t)b * (kN + 1) + kN];
    } else {
        const uint32_t p = idx >> logn, tt = idx & (N - 1u);
        const uint32_t *poly = glwe + ((size_t)b * (k + 1) + p) * N;
        v = tt == 0 ? poly[0] : 0u - poly[N - tt];
        if (idx == 0) body[b] = glwe[((size_t)b * (k + 1) + k) * N];
    }
    int32_t d[L];
    decompose_signed<LOGB, L>(v, d);
    int8_t *o = digits + (size_t)b * kN * L + (size_t)idx * L;
#pragma unroll
    for (int lev = 0; lev < L; lev++) o[lev] = (int8_t)d[lev];
}

// out[b][c] = -(sum_r D[b][r] * KSK[r][c]) (+ body[b] at c == n), wrapping u32.
// CTA tile: 64 ciphertexts x 128 columns, 256 threads, 8x4 register micro-tile, BK = 32.  The device copy of the KSK has
// a row stride that is a multiple of 128 words (zero padded), so tiles are read with aligned 128-bit loads and need no
// column checks; the next tile's global loads are issued into registers before the current tile is multiplied.
constexpr int KS_BM = 64, KS_BN = 128, KS_BK = 32, KS_THREADS = 256;
__global__ void __launch_bounds__(KS_THREADS) ks_gemm_kernel(const int8_t *__restrict__ digits, const uint32_t *__restrict__ ksk,
                                                             const uint32_t *__restrict__ body, uint32_t *__restrict__ out,
                                                             uint32_t KD, uint32_t n, uint32_t batch, uint32_t stride) {
    __shared__ __align__(16) int32_t sD[KS_BK][KS_BM];
    __shared__ __align__(16) uint32_t sK[KS_BK][KS_BN];
    const uint32_t tid = threadIdx.x;
    const uint32_t c0 = blockIdx.x * KS_BN, b0 = blockIdx.y * KS_BM;
    const uint32_t tx = tid % 32, ty = tid / 32;  // tx -> 4 columns, ty -> 8 rows
    const uint32_t ncols = n + 1;
    // loader roles: digits -- row drow, 8 consecutive k from dk0 (a warp covers 32 rows: conflict-free shared stores);
    // key -- 4 x uint4, consecutive lanes read consecutive 16 bytes of a KSK row
    const uint32_t drow = tid % KS_BM, dk0 = (tid / KS_BM) * 8;
    const bool drow_ok = b0 + drow < batch;
    const int8_t *dsrc = digits + (size_t)(b0 + (drow_ok ? drow : 0)) * KD + dk0;
    uint2 dreg = make_uint2(0u, 0u);
    uint4 kreg[4];
    auto prefetch = [&](uint32_t k0) {
        if (drow_ok) dreg = __ldg(reinterpret_cast<const uint2 *>(dsrc + k0));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t idx = tid + i * KS_THREADS, kk = idx / 32, c4 = idx % 32;
            kreg[i] = __ldg(reinterpret_cast<const uint4 *>(ksk + (size_t)(k0 + kk) * stride + c0 + c4 * 4));
        }
    };
    uint32_t accv[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) accv[i][j] = 0;
    prefetch(0);
    for (uint32_t k0 = 0; k0 < KD; k0 += KS_BK) {   // KD is a multiple of KS_BK (checked on the host)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t w = i < 4 ? dreg.x : dreg.y;
            sD[dk0 + i][drow] = (int32_t)(int8_t)(w >> (8 * (i & 3)));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t idx = tid + i * KS_THREADS, kk = idx / 32, c4 = idx % 32;
            *reinterpret_cast<uint4 *>(&sK[kk][c4 * 4]) = kreg[i];
        }
        __syncthreads();
        if (k0 + KS_BK < KD) prefetch(k0 + KS_BK);
#pragma unroll 8
        for (int kk = 0; kk < KS_BK; kk++) {
            int32_t av[8];
            uint32_t bv[4];
            *reinterpret_cast<int4 *>(&av[0]) = *reinterpret_cast<const int4 *>(&sD[kk][ty * 8]);
            *reinterpret_cast<int4 *>(&av[4]) = *reinterpret_cast<const int4 *>(&sD[kk][ty * 8 + 4]);
            *reinterpret_cast<uint4 *>(&bv[0]) = *reinterpret_cast<const uint4 *>(&sK[kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) accv[i][j] += (uint32_t)av[i] * bv[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t b = b0 + ty * 8 + i;
        if (b >= batch) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t c = c0 + tx * 4 + j;
            if (c >= ncols) continue;
            uint32_t v = 0u - accv[i][j];             // key_switching.rs:96 negate
            if (c == n) v += body[b];                 // key_switching.rs:98-100
            out[(size_t)b * ncols + c] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ K4-MMA
// The same wrapping-u32 product on the integer tensor cores, exactly.  key_switching.rs:88 is a matrix product
// (ndarray `dot`): out[b][c] = -sum_r D[b][r] KSK[r][c] mod 2^32 with small signed digits D.  Every key word is the sum
// of its four bytes, KSK[r][c] = sum_pl 2^(8 pl) byte_pl(r, c), so
//     sum_r D[b][r] KSK[r][c]  =  sum_pl 2^(8 pl) * ( sum_r D[b][r] byte_pl(r, c) )          (mod 2^32)
// and each inner sum is an s8 x u8 dot product whose magnitude KD * max|D| * 255 stays below 2^31 (checked on the
// host), i.e. exactly what mma.sync.m16n8k32.s32.s8.u8 computes with its 32-bit accumulators -- no rounding anywhere,
// the same bits as the IMAD kernel above.  A = digits [B][KD] (row-major, as ks_digits_kernel writes them); B = the
// byte-transposed key KSKt[c*4 + pl][KD] (k contiguous per byte column, built once at key upload).
// CTA: 128 ciphertexts x 128 byte columns (= 32 key columns), 8 warps as 4 x 2, each 32 x 64; BK = 64 digits per stage,
// 3-stage cp.async ring, operands through ldmatrix (rows padded to 80 B: conflict-free).
constexpr int KM_BM = 128, KM_BN = 128, KM_BK = 64, KM_STAGES = 3, KM_THREADS = 256, KM_ROW = KM_BK + 16;
constexpr int KM_SMEM = KM_STAGES * (KM_BM + KM_BN) * KM_ROW;
__global__ void ksk_byte_transpose_kernel(const uint32_t *__restrict__ ksk, uint8_t *__restrict__ kskt, uint32_t KD, uint32_t stride) {
    // one thread: key column c, four consecutive rows r..r+3 -> one 32-bit word of each of the four byte-plane rows
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * 4;
    if (c >= stride) return;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = ksk[(size_t)(r + i) * stride + c];
#pragma unroll
    for (int pl = 0; pl < 4; pl++) {
        const uint32_t v = ((w[0] >> (8 * pl)) & 255u) | (((w[1] >> (8 * pl)) & 255u) << 8) | (((w[2] >> (8 * pl)) & 255u) << 16) | (((w[3] >> (8 * pl)) & 255u) << 24);
        *reinterpret_cast<uint32_t *>(kskt + (size_t)(c * 4 + pl) * KD + r) = v;
    }
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_s8u8(int32_t (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(KM_THREADS, 2) ks_mma_kernel(const int8_t *__restrict__ digits, const uint8_t *__restrict__ kskt,
                                                               const uint32_t *__restrict__ body, uint32_t *__restrict__ out,
                                                               uint32_t KD, uint32_t n, uint32_t batch) {
    extern __shared__ __align__(128) uint8_t ksm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t wm = warp & 3, wn = warp >> 2;                 // warp tile: rows wm*32.., byte columns wn*64..
    const uint32_t b0 = blockIdx.y * KM_BM, n0 = blockIdx.x * KM_BN;
    auto stageA = [&](int st) { return ksm + (size_t)st * (KM_BM + KM_BN) * KM_ROW; };
    auto stageB = [&](int st) { return stageA(st) + KM_BM * KM_ROW; };
    // loader: 2 x (128 rows x 4 chunks of 16 B) per operand and stage; rows beyond the batch re-read its last row (never stored)
    auto load_stage = [&](int st, uint32_t k0) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t idx = tid + i * KM_THREADS, row = idx >> 2, ch = idx & 3;
            const uint32_t br = min(b0 + row, batch - 1u);
            cp_async16(stageA(st) + row * KM_ROW + ch * 16, digits + (size_t)br * KD + k0 + ch * 16);
            cp_async16(stageB(st) + row * KM_ROW + ch * 16, kskt + (size_t)(n0 + row) * KD + k0 + ch * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int32_t acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0;
    const uint32_t nk = KD / KM_BK;
#pragma unroll
    for (int st = 0; st < KM_STAGES - 1; st++) {
        if ((uint32_t)st < nk) load_stage(st, st * KM_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ldmatrix row addresses of this lane (see the fragment layouts of mma.m16n8k32 with 8-bit operands):
    //   A x4 = {rows 0-7 | k 0-15, rows 8-15 | k 0-15, rows 0-7 | k 16-31, rows 8-15 | k 16-31}  = a0..a3
    //   B x4 = {n-tile j | k 0-15, n-tile j | k 16-31, n-tile j+1 | k 0-15, n-tile j+1 | k 16-31} = b0, b1 of two tiles
    const uint32_t a_row = wm * 32 + (lane & 7) + ((lane >> 3) & 1) * 8, a_kb = (lane >> 4) * 16;
    const uint32_t b_row = wn * 64 + (lane >> 4) * 8 + (lane & 7), b_kb = ((lane >> 3) & 1) * 16;
    for (uint32_t kt = 0; kt < nk; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(KM_STAGES - 2) : "memory");
        __syncthreads();                                            // stage kt has landed; stage kt-1 is free again
        if (kt + KM_STAGES - 1 < nk) load_stage((kt + KM_STAGES - 1) % KM_STAGES, (kt + KM_STAGES - 1) * KM_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const uint8_t *sA = stageA(kt % KM_STAGES), *sB = stageB(kt % KM_STAGES);
#pragma unroll
        for (int ks = 0; ks < KM_BK / 32; ks++) {
            uint32_t af[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++) ldmatrix_x4(af[i], sA + (a_row + i * 16) * KM_ROW + ks * 32 + a_kb);
#pragma unroll
            for (int jp = 0; jp < 4; jp++) {
                uint32_t bf[4];
                ldmatrix_x4(bf, sB + (b_row + jp * 16) * KM_ROW + ks * 32 + b_kb);
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    mma_s8u8(acc[i][2 * jp], af[i], bf[0], bf[1]);
                    mma_s8u8(acc[i][2 * jp + 1], af[i], bf[2], bf[3]);
                }
            }
        }
    }
    // epilogue.  c-fragment: rows g / g+8 (g = lane/4), byte columns 2 tg, 2 tg + 1 (tg = lane%4) of each 8-wide tile, i.e.
    // byte planes {0,1} (tg even) or {2,3} (tg odd) of key column (n0 + wn*64 + 8 j)/4 + tg/2: lanes tg and tg^1 hold the
    // two halves of one output word.  After the exchange the even lane stores row g, the odd lane row g+8.
    const uint32_t g = lane >> 2, tg = lane & 3, ncols = n + 1;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t sh = (tg & 1u) * 16u;
            const uint32_t lo = ((uint32_t)acc[i][j][0] << sh) + ((uint32_t)acc[i][j][1] << (sh + 8u));   // row g
            const uint32_t hi = ((uint32_t)acc[i][j][2] << sh) + ((uint32_t)acc[i][j][3] << (sh + 8u));   // row g + 8
            const uint32_t lo_o = __shfl_xor_sync(0xFFFFFFFFu, lo, 1), hi_o = __shfl_xor_sync(0xFFFFFFFFu, hi, 1);
            const uint32_t sum = (tg & 1u) ? hi + hi_o : lo + lo_o;
            const uint32_t b = b0 + wm * 32 + i * 16 + g + (tg & 1u) * 8u;
            const uint32_t c = (n0 + wn * 64 + j * 8) / 4 + (tg >> 1);
            if (b < batch && c < ncols) {
                uint32_t v = 0u - sum;                    // key_switching.rs:96 negate
                if (c == n) v += body[b];                 // key_switching.rs:98-100
                out[(size_t)b * ncols + c] = v;
            }
        }
}

// Small-batch key switch (latency path): the KSK (tens of MB) is the only traffic that matters, so the rows are
// split over many CTAs and partial sums are combined with u32 atomics (addition mod 2^32 is associative, so the
// result is bit-identical to the sequential reference sum).  out must be pre-initialised by ks_init_kernel.
__global__ void ks_init_kernel(const uint32_t *__restrict__ body, uint32_t *__restrict__ out, uint32_t n, uint32_t batch) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)batch * (n + 1)) return;
    out[i] = (i % (n + 1) == n) ? body[i / (n + 1)] : 0u;      // key_switching.rs:98-100
}
constexpr int KSV_THREADS = 256, KSV_MAXB = 8;
__global__ void __launch_bounds__(KSV_THREADS) ks_gemv_kernel(const int8_t *__restrict__ digits, const uint32_t *__restrict__ ksk,
                                                              uint32_t *__restrict__ out, uint32_t KD, uint32_t n, uint32_t batch,
                                                              uint32_t rows_per_cta, uint32_t stride) {
    __shared__ int32_t sD[KSV_MAXB][64];
    const uint32_t ncols = n + 1;
    const uint32_t r0 = blockIdx.x * rows_per_cta, r1 = min(KD, r0 + rows_per_cta);
    for (uint32_t cbase = 0; cbase < ncols; cbase += KSV_THREADS * 4) {
        uint32_t accv[KSV_MAXB][4];
#pragma unroll
        for (int b = 0; b < KSV_MAXB; b++)
#pragma unroll
            for (int j = 0; j < 4; j++) accv[b][j] = 0;
        for (uint32_t rb = r0; rb < r1; rb += 64) {
            __syncthreads();
            for (uint32_t e = threadIdx.x; e < KSV_MAXB * 64; e += KSV_THREADS) {
                const uint32_t b = e / 64, rr = rb + e % 64;
                sD[b][e % 64] = (b < batch && rr < r1) ? (int32_t)digits[(size_t)b * KD + rr] : 0;
            }
            __syncthreads();
            const uint32_t rn = min(64u, r1 - rb);
            for (uint32_t rr = 0; rr < rn; rr++) {
                const uint32_t *row = ksk + (size_t)(rb + rr) * stride;
                uint32_t kv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = cbase + j * KSV_THREADS + threadIdx.x;
                    kv[j] = c < ncols ? __ldg(row + c) : 0u;
                }
#pragma unroll
                for (int b = 0; b < KSV_MAXB; b++) {
                    const uint32_t d = (uint32_t)sD[b][rr];
#pragma unroll
                    for (int j = 0; j < 4; j++) accv[b][j] += d * kv[j];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < KSV_MAXB; b++)
            if ((uint32_t)b < batch)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = cbase + j * KSV_THREADS + threadIdx.x;
                    if (c < ncols && accv[b][j]) atomicAdd(out + (size_t)b * ncols + c, 0u - accv[b][j]);  // key_switching.rs:96 negate
                }
    }
}

// ------------------------------------------------------------------------------------------ sub-op: poly_mul
// utils.rs:155-160 poly_mul (Toeplitz product) for a batch of pairs: out[b] = a[b] (*) g[b] in Z_{2^32}[X]/(X^N+1),
// a = small signed coefficients (|a| <= amax checked on the host so that N*amax*2^31 < Q0*Q1/2), g = arbitrary u32.
// Same team code as the blind rotation: NTT(a) * NTT(centred g) * N^-1 per prime, inverse NTT, CRT.
struct PolyMulArgs {
    PrimeTab prime[2];
    TwTables tw[2];
    const int32_t *a;   // [B][N]
    const uint32_t *g;  // [B][N]
    uint32_t *out;      // [B][N]
};
template <class K>
__global__ void __launch_bounds__(K::THREADS) polymul_kernel(const __grid_constant__ PolyMulArgs a) {
    using C = typename K::Ntt;
    __shared__ __align__(16) uint32_t buf[2 * 2 * C::NPAD];
    const uint32_t tid = threadIdx.x, pr = tid / K::T, t = tid % K::T;
    const uint32_t jbB = jbase_B<C>(t);
    const PrimeTab &pt = a.prime[pr];
    const TwTables &tw = a.tw[pr];
    uint32_t *buf0 = buf + (pr * 2) * C::NPAD, *buf1 = buf0 + C::NPAD;
    uint32_t *res_pr = buf0;  // the team's first exchange buffer is free once the last pass has read buf1
    const int32_t *pa = a.a + (size_t)blockIdx.x * K::N;
    const uint32_t *pg = a.g + (size_t)blockIdx.x * K::N;
    TeamRegs<K> R;
    team_init<K>(R, tw, t);
    uint32_t ahat[K::E];
    // forward NTT of a (signed small -> [0, q))
    prefetch_twB<K>(R, tw.fwdB, jbB);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const int32_t v = pa[(e << C::LOGT) | t];
        R.x[e] = v < 0 ? pt.q - (uint32_t)(-v) : (uint32_t)v;
    }
    fwd_pass_A<C>(R.x, pt.fwdA, pt.q, pt.zero);
    store_A<C>(R.x, buf0, t);
    team_bar(pr, K::T);
    phase_F2<K>(R, jbB, pt, buf0, buf1);
    team_bar(pr, K::T);
    phase_F3a<K>(R, t, pt, tw, buf1);
#pragma unroll
    for (int e = 0; e < K::E; e++) ahat[e] = csub(shoup_mul(R.x[e], 1u, pt.one_s, pt.q), pt.q);  // reduce to [0, q)
    // forward NTT of g, pointwise product, scale by N^-1
    phase_T1<K>(R, t, jbB, pt, tw, pg, buf0);
    team_bar(pr, K::T);
    phase_F2<K>(R, jbB, pt, buf0, buf1);
    team_bar(pr, K::T);
    phase_F3a<K>(R, t, pt, tw, buf1);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t ghat = csub(shoup_mul(R.x[e], pt.ninv, pt.ninv_s, pt.q), pt.q);
        R.acc[0][e] = (uint64_t)ahat[e] * ghat;
    }
    phase_I1<K>(R, t, jbB, 0, pt, tw, buf0);
    team_bar(pr, K::T);
    phase_I2<K>(R, jbB, pt, buf0, buf1);
    team_bar(pr, K::T);
    phase_I3<K>(R, t, pt, buf1, res_pr);
    __syncthreads();
    uint32_t *o = a.out + (size_t)blockIdx.x * K::N;
    for (uint32_t j = tid; j < (uint32_t)K::N; j += K::THREADS) o[j] = crt_to_u32(buf[j], buf[2 * C::NPAD + j]);
}

// ------------------------------------------------------------------------------------------ K5
// boolean.rs:18  ct_in = 2*ct1 + ct0   (lwe.rs:9-23)
__global__ void gate_linear_kernel(const uint32_t *__restrict__ ct0, const uint32_t *__restrict__ ct1, uint32_t *__restrict__ out, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = ct1[i] * 2u + ct0[i];
}
// general operand of tfhe_negacyclic_mul: byte `limb` of every word of a (unsigned, < 256: inside the exact range of the transform)
__global__ void byte_limb_kernel(const int32_t *__restrict__ a, int32_t *__restrict__ out, int limb, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = (int32_t)(((uint32_t)a[i] >> (8 * limb)) & 0xFFu);
}
// acc (+)= part << shift  (mod 2^32); first = 1 overwrites
__global__ void shl_accumulate_kernel(uint32_t *__restrict__ acc, const uint32_t *__restrict__ part, int shift, int first, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) acc[i] = (first ? 0u : acc[i]) + (part[i] << shift);
}
// NAND/NOR/XNOR = trivial(1) - gate:  out[j] = -x[j];  out[n] += encode(1)   
__global__ void gate_negate_kernel(uint32_t *__restrict__ x, const uint8_t *__restrict__ gates, int all, uint32_t n, uint32_t batch, uint32_t one) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)batch * (n + 1)) return;
    const uint32_t b = (uint32_t)(i / (n + 1)), c = (uint32_t)(i % (n + 1));
    if (all || gates[b] >= 3) {
        uint32_t v = 0u - x[i];
        if (c == n) v += one;
        x[i] = v;
    }
}
__global__ void switch_modulus_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t len, int logn) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = mod_switch(in[i], logn);
}
template <int LOGB, int L>
__global__ void decompose_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    int32_t d[L];
    decompose_signed<LOGB, L>(in[i], d);
#pragma unroll
    for (int lev = 0; lev < L; lev++) out[i * L + lev] = (uint32_t)d[lev];
}
// glwe.rs:20-34: every polynomial of GLWE b times X^{rot[b]}, rot already reduced to [0, 2N)
__global__ void mul_monomial_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ rot, uint32_t *__restrict__ out,
                                    uint32_t polys_per_ct, int logn, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t N = 1u << logn;
    const size_t poly = i >> logn;
    const uint32_t j = (uint32_t)(i & (N - 1u));
    out[i] = rot_coeff(in + poly * N, j, rot[poly / polys_per_ct], logn);
}
__global__ void sample_extract_kernel(const uint32_t *__restrict__ glwe, uint32_t *__restrict__ out, uint32_t k, int logn, uint32_t batch) {
    const uint32_t N = 1u << logn, kN = k * N;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * (kN + 1)) return;
    const uint32_t b = (uint32_t)(gid / (kN + 1)), idx = (uint32_t)(gid % (kN + 1));
    const uint32_t *g = glwe + (size_t)b * (k + 1) * N;
    uint32_t v;
    if (idx == kN) v = g[(size_t)k * N];
    else {
        const uint32_t p = idx >> logn, tt = idx & (N - 1u);
        v = tt == 0 ? g[(size_t)p * N] : 0u - g[(size_t)p * N + N - tt];
    }
    out[gid] = v;
}

// ------------------------------------------------------------------------------------------ peaks
// Integer-pipe peak microbenchmarks: 8 independent dependency chains per thread, 8-way unrolled.
//   KIND 0 IMAD, 1 IMAD.HI, 2 IMAD.WIDE (64 lane-ops per inner iteration each);
//   KIND 3 lazy Shoup/Harvey butterflies in registers exactly as the NTT passes issue them
//          (IMAD.HI + 2 IMAD + 2 IADD3 each): 32 butterflies per inner iteration -- the achievable
//          butterfly rate is the practical ceiling of the blind-rotation kernel.
template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *sink, uint32_t a, uint32_t b, int iters) {
    uint32_t x[8];
    unsigned long long y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = x[i]; }
    if (KIND >= 3) {
        // KIND 3: one (w, ws) pair shared by all butterflies, q in a register.
        // KIND 4: a different (w, ws) register pair per butterfly (as in the NTT passes), q in a register.
        // KIND 5: like 4 with q as a compile-time immediate.
        const uint32_t qr = (KIND == 5) ? kQ0 : (kQ0 + (b & 0u));
        const uint32_t w = a % kQ0, ws = (uint32_t)(((unsigned long long)w << 32) / kQ0), z = b & 0u;
        uint32_t wv[4], wsv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { wv[i] = w + 17u * i * (KIND >= 4) + threadIdx.x * (KIND >= 4); wsv[i] = ws + 13u * i * (KIND >= 4); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (KIND == 5) ct_bfly(x[i], x[i + 4], wv[i] + u, wsv[i], kQ0, z);
                    else ct_bfly(x[i], x[i + 4], wv[i] + u, wsv[i], qr, z);
                }
                const uint32_t q = kQ0;
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = umin_u32(x[i], x[i] - 8u * q);  // keep the lazy range bounded (ALU only)
            }
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (KIND == 0) x[i] = x[i] * a + b;                                  // IMAD
                    else if (KIND == 1) x[i] = __umulhi(x[i], a) + b;                    // IMAD.HI
                    else y[i] = (unsigned long long)(uint32_t)y[i] * a + y[i];         // IMAD.WIDE (loop-variant multiplicand)
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + (uint32_t)y[i] + (uint32_t)(y[i] >> 32);
    if (s == 0x12345678u) sink[0] = s;
}

// FMA-bound butterfly stream with realistic register variety: 16 values, 8 butterflies per stage with 8 distinct
// (w, ws) register pairs, 4 stages per round like one register pass; no range corrections (values may wrap --
// only the instruction stream matters).  64 butterflies... 32 per inner iteration (4 stages x 8).
__global__ void __launch_bounds__(128, 3) bfly_stream_kernel(uint32_t *sink, uint32_t a, uint32_t b, int iters) {
    uint32_t x[16];
    uint2 tw[8];
    const uint32_t q = kQ0 + (b & 0u), z = b & 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 7u + i;
#pragma unroll
    for (int i = 0; i < 8; i++) tw[i] = uint2{a % kQ0 + 17u * i + threadIdx.x, a + 13u * i};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 8 >> s;
#pragma unroll
            for (int e = 0; e < 16; e++)
                if ((e & half) == 0) ct_bfly(x[e], x[e + half], tw[(e + s) & 7].x, tw[(e + s) & 7].y, q, z);
        }
    }
    uint32_t sacc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sacc += x[i];
    if (sacc == 0x12345678u) sink[0] = sacc;
}

}  // namespace tfhe
