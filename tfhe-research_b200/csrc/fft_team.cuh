// fft_team.cuh -- exact negacyclic products mod 2^32 through a double-precision FFT with a limb-split key
// (second arithmetic path of the blind rotation; the first is the 2-prime NTT of pbs_team.cuh).
//
// Why: on B200 the FP64 pipe (DFMA/DADD/DMUL, 64 lanes/clk/SM = 18.5 T op/s measured, profiles/r01_fp64_peak.json)
// is a separate pipe from the FMA-heavy pipe that IMAD / IMAD.HI / IMAD.WIDE share, and the 2-prime NTT is bound by
// that integer pipe.  A folded complex FFT of size M = N/2 needs 6 DP ops per butterfly for two real coefficients
// where the NTT needs 2 primes x (IMAD.HI + 2 IMAD) = 16 pipe cycles.
//
// Reference semantics reproduced (bit for bit): ggsw.rs:132-161 external_product, utils.rs:113-173 poly_dot_product
//   out[c] = sum_r dec[r] (*) G[r][c]  in Z_{2^32}[X]/(X^N+1).
// Exactness argument:
//   * digits d are small signed integers, |d| <= B = 2^LOGB (decomposer.rs:42-80 incl. the +B quirk);
//   * every key word g is lifted to its centred representative s in [-2^31, 2^31) and split into two centred limbs
//     s = lo + 2^16 hi, lo in [-2^15, 2^15), hi in [-2^15, 2^15];  out = sum d(*)lo + 2^16 sum d(*)hi  (mod 2^32);
//   * each limb convolution is an integer of magnitude <= S = ROWS*N*B*2^15 <= 2^36.6 (P2); the floating-point result
//     differs from it by at most  c * 2^-53 * sum_r ||d_r||_2 ||limb_r||_2,  c = 3 (log2(M) (1 + sqrt 5) + sum over stages of
//     the twiddle error in ulps) <= 332 = 2^8.4 INCLUDING the twiddles derived by repeated squaring (derive_pass_tw: an
//     entry v squarings below the loaded one is off by <= 2^v + 3.2 (2^v - 1) + 3 ulps), i.e. <= 2^-8.0 (P2), 2^-9.3 (P1),
//     2^-12.8 (P0), so rounding to nearest recovers the exact integer;  the static_assert below is S c 2^-53 < 1/4 for any
//     c <= 2^9, and the kernel can record the largest distance to an integer it ever saw (CHECK), which the GPU tests
//     assert to be < 2^-6 on random AND adversarial inputs (tests/test_gpu_exactness.py).
//
// Transform: the fold z_j = a_j + i a_{j+M} maps R[X]/(X^N+1) to C[X]/(X^M - i); the forward transform evaluates at
// the M roots of X^M = i (zeta^(4k+1), zeta = exp(2 pi i / 4M)) with a Cooley-Tukey flow (natural in, bit-reversed
// out, merged twiddles: no separate twist), the inverse is the mirrored Gentleman-Sande flow with conjugate
// twiddles; the 1/M scaling is folded into the stored key (a power of two: exact).
//
// Everything here is __host__ __device__ so tests/emu can step the identical code on the CPU.
#pragma once
#include <math.h>

#include "tfhe_core.cuh"

// slots of the key ring (see FftPbsCfg): one slot = one GGSW row (both limbs)
#ifndef TFHE_FFT_NSLOT
#define TFHE_FFT_NSLOT 2
#endif

namespace tfhe {
namespace fft {

struct alignas(16) cplx {
    double re, im;
};

// ---- pinned FP64 operations (no compiler contraction: the CPU emulation reproduces the GPU doubles) ----
TFHE_HD double fma_d(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return ::fma(a, b, c);
#endif
}
TFHE_HD double mul_d(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
TFHE_HD double add_d(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// small signed integer -> double without the conversion unit: bits of (2^52 + 2^31 + d), minus the bias
TFHE_HD double i2d(int32_t d) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__hiloint2double(0x43300000, (int)((uint32_t)d ^ 0x80000000u)), -4503601774854144.0);
#else
    return (double)d;
#endif
}
// round to nearest integer, returned mod 2^32 (|x| < 2^51); *frac receives |x - rint(x)| when CHECK
template <bool CHECK>
TFHE_HD uint32_t round_u32(double x, double &maxfrac) {
#if defined(__CUDA_ARCH__)
    const double t = __dadd_rn(x, 6755399441055744.0);  // 1.5 * 2^52: low mantissa word = rint(x) mod 2^32
    if (CHECK) maxfrac = fmax(maxfrac, fabs(__dadd_rn(x, -__dadd_rn(t, -6755399441055744.0))));
    return (uint32_t)__double2loint(t);
#else
    const double r = nearbyint(x);
    if (CHECK) maxfrac = fmax(maxfrac, fabs(x - r));
    return (uint32_t)(int64_t)r;
#endif
}

// Cooley-Tukey butterfly  X' = X + wY,  Y' = X - wY = 2X - X'   (6 DFMA)
TFHE_HD void ct_bfly(cplx &X, cplx &Y, const cplx w) {
    const double xr = fma_d(-Y.im, w.im, fma_d(Y.re, w.re, X.re));
    const double xi = fma_d(Y.im, w.re, fma_d(Y.re, w.im, X.im));
    Y.re = fma_d(2.0, X.re, -xr);
    Y.im = fma_d(2.0, X.im, -xi);
    X.re = xr;
    X.im = xi;
}
// Gentleman-Sande butterfly of the inverse  X' = X + Y,  Y' = (X - Y) conj(w)   (4 DADD + 2 DMUL + 2 DFMA)
TFHE_HD void gs_bfly(cplx &X, cplx &Y, const cplx w) {
    const double dr = add_d(X.re, -Y.re), di = add_d(X.im, -Y.im);
    X.re = add_d(X.re, Y.re);
    X.im = add_d(X.im, Y.im);
    Y.re = fma_d(di, w.im, mul_d(dr, w.re));
    Y.im = fma_d(di, w.re, -mul_d(dr, w.im));
}

// ---------------------------------------------------------------- team geometry (mirrors NttCfg)
// M = 2^LOGM complex points per polynomial, a TEAM of T = M/E threads, E = 2^LOGE points per thread.
// Index bits of j (MSB first):  hA[LOGE] | mid[QB] | lo[LOGE].
//   layout A: regs <-> hA                      thread t <-> (mid|lo)          j = (e << LOGT) | t
//   layout B: regs <-> mid | top XB bits of lo thread   <-> hA + low QB bits of lo
//   layout C: regs <-> lo                      thread t <-> (hA|mid)          j = (t << LOGE) | e
// Exchange buffers hold 16-byte elements at phys(j) = j + (j >> LOGE): additive over disjoint bit sets (register
// offsets are immediates) and conflict-free for 128-bit accesses in all three layouts (8 consecutive lanes hit
// 8 distinct 16-byte bank groups).
template <int LOGM_, int LOGE_>
struct FftCfg {
    static constexpr int LOGM = LOGM_, LOGE = LOGE_;
    static constexpr int M = 1 << LOGM, E = 1 << LOGE, LOGT = LOGM - LOGE, T = 1 << LOGT;
    static constexpr int QB = LOGM - 2 * LOGE, XB = LOGE - QB;
    static constexpr int NA_TW = E - 1, NB_TW = (1 << QB) - 1, NC_TW = E - 1;
    static constexpr int MPAD = M + (M >> LOGE);
    static_assert(QB >= 1 && QB <= LOGE, "unsupported (LOGM, LOGE)");
    static_assert(LOGT >= 5, "team must be at least one warp");
};
template <class C>
TFHE_HD constexpr uint32_t cphys(uint32_t j) { return j + (j >> C::LOGE); }
template <class C>
TFHE_HD constexpr uint32_t jbase_B(uint32_t t) {
    return (t & ((1u << C::QB) - 1u)) | ((t >> C::QB) << (C::LOGE + C::QB));
}

template <class C> TFHE_HD void store_A(const cplx *x, cplx *buf, uint32_t t) {
    cplx *b = buf + cphys<C>(t);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; b[cphys<C>(e << C::LOGT)] = x[e]; });
}
template <class C> TFHE_HD void load_A(cplx *x, const cplx *buf, uint32_t t) {
    const cplx *b = buf + cphys<C>(t);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = b[cphys<C>(e << C::LOGT)]; });
}
template <class C> TFHE_HD void store_B(const cplx *x, cplx *buf, uint32_t jbB) {
    cplx *b = buf + cphys<C>(jbB);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; b[cphys<C>(e << C::QB)] = x[e]; });
}
template <class C> TFHE_HD void load_B(cplx *x, const cplx *buf, uint32_t jbB) {
    const cplx *b = buf + cphys<C>(jbB);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = b[cphys<C>(e << C::QB)]; });
}
// ---- the one shared-memory exchange of the tensor-memory-tail transform (M = 512, T = 64; fft_tmem.cuh).  Layout A as above;
// layout B'': registers <-> mid (j5 j4 j3), warp <-> j8, lane (L4..L0) <-> (j2 j1 j0 j6 j7) -- the bits the two tensor-memory swaps
// take next, (j2 j1) and then j0, are the top lane bits.  Elements sit unpadded at swz(j) = j ^ (((j >> 6) & 3) << 1): a quarter-warp
// covers 8 distinct 16-byte bank groups in layout A (lanes <-> j2 j1 j0) and in layout B'' (lanes <-> j0 j6 j7).
TFHE_HD constexpr uint32_t swz9(uint32_t j) { return j ^ (((j >> 6) & 3u) << 1); }
TFHE_HD constexpr uint32_t jbase_Bsw(uint32_t t) {   // t = warp << 5 | lane
    return (((t >> 5) & 1u) << 8) | ((t & 1u) << 7) | (((t >> 1) & 1u) << 6) | (((t >> 4) & 1u) << 2) | (((t >> 3) & 1u) << 1) | ((t >> 2) & 1u);
}
TFHE_HD constexpr uint32_t hA_Bsw(uint32_t t) { return jbase_Bsw(t) >> 6; }
template <class C> TFHE_HD void store_Asw(const cplx *x, cplx *buf, uint32_t t) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; buf[(e << C::LOGT) + (t ^ ((e & 3u) << 1))] = x[e]; });
}
template <class C> TFHE_HD void load_Asw(cplx *x, const cplx *buf, uint32_t t) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = buf[(e << C::LOGT) + (t ^ ((e & 3u) << 1))]; });
}
template <class C> TFHE_HD void store_Bsw(const cplx *x, cplx *buf, uint32_t jb_swz) {   // jb_swz = swz9(jbase_Bsw(t))
    cplx *b = buf + jb_swz;
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; b[e << C::QB] = x[e]; });
}
template <class C> TFHE_HD void load_Bsw(cplx *x, const cplx *buf, uint32_t jb_swz) {
    const cplx *b = buf + jb_swz;
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = b[e << C::QB]; });
}
// ---- the same for M = 1024 at 16 points per thread.  Layout B''16: registers <-> (j5 j4 j3 j2), warp <-> j9, lane (L4..L0) <->
// (j1 j0 j8 j7 j6); elements at swz10(j) = j ^ ((j >> 6) & 7).  Bit 2 of that XOR (j8, a thread bit) flips register bit j2: a thread
// addresses its even and its odd registers from two bases.
TFHE_HD constexpr uint32_t swz10(uint32_t j) { return j ^ ((j >> 6) & 7u); }
TFHE_HD constexpr uint32_t jbase_Bsw16(uint32_t t) {   // t = warp << 5 | lane
    return (((t >> 5) & 1u) << 9) | (((t >> 2) & 1u) << 8) | (((t >> 1) & 1u) << 7) | ((t & 1u) << 6) | (((t >> 4) & 1u) << 1) | ((t >> 3) & 1u);
}
template <class C> TFHE_HD void store_Asw16(const cplx *x, cplx *buf, uint32_t t) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; buf[(e << C::LOGT) + (t ^ (e & 7u))] = x[e]; });
}
template <class C> TFHE_HD void load_Asw16(cplx *x, const cplx *buf, uint32_t t) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = buf[(e << C::LOGT) + (t ^ (e & 7u))]; });
}
// even / odd: buf + swz10(jbase) for register 0, and the same with bit 2 flipped back for the odd registers (see above)
template <class C> TFHE_HD void store_Bsw16(const cplx *x, cplx *even, cplx *odd) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; ((e & 1u) ? odd : even)[(e >> 1) << 3] = x[e]; });
}
template <class C> TFHE_HD void load_Bsw16(cplx *x, const cplx *even, const cplx *odd) {
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = ((e & 1u) ? odd : even)[(e >> 1) << 3]; });
}
template <class C> TFHE_HD void store_C(const cplx *x, cplx *buf, uint32_t t) {
    cplx *b = buf + cphys<C>(t << C::LOGE);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; b[cphys<C>(e)] = x[e]; });
}
template <class C> TFHE_HD void load_C(cplx *x, const cplx *buf, uint32_t t) {
    const cplx *b = buf + cphys<C>(t << C::LOGE);
    static_for<0, C::E>([&](auto ei) { constexpr uint32_t e = decltype(ei)::value; x[e] = b[cphys<C>(e)]; });
}

// register passes: stage u pairs register bit LOGE-1-u; twiddle index 2^u - 1 + (top u bits of e)
template <int LOGE, int NST>
TFHE_HD void fwd_pass(cplx *x, const cplx *tw) {
    static_for<0, NST>([&](auto ui) {
        constexpr int u = decltype(ui)::value;
        constexpr int bit = 1 << (LOGE - 1 - u);
        static_for<0, (1 << LOGE)>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) ct_bfly(x[e], x[e + bit], tw[(1 << u) - 1 + (e >> (LOGE - u))]);
        });
    });
}
// stages [U0, U1) of a pass (same twiddle indexing): lets a caller fetch the twiddles of the late stages late
template <int LOGE, int U0, int U1>
TFHE_HD void fwd_pass_range(cplx *x, const cplx *tw) {
    static_for<U0, U1>([&](auto ui) {
        constexpr int u = decltype(ui)::value;
        constexpr int bit = 1 << (LOGE - 1 - u);
        static_for<0, (1 << LOGE)>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) ct_bfly(x[e], x[e + bit], tw[(1 << u) - 1 + (e >> (LOGE - u))]);
        });
    });
}
template <int LOGE, int U0, int U1>
TFHE_HD void inv_pass_range(cplx *x, const cplx *tw) {   // stages U1-1 down to U0
    static_for<0, U1 - U0>([&](auto ui) {
        constexpr int u = U1 - 1 - decltype(ui)::value;
        constexpr int bit = 1 << (LOGE - 1 - u);
        static_for<0, (1 << LOGE)>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) gs_bfly(x[e], x[e + bit], tw[(1 << u) - 1 + (e >> (LOGE - u))]);
        });
    });
}
template <int LOGE, int NST>
TFHE_HD void inv_pass(cplx *x, const cplx *tw) {
    static_for<0, NST>([&](auto ui) {
        constexpr int u = NST - 1 - decltype(ui)::value;
        constexpr int bit = 1 << (LOGE - 1 - u);
        static_for<0, (1 << LOGE)>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) gs_bfly(x[e], x[e + bit], tw[(1 << u) - 1 + (e >> (LOGE - u))]);
        });
    });
}

// ---------------------------------------------------------------- blind-rotation configuration
constexpr int kMaxTwA = 15;
struct TwTablesF {
    cplx twA[kMaxTwA];   // pass A (stages 0..LOGE-1): the same for every thread -> kernel-parameter constant bank
    const cplx *twB;     // [2^LOGE (hA)][NB_TW]
    const cplx *twC;     // [NC_TW][T]  (consecutive lanes read consecutive 16 bytes)
    const cplx *ztab;    // [2N] zeta^m, zeta = exp(2 pi i / 2N): monomial factors of the BMMP variant
    const cplx *twX;     // [52] per-lane entries of the tensor-memory-exchange passes (fft_tmem.cuh TmemTw), M = 256 only
};

// Work split of one ciphertext (one "team"): P = k+1 SUB-TEAMS of T threads.  Sub-team s owns polynomial s of the GLWE
// being decomposed (its digits and their L forward transforms) AND column s of the result (its multiply-accumulate
// against column s of every GGSW row, its two inverse transforms -- low and high key limb -- and the update of acc[s]).
// Per level the sub-teams exchange their transformed digit rows through shared memory (xbuf), so every forward
// transform is computed once and every thread carries only 2 x E complex accumulators.
// SINGLE_BUF: one exchange buffer per sub-team instead of two (N = 2048: 17 KB each), at the price of a barrier between
// every load and the next store.  HALVES: a key slot holds both limbs of one GGSW row for 1/HALVES of the points, so the
// two-slot TMA ring stays at 2 x 32 KB when a row is 64 KB.
// NSLOT: depth of the key ring.  Two slots feed a CTA whose 2-4 ciphertexts keep the SM busy; a CTA that holds ONE ciphertext
// (small batches: latency) is bound by the round trip of each ring refill instead, and uses the idle shared memory for a deep ring.
// XCHG: 0 = the register passes of a transform exchange through shared memory (store_A/load_B, ...); 1 = through tensor memory
// (fft_tmem.cuh: warp-sized sub-teams, M = 256), which also changes the spectral layout of the stored key; 2 = shared-memory
// exchanges with the derived twiddles of passes B and C kept in tensor memory instead of being re-derived in every pass;
// 3 = the second shared-memory exchange and pass C replaced by tensor-memory swaps (another spectral layout of the key).
template <int LOGN_, int LOGE_, int K_, int L_, int LOGB_, int CTS_, bool CHECK_ = true, bool SINGLE_BUF_ = false, int HALVES_ = 1, int NSLOT_ = TFHE_FFT_NSLOT, int XCHG_ = 0>
struct FftPbsCfg {
    using F = FftCfg<LOGN_ - 1, LOGE_>;
    static constexpr int LOGN = LOGN_, N = 1 << LOGN_, M = N / 2;
    static constexpr int K = K_, P = K_ + 1, L = L_, LOGB = LOGB_, ROWS = P * L;
    static constexpr int E = F::E, T = F::T;
    static constexpr int CTS = CTS_;                    // ciphertexts (teams) per CTA, sharing one key stream
    static constexpr bool CHECK = CHECK_, SINGLE_BUF = SINGLE_BUF_, XCHG = XCHG_ == 1;
    static constexpr bool TWT = XCHG_ == 2;
    static constexpr bool TAIL = XCHG_ == 3;  // M = 512, two warps per sub-team: pass A, ONE shared-memory exchange, pass B, then the last three stages with tensor-memory swaps inside each warp (fft_tmem.cuh)
    static_assert(!TAIL || (LOGN_ == 10 && LOGE_ == 3 && !SINGLE_BUF_ && HALVES_ == 1) || (LOGN_ == 11 && LOGE_ == 4 && SINGLE_BUF_), "tensor-memory tail: M = 512 at 8 points per thread, or M = 1024 at 16 points per thread with the single exchange buffer");
    static constexpr bool TAIL16 = TAIL && LOGE_ == 4;   // shared-memory exchanges, but a thread's derived pass twiddles wait in tensor memory (own-row-first loop)
    static_assert(!XCHG || (LOGN_ == 9 && LOGE_ == 3 && !SINGLE_BUF_ && HALVES_ == 1 && CTS_ == 4 && K_ <= 3), "tensor-memory exchanges: M = 256, 8 points per thread, one warp per sub-team, one team per lane quarter");
    static constexpr int HALVES = HALVES_, EH = E / HALVES_, MH = M / HALVES_;   // points per thread / per polynomial in one slot
    static constexpr int WARPS_PER_SUB = T / 32, TEAM_THREADS = P * T, THREADS = CTS * TEAM_THREADS;
    static_assert(LOGB * L <= 32 && 32 % LOGB == 0, "decomposer must divide log_q (SURVEY 9-B H2)");
    // digits of the levels after the first wait in a thread-private stash: int8 when they fit ([-B/2, B], B <= 64)
    using stash_t = typename std::conditional<(LOGB <= 6), int8_t, int16_t>::type;
    static_assert(LOGB <= 14, "digits are stashed as int16 at most");
    // exactness: largest limb convolution * 2^9 (error constant incl. safety) must stay below 2^51
    static_assert((double)ROWS * N * (double)(1 << LOGB) * 32768.0 * 512.0 < 2251799813685248.0, "FP64 exactness bound");
    // key stream: one SLOT = one GGSW row (or 1/HALVES of its points), both limbs: [2 limbs][P columns][MH] complex in slot
    // order; rows are stored in consumption order (level-major: row index lev*P + p for polynomial p, level lev)
    static constexpr int POLY_BYTES = MH * 16, LIMB_BYTES = P * POLY_BYTES, SLOT_BYTES = 2 * LIMB_BYTES, SLOTS_PER_STEP = ROWS * HALVES;
    static constexpr int NSLOT = NSLOT_;
    static constexpr size_t GGSW_BYTES = (size_t)SLOTS_PER_STEP * SLOT_BYTES;
    // shared memory per team: acc, then per sub-team {stash, buf0, buf1}, then the mod-switched mask
    static constexpr int TM_ACC = 0;                                   // u32 acc[P][N]
    static constexpr int STASH_BYTES = ((L > 1 ? (L - 1) : 1) * 2 * E * T * (int)sizeof(stash_t) + 15) & ~15;  // [(L-1)*2E][T]
    static constexpr int NBUF = XCHG ? 0 : SINGLE_BUF ? 1 : 2;   // tensor-memory exchanges: no exchange buffer, rows are published in tensor memory too
    static constexpr int SUB_BYTES = STASH_BYTES + NBUF * F::MPAD * 16;   // + cplx buf[NBUF][MPAD]
    static constexpr int TM_SUB = TM_ACC + P * N * 4;
    static constexpr int TM_AT = TM_SUB + P * SUB_BYTES;               // u16 at[n+1] (size known at launch)
    static constexpr int team_bytes(int n) { return (TM_AT + (n + 1) * 2 + 127) & ~127; }
    // single-ciphertext modes borrow idle shared memory for the polynomial to decompose and a zero subtrahend: the
    // accumulators of teams 1 and 2, or (two teams) the accumulator and the first exchange buffer of team 1
    static constexpr bool HAS_SINGLE_MODES = !XCHG && (CTS >= 3 || (CTS == 2 && F::MPAD * 16 >= P * N * 4));   // CTS == 1 (latency configuration): blind rotation only
    static constexpr int SPARE_DIN = 1 * 0 + TM_ACC;                         // offset inside team 1
    static constexpr int SPARE_ZERO_TEAM = CTS >= 3 ? 2 : 1;
    static constexpr int SPARE_ZERO = CTS >= 3 ? TM_ACC : TM_SUB + STASH_BYTES;   // offset inside team SPARE_ZERO_TEAM
};
// index of GGSW row (polynomial p, level lev) in the stored key = its position in the consumption order
template <class K>
TFHE_HD constexpr uint32_t key_row_index(uint32_t p, uint32_t lev) { return lev * K::P + p; }
// Device layout ("diagonal-major"): the P x P polynomials (GGSW row of polynomial p, column c) of a level are regrouped so
// that slot d of the level holds, at column position c, the polynomial of row p = (c + d) mod P.  Sub-team s always reads
// column position s; at slot d it multiplies the transformed digits of polynomial (s + d) mod P -- its OWN row at d = 0,
// whatever s is.  So every sub-team can start a level on the row it still holds in registers while the slots stream in one
// fixed order through a two-slot ring.
template <class K>
TFHE_HD constexpr uint32_t key_slot_index(uint32_t p, uint32_t c, uint32_t lev) { return lev * K::P + (p + K::P - c) % K::P; }
// every (slot d, column position c) of a level is hit exactly once, by the row p = (c + d) mod P the kernels multiply there
template <class K>
constexpr bool key_slot_layout_ok() {
    for (uint32_t lev = 0; lev < (uint32_t)K::L; lev++)
        for (uint32_t d = 0; d < (uint32_t)K::P; d++)
            for (uint32_t c = 0; c < (uint32_t)K::P; c++)
                if (key_slot_index<K>((c + d) % K::P, c, lev) != lev * K::P + d) return false;
    return true;
}

template <class K>
struct FftRegs {
    cplx x[K::E];
    cplx acc[2][K::E];  // [limb][point] of this sub-team's column
};
template <class K>
TFHE_HD void zero_acc(FftRegs<K> &r) {
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
        for (int e = 0; e < K::E; e++) r.acc[l][e] = cplx{0.0, 0.0};
}

// centred limbs of a key word (see header): s = lo + 2^16 hi
TFHE_HD int32_t key_limb(uint32_t g, int limb) {
    const int32_t s = (int32_t)g;
    const int32_t lo = (int32_t)(int16_t)(uint16_t)(g & 0xFFFFu);
    return limb == 0 ? lo : (int32_t)(((int64_t)s - lo) >> 16);
}

// ---- F1: digits of row `row` = (polynomial p, level lev), folded to complex, pass A, store to buf0.
// minuend(p, j) - subtrahend(p, j) is coefficient j of polynomial p of the GLWE to decompose (glwe.rs:69-108).  The
// decomposition (decomposer.rs:27-80) of a thread's 2E coefficients is done once per polynomial, at level 0; the
// other levels' digits wait in a thread-private stash (no barrier: written and read by the same thread).
// F1 = F1a (registers only: digits, pass A) followed by store_A into buf0.
// WHICH: 0 = any level (run-time test), 1 = the caller knows lev == 0, 2 = the caller knows lev > 0 (then diff is not used:
// a level loop peeled this way does not keep the operands of diff alive across the later levels)
template <class K, int WHICH = 0, class DiffFn>
TFHE_HD void phase_F1a(FftRegs<K> &r, uint32_t t, uint32_t p, uint32_t lev, typename K::stash_t *stash, const cplx *twA, DiffFn diff) {
    using C = typename K::F;
    // stash word = the two digits (coefficients j and j + M: real and imaginary part of point e) of one level
    using pair_t = typename std::conditional<sizeof(typename K::stash_t) == 1, uint16_t, uint32_t>::type;
    constexpr int SB = 8 * (int)sizeof(typename K::stash_t);
    pair_t *sp = reinterpret_cast<pair_t *>(stash);
    if (WHICH == 1 || (WHICH == 0 && lev == 0)) {
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            int32_t d[2][K::L];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t j = (((uint32_t)e << C::LOGT) | t) + (uint32_t)h * K::M;
                // optional third argument: 2e + h, the position among this thread's coefficients (register-resident operands)
                if constexpr (std::is_invocable<DiffFn, uint32_t, uint32_t, int>::value) decompose_signed<K::LOGB, K::L>(diff(p, j, 2 * e + h), d[h]);
                else decompose_signed<K::LOGB, K::L>(diff(p, j), d[h]);
            }
#pragma unroll
            for (int l = 1; l < K::L; l++)
                sp[((l - 1) * K::E + e) * K::T + t] = (pair_t)(((uint32_t)d[0][l] & ((1u << SB) - 1u)) | ((uint32_t)d[1][l] << SB));
            r.x[e] = cplx{i2d(d[0][0]), i2d(d[1][0])};
        }
    } else {
        const pair_t *s = sp + (size_t)(lev - 1) * K::E * K::T + t;
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            const uint32_t w = s[e * K::T];
            r.x[e] = cplx{i2d((int32_t)(typename K::stash_t)(w & ((1u << SB) - 1u))), i2d((int32_t)(typename K::stash_t)(w >> SB))};
        }
    }
    fwd_pass<C::LOGE, C::LOGE>(r.x, twA);
}
template <class K, class DiffFn>
TFHE_HD void phase_F1(FftRegs<K> &r, uint32_t t, uint32_t p, uint32_t lev, typename K::stash_t *stash, const cplx *twA, cplx *buf0, DiffFn diff) {
    phase_F1a<K>(r, t, p, lev, stash, twA, diff);
    store_A<typename K::F>(r.x, buf0, t);
}
// Twiddles of a pass, tw[2^u - 1 + m] for stage u and m in [0, 2^u).  With b0 the thread-dependent block index of the
// pass and S its first stage, w(S+u, (b0 << u) | m) = base_u * rho_u^brv_u(m), rho_u = exp(2 pi i / 2^(u+1)), and
// base_u = base_(u+1)^2 (the exponent of zeta splits into a b0 part and an m part; zeta^M = i).  So ONE table entry per
// pass is loaded -- the last stage's m = 0 twiddle -- and the other 2^NST - 2 are derived by squaring and by
// multiplications with i (a swap) and exp(i pi / 4): a quarter of the twiddle traffic on the shared-memory / L1 data
// pipe, which is this kernel's binding resource, for ~14 FP64 operations per pass.
TFHE_HD cplx csq(const cplx a) { return cplx{mul_d(add_d(a.re, -a.im), add_d(a.re, a.im)), mul_d(add_d(a.re, a.re), a.im)}; }
TFHE_HD cplx cmul_i(const cplx a) { return cplx{-a.im, a.re}; }
TFHE_HD cplx cmul_w8(const cplx a) {
    constexpr double r = 0.70710678118654752440;
    return cplx{mul_d(add_d(a.re, -a.im), r), mul_d(add_d(a.re, a.im), r)};
}
TFHE_HD cplx cmul_c(const cplx a, double cr, double ci) {
    return cplx{fma_d(-a.im, ci, mul_d(a.re, cr)), fma_d(a.im, cr, mul_d(a.re, ci))};
}
// the one table entry a pass of NST stages needs (a thread's entries never change: the kernel keeps them in registers)
template <int NST>
TFHE_HD cplx pass_tw_base(const cplx *table, uint32_t stride) { return table[(size_t)((1 << (NST - 1)) - 1) * stride]; }
template <int NST>
TFHE_HD void derive_pass_tw(cplx *tw, cplx b) {
    static_assert(NST >= 1 && NST <= 4, "derived twiddles are written out for passes of up to 4 stages");
    static_for<0, NST>([&](auto vi) {
        constexpr int u = NST - 1 - decltype(vi)::value;
        if constexpr (u < NST - 1) b = csq(b);
        constexpr int o = (1 << u) - 1;
        tw[o] = b;
        if constexpr (u == 1) tw[o + 1] = cmul_i(b);
        if constexpr (u == 2) {
            tw[o + 1] = cmul_i(b);          // brv_2(1) = 2: rho^2 = i
            tw[o + 2] = cmul_w8(b);         // brv_2(2) = 1
            tw[o + 3] = cmul_i(tw[o + 2]);  // brv_2(3) = 3
        }
        if constexpr (u == 3) {             // rho = exp(2 pi i / 16); brv_3(m) = 0, 4, 2, 6, 1, 5, 3, 7
            constexpr double c16 = 0.92387953251128675613, s16 = 0.38268343236508977173;
            tw[o + 1] = cmul_i(b);
            tw[o + 2] = cmul_w8(b);
            tw[o + 3] = cmul_i(tw[o + 2]);
            tw[o + 4] = cmul_c(b, c16, s16);
            tw[o + 5] = cmul_i(tw[o + 4]);
            tw[o + 6] = cmul_c(b, s16, c16);
            tw[o + 7] = cmul_i(tw[o + 6]);
        }
    });
}
template <int NST>
TFHE_HD void load_pass_tw(cplx *tw, const cplx *table, uint32_t stride) { derive_pass_tw<NST>(tw, pass_tw_base<NST>(table, stride)); }
// F2: layout B, pass B.   F3: layout C, pass C.
// (the *v variants take the pass's table entry by value: pass_tw_base<QB>(twB_thread, 1) / pass_tw_base<LOGE>(twC + t, T))
template <class K>
TFHE_HD void phase_F2v(FftRegs<K> &r, uint32_t jbB, const cplx twB_base, const cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NB_TW];
    derive_pass_tw<C::QB>(tw, twB_base);
    load_B<C>(r.x, buf0, jbB);
    fwd_pass<C::LOGE, C::QB>(r.x, tw);
    store_B<C>(r.x, buf1, jbB);
}
// (the *w variants take the pass's complete twiddle array, prepared by the caller)
template <class K>
TFHE_HD void phase_F2w(FftRegs<K> &r, uint32_t jbB, const cplx *tw, const cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    load_B<C>(r.x, buf0, jbB);
    fwd_pass<C::LOGE, C::QB>(r.x, tw);
    store_B<C>(r.x, buf1, jbB);
}
template <class K>
TFHE_HD void phase_F3w(FftRegs<K> &r, uint32_t t, const cplx *tw, const cplx *buf1) {
    using C = typename K::F;
    load_C<C>(r.x, buf1, t);
    fwd_pass<C::LOGE, C::LOGE>(r.x, tw);
}
template <class K>
TFHE_HD void phase_J1w(FftRegs<K> &r, uint32_t t, const cplx *tw, cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], tw);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], tw);
    store_C<C>(r.acc[0], buf0, t);
    store_C<C>(r.acc[1], buf1, t);
}
template <class K>
TFHE_HD void phase_J2aw(FftRegs<K> &r, uint32_t jbB, const cplx *tw, const cplx *buf0, const cplx *buf1) {
    using C = typename K::F;
    load_B<C>(r.acc[0], buf0, jbB);
    load_B<C>(r.acc[1], buf1, jbB);
    inv_pass<C::LOGE, C::QB>(r.acc[0], tw);
    inv_pass<C::LOGE, C::QB>(r.acc[1], tw);
}
template <class K>
TFHE_HD void phase_F2(FftRegs<K> &r, uint32_t jbB, const cplx *twB_thread, const cplx *buf0, cplx *buf1) {
    phase_F2v<K>(r, jbB, pass_tw_base<K::F::QB>(twB_thread, 1), buf0, buf1);
}
template <class K>
TFHE_HD void phase_F3v(FftRegs<K> &r, uint32_t t, const cplx twC_base, const cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NC_TW];
    derive_pass_tw<C::LOGE>(tw, twC_base);
    load_C<C>(r.x, buf1, t);
    fwd_pass<C::LOGE, C::LOGE>(r.x, tw);
}
template <class K>
TFHE_HD void phase_F3(FftRegs<K> &r, uint32_t t, const cplx *twC, const cplx *buf1) {
    phase_F3v<K>(r, t, pass_tw_base<K::F::LOGE>(twC + t, K::F::T), buf1);
}
// single-buffer variants: the load half and the store half of a pass are separate phases (a barrier goes between them)
template <class K>
TFHE_HD void phase_F2a(FftRegs<K> &r, uint32_t jbB, const cplx *twB_thread, const cplx *buf) {
    using C = typename K::F;
    cplx tw[C::NB_TW];
    load_pass_tw<C::QB>(tw, twB_thread, 1);
    load_B<C>(r.x, buf, jbB);
    fwd_pass<C::LOGE, C::QB>(r.x, tw);
}
template <class K>
TFHE_HD void phase_F2b(const FftRegs<K> &r, uint32_t jbB, cplx *buf) { store_B<typename K::F>(r.x, buf, jbB); }
// X: publish this sub-team's transformed digit row for the other sub-teams (slot order: point (t<<LOGE)|e at e*T + t)
template <class K>
TFHE_HD void phase_xstore(const FftRegs<K> &r, uint32_t t, cplx *xbuf) {
#pragma unroll
    for (int e = 0; e < K::E; e++) xbuf[e * K::T + t] = r.x[e];
}
// multiply-accumulate of one transformed digit row against column `col` of one GGSW row, both limbs:
// slot = [2 limbs][P][M] complex in slot order (consecutive lanes read consecutive 16 bytes); xsrc = this thread's own
// r.x (OWN) or the publishing sub-team's xbuf.
template <class K, bool OWN>
TFHE_HD void phase_mac(FftRegs<K> &r, uint32_t t, uint32_t col, const cplx *slot, const cplx *xbuf, uint32_t h = 0) {
    const cplx *g0 = slot + col * K::MH + t, *g1 = g0 + K::P * K::MH;
    static_for<0, K::HALVES>([&](auto hi) {      // compile-time register indices: h selects which EH of the E points
        constexpr int hh = decltype(hi)::value;
        if (h == (uint32_t)hh) {
#pragma unroll
            for (int q = 0; q < K::EH; q++) {
                constexpr int e0 = hh * K::EH;
                const cplx xv = OWN ? r.x[e0 + q] : xbuf[(e0 + q) * K::T + t];
                const cplx a = g0[q * K::T], b = g1[q * K::T];
                r.acc[0][e0 + q].re = fma_d(-xv.im, a.im, fma_d(xv.re, a.re, r.acc[0][e0 + q].re));
                r.acc[0][e0 + q].im = fma_d(xv.im, a.re, fma_d(xv.re, a.im, r.acc[0][e0 + q].im));
                r.acc[1][e0 + q].re = fma_d(-xv.im, b.im, fma_d(xv.re, b.re, r.acc[1][e0 + q].re));
                r.acc[1][e0 + q].im = fma_d(xv.im, b.re, fma_d(xv.re, b.im, r.acc[1][e0 + q].im));
            }
        }
    });
}
// bit reversal of the low `bits` bits
TFHE_HD uint32_t brv_bits(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
    return __brev(x) >> (32 - bits);
#else
    return brv_c(x, bits);
#endif
}
// BMMP variant (notes/BMMP Bootstrapping.md:13-25): multiply-accumulate against (X^ex - 1) * G.  In the transform
// domain the plaintext polynomial X^ex - 1 is a pointwise factor: spectral position j holds the evaluation at
// X_j = zeta^(1 + 4 brv(j)) (fft_team.cuh header), so the factor is zeta^((1 + 4 brv(j)) ex mod 2N) - 1, read from ztab.
// Result bits are unchanged (the factor only adds a relative error of a few ulp to a product that is rounded to an
// integer with > 2^5 margin; the exact statement is acc += sum_k (X^e_k - 1) ExtProd(bk_k, acc) mod 2^32).
// The exponent splits as (1 + 4 brv(j)) ex = c_t ex + (brv_LOGE(e) << (LOGT+2)) ex with c_t = 1 + 4 brv_LOGT(t): one
// per-thread table entry (bmmp_base, fetched once per step and key) times an entry whose address is the same for every
// thread (a broadcast load) -- no per-element gathers.
template <class K>
TFHE_HD cplx bmmp_base(const cplx *ztab, uint32_t t, uint32_t ex) {
    using C = typename K::F;
    return ztab[((1u + 4u * brv_bits(t, C::LOGT)) * ex) & (2u * K::N - 1u)];
}
template <class K, bool OWN>
TFHE_HD void phase_mac_bmmp(FftRegs<K> &r, uint32_t t, uint32_t col, const cplx *slot, const cplx *xbuf, const cplx *ztab, uint32_t ex,
                            const cplx base) {
    using C = typename K::F;
    static_assert(K::HALVES == 1, "the BMMP variant is instantiated for whole-row key slots only");
    // exactness: three keys per row, each times a monomial factor of modulus <= 2 -> the limb convolution is up to 6x larger
    static_assert(6.0 * (double)K::ROWS * K::N * (double)(1 << K::LOGB) * 32768.0 * 512.0 < 2251799813685248.0, "FP64 exactness bound (BMMP)");
    const cplx *g0 = slot + col * K::M + t, *g1 = g0 + K::P * K::M;
    static_for<0, K::E>([&](auto ei) {
        constexpr int e = decltype(ei)::value;
        constexpr uint32_t step = brv_c((uint32_t)e, C::LOGE) << (C::LOGT + 2);
        const cplx u = ztab[(step * ex) & (2u * K::N - 1u)];
        const double fr = add_d(fma_d(-base.im, u.im, mul_d(base.re, u.re)), -1.0);   // Re(base * u) - 1
        const double fi = fma_d(base.im, u.re, mul_d(base.re, u.im));
        const cplx xs = OWN ? r.x[e] : xbuf[e * K::T + t];
        cplx xv;
        xv.re = fma_d(-xs.im, fi, mul_d(xs.re, fr));
        xv.im = fma_d(xs.im, fr, mul_d(xs.re, fi));
        const cplx a = g0[e * K::T], b = g1[e * K::T];
        r.acc[0][e].re = fma_d(-xv.im, a.im, fma_d(xv.re, a.re, r.acc[0][e].re));
        r.acc[0][e].im = fma_d(xv.im, a.re, fma_d(xv.re, a.im, r.acc[0][e].im));
        r.acc[1][e].re = fma_d(-xv.im, b.im, fma_d(xv.re, b.re, r.acc[1][e].re));
        r.acc[1][e].im = fma_d(xv.im, b.re, fma_d(xv.re, b.im, r.acc[1][e].im));
    });
}
// ---- paired inverse: the low- and high-limb accumulators of the column are transformed TOGETHER, in place in their
// registers (twice the independent butterflies per pass, three barriers for the pair instead of four): limb 0 travels
// through buf0, limb 1 through buf1; each buffer is stored, (barrier), loaded, and only re-stored after the next barrier.
//   J1: pass C on both, store_C        | barrier |  J2a: load_B both, pass B  | barrier |  J2b: store_B both  | barrier |
//   J3: load_A both, pass A, round both, acc += lo + (hi << 16)
template <class K>
TFHE_HD void phase_J1v(FftRegs<K> &r, uint32_t t, const cplx twC_base, cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NC_TW];
    derive_pass_tw<C::LOGE>(tw, twC_base);
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], tw);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], tw);
    store_C<C>(r.acc[0], buf0, t);
    store_C<C>(r.acc[1], buf1, t);
}
template <class K>
TFHE_HD void phase_J2av(FftRegs<K> &r, uint32_t jbB, const cplx twB_base, const cplx *buf0, const cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NB_TW];
    derive_pass_tw<C::QB>(tw, twB_base);
    load_B<C>(r.acc[0], buf0, jbB);
    load_B<C>(r.acc[1], buf1, jbB);
    inv_pass<C::LOGE, C::QB>(r.acc[0], tw);
    inv_pass<C::LOGE, C::QB>(r.acc[1], tw);
}
template <class K>
TFHE_HD void phase_J1(FftRegs<K> &r, uint32_t t, const cplx *twC, cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NC_TW];
    load_pass_tw<C::LOGE>(tw, twC + t, C::T);
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], tw);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], tw);
    store_C<C>(r.acc[0], buf0, t);
    store_C<C>(r.acc[1], buf1, t);
}
template <class K>
TFHE_HD void phase_J2a(FftRegs<K> &r, uint32_t jbB, const cplx *twB_thread, const cplx *buf0, const cplx *buf1) {
    using C = typename K::F;
    cplx tw[C::NB_TW];
    load_pass_tw<C::QB>(tw, twB_thread, 1);
    load_B<C>(r.acc[0], buf0, jbB);
    load_B<C>(r.acc[1], buf1, jbB);
    inv_pass<C::LOGE, C::QB>(r.acc[0], tw);
    inv_pass<C::LOGE, C::QB>(r.acc[1], tw);
}
template <class K>
TFHE_HD void phase_J2b(const FftRegs<K> &r, uint32_t jbB, cplx *buf0, cplx *buf1) {
    using C = typename K::F;
    store_B<C>(r.acc[0], buf0, jbB);
    store_B<C>(r.acc[1], buf1, jbB);
}
template <class K>
TFHE_HD void phase_J3(FftRegs<K> &r, uint32_t t, const cplx *twA, const cplx *buf0, const cplx *buf1, uint32_t *acc_c, double &maxfrac) {
    using C = typename K::F;
    load_A<C>(r.acc[0], buf0, t);
    load_A<C>(r.acc[1], buf1, t);
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], twA);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], twA);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        acc_c[j] += round_u32<K::CHECK>(r.acc[0][e].re, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].re, maxfrac) << 16);
        acc_c[j + K::M] += round_u32<K::CHECK>(r.acc[0][e].im, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].im, maxfrac) << 16);
    }
}
// J3 that also hands the thread's 2E updated accumulator words to the next step in registers
// (accv[2e + h] = acc_c[((e << LOGT) | t) + h M]): they are the subtrahend of the next step's decomposition
template <class K>
TFHE_HD void phase_J3r(FftRegs<K> &r, uint32_t t, const cplx *twA, const cplx *buf0, const cplx *buf1, uint32_t *acc_c, uint32_t *accv, double &maxfrac) {
    using C = typename K::F;
    load_A<C>(r.acc[0], buf0, t);
    load_A<C>(r.acc[1], buf1, t);
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], twA);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], twA);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc_c[j] + round_u32<K::CHECK>(r.acc[0][e].re, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].re, maxfrac) << 16);
        accv[2 * e + 1] = acc_c[j + K::M] + round_u32<K::CHECK>(r.acc[0][e].im, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].im, maxfrac) << 16);
        acc_c[j] = accv[2 * e];
        acc_c[j + K::M] = accv[2 * e + 1];
    }
}
// the same when both accumulators already sit in layout A in their registers (tensor-memory exchanges, fft_tmem.cuh)
template <class K>
TFHE_HD void phase_J3r_regs(FftRegs<K> &r, uint32_t t, const cplx *twA, uint32_t *acc_c, uint32_t *accv, double &maxfrac) {
    using C = typename K::F;
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], twA);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], twA);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc_c[j] + round_u32<K::CHECK>(r.acc[0][e].re, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].re, maxfrac) << 16);
        accv[2 * e + 1] = acc_c[j + K::M] + round_u32<K::CHECK>(r.acc[0][e].im, maxfrac) + (round_u32<K::CHECK>(r.acc[1][e].im, maxfrac) << 16);
        acc_c[j] = accv[2 * e];
        acc_c[j + K::M] = accv[2 * e + 1];
    }
}
// the same after the swizzled exchange of the tensor-memory-tail transform
template <class K>
TFHE_HD void phase_J3r_sw(FftRegs<K> &r, uint32_t t, const cplx *twA, const cplx *buf0, const cplx *buf1, uint32_t *acc_c, uint32_t *accv, double &maxfrac) {
    using C = typename K::F;
    load_Asw<C>(r.acc[0], buf0, t);
    load_Asw<C>(r.acc[1], buf1, t);
    phase_J3r_regs<K>(r, t, twA, acc_c, accv, maxfrac);
}
// single-buffer inverse of ONE limb, in place in its accumulator registers:
//   K1: pass C, store_C | barrier | K2a: load_B, pass B | barrier | K2b: store_B | barrier | K3: load_A, pass A
// then (limb 0) keep the rounded words, (limb 1) acc += lo + (hi << 16).
template <class K, int LIMB>
TFHE_HD void phase_K1(FftRegs<K> &r, uint32_t t, const cplx *twC, cplx *buf) {
    using C = typename K::F;
    cplx tw[C::NC_TW];
    load_pass_tw<C::LOGE>(tw, twC + t, C::T);
    inv_pass<C::LOGE, C::LOGE>(r.acc[LIMB], tw);
    store_C<C>(r.acc[LIMB], buf, t);
}
template <class K, int LIMB>
TFHE_HD void phase_K2a(FftRegs<K> &r, uint32_t jbB, const cplx *twB_thread, const cplx *buf) {
    using C = typename K::F;
    cplx tw[C::NB_TW];
    load_pass_tw<C::QB>(tw, twB_thread, 1);
    load_B<C>(r.acc[LIMB], buf, jbB);
    inv_pass<C::LOGE, C::QB>(r.acc[LIMB], tw);
}
template <class K, int LIMB>
TFHE_HD void phase_K2b(const FftRegs<K> &r, uint32_t jbB, cplx *buf) { store_B<typename K::F>(r.acc[LIMB], buf, jbB); }
template <class K>
TFHE_HD void phase_K3_lo(FftRegs<K> &r, uint32_t t, const cplx *twA, const cplx *buf, uint32_t *lo, double &maxfrac) {
    using C = typename K::F;
    load_A<C>(r.acc[0], buf, t);
    inv_pass<C::LOGE, C::LOGE>(r.acc[0], twA);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        lo[2 * e] = round_u32<K::CHECK>(r.acc[0][e].re, maxfrac);
        lo[2 * e + 1] = round_u32<K::CHECK>(r.acc[0][e].im, maxfrac);
    }
}
template <class K>
TFHE_HD void phase_K3_hi(FftRegs<K> &r, uint32_t t, const cplx *twA, const cplx *buf, const uint32_t *lo, uint32_t *acc_c, double &maxfrac) {
    using C = typename K::F;
    load_A<C>(r.acc[1], buf, t);
    inv_pass<C::LOGE, C::LOGE>(r.acc[1], twA);
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        acc_c[j] += lo[2 * e] + (round_u32<K::CHECK>(r.acc[1][e].re, maxfrac) << 16);
        acc_c[j + K::M] += lo[2 * e + 1] + (round_u32<K::CHECK>(r.acc[1][e].im, maxfrac) << 16);
    }
}

// ---- one-off key transform: raw GGSW polynomial g[N] (u32) -> limb `limb`, folded, forward FFT, scaled by 1/M,
// stored in slot order.  T1 -> (barrier) -> F2 -> (barrier) -> T3.
template <class K>
TFHE_HD void phase_T1(FftRegs<K> &r, uint32_t t, int limb, const uint32_t *g, const cplx *twA, cplx *buf0) {
    using C = typename K::F;
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        r.x[e] = cplx{i2d(key_limb(g[j], limb)), i2d(key_limb(g[j + K::M], limb))};
    }
    fwd_pass<C::LOGE, C::LOGE>(r.x, twA);
    store_A<C>(r.x, buf0, t);
}
template <class K>
TFHE_HD void phase_T3(FftRegs<K> &r, uint32_t t, const cplx *twC, const cplx *buf1, cplx *out) {
    phase_F3<K>(r, t, twC, buf1);
    constexpr double scale = 1.0 / (double)K::M;  // power of two: exact
#pragma unroll
    for (int e = 0; e < K::E; e++)   // point e of this thread: half e / EH, position (e % EH) * T + t inside the half
        out[(size_t)(e / K::EH) * (2 * K::P * K::MH) + (e % K::EH) * K::T + t] = cplx{mul_d(r.x[e].re, scale), mul_d(r.x[e].im, scale)};
}

}  // namespace fft
}  // namespace tfhe
