// pbs_team.cuh -- per-thread phases of one CMUX step (external product) of the blind rotation.
//
// Reference path: bootstrapping.rs:90-105 -> ggsw.rs:164-178 (cmux) -> ggsw.rs:132-161
// (external_product) -> glwe.rs:69-108 (decompose) + utils.rs:155-173 (poly_dot_product).
//
// One CTA owns one GLWE accumulator (shared memory) and runs two TEAMS of T threads, one per RNS
// prime; both teams execute the SAME instructions (the prime and its twiddles are runtime values).
// A CMUX step is a fixed sequence of phases separated by barriers:
//   D   (all threads)  diff = rot(acc, a) - acc, signed digits of every coefficient -> dig[]
//   F_r (each team)    forward NTT of digit polynomial r (3 register passes, 2 exchanges) and
//                      multiply-accumulate against GGSW row r into 64-bit register accumulators
//   I_c (each team)    reduce accumulator column c, inverse NTT, store residues res[prime][c][]
//   R   (all threads)  CRT + centred lift mod 2^32, acc += result
// The phase functions below are __host__ __device__ so tests/emu can run the identical code for
// every thread of a CTA in lock step on the CPU.
#pragma once
#include "tfhe_core.cuh"

// unroll factor of the per-coefficient loops of the digit (D) and CRT (R) phases: independent iterations,
// so unrolling lets the shared-memory loads of several coefficients overlap
#ifndef TFHE_DR_UNROLL
#define TFHE_DR_UNROLL 16
#endif

namespace tfhe {

constexpr int kDrUnroll = TFHE_DR_UNROLL;

template <int LOGN_, int LOGE_, int K_, int L_, int LOGB_, bool STAGE_G_ = true, bool TWC_GLOBAL_ = false>
struct PbsCfg {
    using Ntt = NttCfg<LOGN_, LOGE_>;
    static constexpr int LOGN = LOGN_, N = 1 << LOGN_;
    static constexpr int K = K_, P = K_ + 1;  // GLWE dimension k, polynomials per GLWE
    static constexpr int L = L_, LOGB = LOGB_;  // PBS decomposer (decomposer.rs:2-6)
    static constexpr int ROWS = P * L;          // rows of a GGSW (ggsw.rs:37-41)
    static constexpr int E = Ntt::E, T = Ntt::T;
    static constexpr int THREADS = 2 * T;       // two primes
    static constexpr int DIG_OFF = 1 << (LOGB - 1);
    static constexpr int DIG_BYTES = (LOGB <= 7) ? 1 : 2;  // stored digit = d + B/2 in [0, 3B/2]
    static constexpr int DIG_WORDS = E * DIG_BYTES / 4;    // 32-bit words per thread per digit row
    static constexpr int DIG_SH = DIG_WORDS == 8 ? 2 : DIG_WORDS == 4 ? 3 : 4;  // log2(32 / DIG_WORDS)
    static_assert(LOGB * L <= 32 && 32 % LOGB == 0, "decomposer must divide log_q (SURVEY 9-B H2)");
    static_assert(DIG_WORDS == 2 || DIG_WORDS == 4 || DIG_WORDS == 8, "digit row geometry");
    // exactness: |sum| <= ROWS*N*B*2^31 must stay below Q0*Q1/2; lazy u64 accumulators must not wrap
    static_assert((double)ROWS * N * (1 << LOGB) * 2147483648.0 < (double)kHalfQ0Q1, "CRT range");
    static_assert((double)ROWS * (2 * LOGN + 2) * (double)kQ1 * (double)kQ1 < 18446744073709551616.0, "u64 MAC range");
    // GGSW rows are streamed global -> shared with cp.async.bulk (TMA) when STAGE_G, else read with LDG
    static constexpr bool STAGE_G = STAGE_G_;
    // forward pass-C twiddles: resident in registers for the whole kernel, or re-loaded from global/L1 for
    // every transform (saves 2*(E-1) registers; pays off when the kernel would otherwise spill)
    static constexpr bool TWC_GLOBAL = TWC_GLOBAL_;
    static constexpr int G_ROW_BYTES = P * N * 4;  // one GGSW row of one prime
    // shared memory carve-up (bytes).  res[2][P][N] (inverse-NTT residues) shares bytes with the staged GGSW
    // rows when STAGE_G (a team only writes its own half, after its own last MAC; team barriers order the
    // two), else with dig[ROWS][N] (a CTA barrier separates the last digit read from the first residue write).
    static constexpr int SM_ACC = 0;
    static constexpr int SM_DIG = SM_ACC + P * N * 4;
    static constexpr int DIG_BYTES_TOTAL = ROWS * N * DIG_BYTES, RES_BYTES_TOTAL = 2 * P * N * 4;
    static constexpr int DIG_REGION = STAGE_G ? DIG_BYTES_TOTAL : (DIG_BYTES_TOTAL > RES_BYTES_TOTAL ? DIG_BYTES_TOTAL : RES_BYTES_TOTAL);
    static constexpr int SM_BUF = SM_DIG + ((DIG_REGION + 127) & ~127);
    static constexpr int SM_G = SM_BUF + ((2 * 2 * Ntt::NPAD * 4 + 127) & ~127);      // [2 primes][G_ROW_BYTES]
    static constexpr int SM_RES = STAGE_G ? SM_G : SM_DIG;
    static constexpr int SM_BAR = SM_G + (STAGE_G ? 2 * G_ROW_BYTES : 0);              // 2 mbarriers
    static constexpr int SM_AT = SM_BAR + 16;  // mod-switched mask (u16), n+1 entries follow
};

// Byte offset of the digit of natural coefficient j inside digit row `row`.  Thread t of layout A
// holds j = (e<<LOGT)|t for e = 0..E-1 and wants those E digits in its own DIG_WORDS words; the word
// index is XOR-swizzled with bits of t so that both the writers (consecutive t, fixed e) and the
// readers (consecutive t, fixed word) hit 32 distinct banks.
template <class K>
TFHE_HD constexpr uint32_t dig_word(uint32_t t, uint32_t w) {
    return t * K::DIG_WORDS + (w ^ ((t >> K::DIG_SH) & (K::DIG_WORDS - 1u)));
}
template <class K>
TFHE_HD constexpr uint32_t dig_byte_off(uint32_t row, uint32_t j) {
    using C = typename K::Ntt;
    const uint32_t t = j & (C::T - 1u), e = j >> C::LOGT;
    const uint32_t byte_in_row = e * K::DIG_BYTES;
    return row * (K::N * K::DIG_BYTES) + dig_word<K>(t, byte_in_row >> 2) * 4u + (byte_in_row & 3u);
}

TFHE_HD void ld_global4(uint32_t *dst, const uint32_t *src) {
#if defined(__CUDA_ARCH__)
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
#else
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
#endif
}
TFHE_HD uint2 ld_global_tw(const uint2 *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- phase D: digits of diff (all THREADS threads of the CTA) -------------------------------
// diff_fn(p, j) returns coefficient j of polynomial p of the GLWE to decompose.
template <class K, class DiffFn>
TFHE_HD void phase_digits(uint32_t tid, uint8_t *dig_bytes, DiffFn diff_fn) {
#pragma unroll(kDrUnroll)
    for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::THREADS) {
        const uint32_t p = idx >> K::LOGN, j = idx & (K::N - 1u);
        int32_t d[K::L];
        decompose_signed<K::LOGB, K::L>(diff_fn(p, j), d);
#pragma unroll
        for (int lev = 0; lev < K::L; lev++) {
            const uint32_t v = (uint32_t)(d[lev] + K::DIG_OFF);
            const uint32_t off = dig_byte_off<K>(p * K::L + lev, j);
            if constexpr (K::DIG_BYTES == 1) dig_bytes[off] = (uint8_t)v;
            else *reinterpret_cast<uint16_t *>(dig_bytes + off) = (uint16_t)v;
        }
    }
}

// ---- per-team register state ------------------------------------------------------------------
template <class K>
struct TeamRegs {
    uint32_t x[K::E];
    uint64_t acc[K::P][K::E];
    uint2 twC[K::TWC_GLOBAL ? 1 : K::Ntt::NC_TW];  // forward pass-C twiddles when register resident
    uint2 twB[K::Ntt::NB_TW];  // pass-B twiddles, prefetched at the start of each transform
};

// twiddle tables in global memory, one set per prime (built on the host, see host_tables.hpp):
//   fwdB/invB: [2^LOGE (hA)][NB_TW] (w, ws)      fwdC/invC: [T (t)][NC_TW] (w, ws)
struct TwTables {
    const uint2 *fwdB, *fwdC, *invB, *invC;
};

template <class K>
TFHE_HD void team_init(TeamRegs<K> &r, const TwTables &tw, uint32_t t) {
    using C = typename K::Ntt;
    if constexpr (!K::TWC_GLOBAL) {
#pragma unroll
        for (int i = 0; i < C::NC_TW; i++) r.twC[i] = ld_global_tw(tw.fwdC + t * C::NC_TW + i);
    }
}
template <class K>
TFHE_HD void team_zero_acc(TeamRegs<K> &r) {
#pragma unroll
    for (int c = 0; c < K::P; c++)
#pragma unroll
        for (int e = 0; e < K::E; e++) r.acc[c][e] = 0;
}
template <class K>
TFHE_HD void prefetch_twB(TeamRegs<K> &r, const uint2 *tabB, uint32_t jbB) {
    using C = typename K::Ntt;
    const uint32_t hA = jbB >> (C::LOGN - C::LOGE);
#pragma unroll
    for (int i = 0; i < C::NB_TW; i++) r.twB[i] = ld_global_tw(tabB + hA * C::NB_TW + i);
}

// F1: load the E digits of row `row` (layout A), lift to [q-B/2, q+B], pass A, store to buf0.
template <class K>
TFHE_HD void phase_F1x(uint32_t *x, uint32_t t, const PrimeTab &pt, const uint8_t *dig_bytes, uint32_t row, uint32_t *buf0);
template <class K>
TFHE_HD void phase_F1(TeamRegs<K> &r, uint32_t t, uint32_t jbB, const PrimeTab &pt, const TwTables &tw, const uint8_t *dig_bytes,
                      uint32_t row, uint32_t *buf0) {
    prefetch_twB<K>(r, tw.fwdB, jbB);
    phase_F1x<K>(r.x, t, pt, dig_bytes, row, buf0);
}
// same on an explicit register array (used when row r+1's first pass is overlapped with row r's last pass)
template <class K>
TFHE_HD void phase_F1x(uint32_t *x, uint32_t t, const PrimeTab &pt, const uint8_t *dig_bytes, uint32_t row, uint32_t *buf0) {
    using C = typename K::Ntt;
    const uint32_t q = pt.q;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(dig_bytes + (size_t)row * K::N * K::DIG_BYTES);
    uint32_t w[K::DIG_WORDS];
#pragma unroll
    for (int i = 0; i < K::DIG_WORDS; i++) w[i] = src[dig_word<K>(t, i)];
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        uint32_t d;
        if constexpr (K::DIG_BYTES == 1) d = (w[e >> 2] >> (8 * (e & 3))) & 0xFFu;
        else d = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
        x[e] = d + (q - K::DIG_OFF);
    }
    fwd_pass_A<C>(x, pt.fwdA, q, pt.zero);
    store_A<C>(x, buf0, t);
}
// F2: layout B, pass B, store to buf1.
template <class K>
TFHE_HD void phase_F2(TeamRegs<K> &r, uint32_t jbB, const PrimeTab &pt, const uint32_t *buf0, uint32_t *buf1) {
    using C = typename K::Ntt;
    load_B<C>(r.x, buf0, jbB);
    fwd_pass_B<C>(r.x, r.twB, pt.q, pt.zero);
    store_B<C>(r.x, buf1, jbB);
}
// F3a: layout C, pass C.   F3b: multiply-accumulate against GGSW row g_row = [P][N] words (bsk_slot order;
// shared memory when staged by TMA, else global).
template <class K>
TFHE_HD void phase_F3a(TeamRegs<K> &r, uint32_t t, const PrimeTab &pt, const TwTables &tw, const uint32_t *buf1) {
    using C = typename K::Ntt;
    if constexpr (K::TWC_GLOBAL) {
        uint2 twc[C::NC_TW];
#pragma unroll
        for (int i = 0; i < C::NC_TW; i++) twc[i] = ld_global_tw(tw.fwdC + t * C::NC_TW + i);
        load_C<C>(r.x, buf1, t);
        fwd_pass_C<C>(r.x, twc, pt.q, pt.zero);
    } else {
        load_C<C>(r.x, buf1, t);
        fwd_pass_C<C>(r.x, r.twC, pt.q, pt.zero);
    }
}
template <class K, bool SHARED_G>
TFHE_HD void phase_F3b(TeamRegs<K> &r, uint32_t t, const uint32_t *g_row) {
    using C = typename K::Ntt;
#pragma unroll
    for (int c = 0; c < K::P; c++) {
#pragma unroll
        for (int ch = 0; ch < K::E / 4; ch++) {
            uint32_t g[4];
            const uint32_t *src = g_row + c * K::N + ((ch * C::T + t) << 2);
            if constexpr (SHARED_G) ld_vec<4>(g, src);
            else ld_global4(g, src);
#pragma unroll
            for (int w = 0; w < 4; w++) r.acc[c][4 * ch + w] += (uint64_t)r.x[4 * ch + w] * g[w];
        }
    }
}
// I1: reduce accumulator column c to [0,2q), inverse pass C, store (layout C) to buf0.
template <class K>
TFHE_HD void phase_I1(TeamRegs<K> &r, uint32_t t, uint32_t jbB, int c, const PrimeTab &pt, const TwTables &tw, uint32_t *buf0) {
    using C = typename K::Ntt;
    prefetch_twB<K>(r, tw.invB, jbB);
    uint2 twc[C::NC_TW];
#pragma unroll
    for (int i = 0; i < C::NC_TW; i++) twc[i] = ld_global_tw(tw.invC + t * C::NC_TW + i);
    static_for<0, K::P>([&](auto ci) {
        constexpr int cc = decltype(ci)::value;
        if (c == cc) {
#pragma unroll
            for (int e = 0; e < K::E; e++) r.x[e] = reduce_acc64_rt(r.acc[cc][e], pt);
        }
    });
    inv_pass_C<C>(r.x, twc, pt.q, pt.zero);
    store_C<C>(r.x, buf0, t);
}
template <class K>
TFHE_HD void phase_I2(TeamRegs<K> &r, uint32_t jbB, const PrimeTab &pt, const uint32_t *buf0, uint32_t *buf1) {
    using C = typename K::Ntt;
    load_B<C>(r.x, buf0, jbB);
    inv_pass_B<C>(r.x, r.twB, pt.q, pt.zero);
    store_B<C>(r.x, buf1, jbB);
}
// I3: layout A, inverse pass A, store residues in natural order (unpadded) to res_c[N].
template <class K>
TFHE_HD void phase_I3(TeamRegs<K> &r, uint32_t t, const PrimeTab &pt, const uint32_t *buf1, uint32_t *res_c) {
    using C = typename K::Ntt;
    load_A<C>(r.x, buf1, t);
    inv_pass_A<C>(r.x, pt.invA, pt.q, pt.zero);
#pragma unroll
    for (int e = 0; e < K::E; e++) res_c[(e << C::LOGT) | t] = r.x[e];
}

// ---- phase R: CRT + accumulate (all THREADS threads) --------------------------------------------
// res = [2 primes][P][N]; out[c][j] = base[c][j] + lift(res)  (ggsw.rs:175 `res += glwe_ciphertext0`)
template <class K>
TFHE_HD void phase_crt(uint32_t tid, const uint32_t *res, uint32_t *acc) {
#pragma unroll(kDrUnroll)
    for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::THREADS)
        acc[idx] += crt_to_u32(res[idx], res[K::P * K::N + idx]);
}

// ---- one-off key transform (K0): raw GGSW polynomial -> NTT domain, pre-scaled by N^-1 ----------
// u32 word -> centred representative in [-2^31, 2^31) -> residue in [0, q)
TFHE_HD uint32_t centred_residue(uint32_t g, uint32_t q) {
    const bool neg = (g >> 31) != 0;
    const uint32_t a = neg ? 0u - g : g;  // |g| as u32 (2^31 maps to itself)
    const uint32_t m = a % q;
    return neg ? (m ? q - m : 0u) : m;
}
// T1: load polynomial g[N] (natural order), lift, pass A, store to buf0.   (T2 == phase_F2)
template <class K>
TFHE_HD void phase_T1(TeamRegs<K> &r, uint32_t t, uint32_t jbB, const PrimeTab &pt, const TwTables &tw, const uint32_t *g, uint32_t *buf0) {
    using C = typename K::Ntt;
    prefetch_twB<K>(r, tw.fwdB, jbB);
#pragma unroll
    for (int e = 0; e < K::E; e++) r.x[e] = centred_residue(g[(e << C::LOGT) | t], pt.q);
    fwd_pass_A<C>(r.x, pt.fwdA, pt.q, pt.zero);
    store_A<C>(r.x, buf0, t);
}
// T3: pass C, scale by N^-1 (so the inverse NTT needs no final scaling), reduce to [0,q), store in
// bsk_slot order.
template <class K>
TFHE_HD void phase_T3(TeamRegs<K> &r, uint32_t t, const PrimeTab &pt, const TwTables &tw, const uint32_t *buf1, uint32_t *out) {
    phase_F3a<K>(r, t, pt, tw, buf1);
    using C = typename K::Ntt;
#pragma unroll
    for (int e = 0; e < K::E; e++) out[bsk_slot<C>(t, e)] = csub(shoup_mul(r.x[e], pt.ninv, pt.ninv_s, pt.q), pt.q);
}

}  // namespace tfhe
