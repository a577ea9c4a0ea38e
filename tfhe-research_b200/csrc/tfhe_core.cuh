// tfhe_core.cuh -- arithmetic core of the B200 PBS path (header only, __host__ __device__).
//
// Everything the hot kernels compute per thread lives here so that the SAME code can be stepped on
// the CPU by tests/emu (a lock-step emulation of one CTA) -- there is no GPU in the build container.
//
// Exact negacyclic products mod 2^32 (reference: utils.rs:113-173 teoplitz/poly_mul/poly_dot_product)
// are computed with a 2-prime RNS number-theoretic transform:
//   * signed gadget digits d (|d| <= B, decomposer.rs:42-80) and centred BSK words g in [-2^31,2^31)
//     give an integer result |sum| <= (k+1)*l*N*B*2^31 < Q0*Q1/2, so CRT reconstruction followed by
//     reduction mod 2^32 reproduces the reference's wrapping u32 result bit for bit.
//   * both primes are < 2^32/26, so Harvey/Shoup butterflies run WITHOUT any conditional correction
//     for up to 11 stages (values stay < 26q < 2^32): a forward butterfly is 3 IMAD-class + 2 ALU ops.
#pragma once
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define TFHE_HD __host__ __device__ __forceinline__
#else
#define TFHE_HD inline
#endif

#if !defined(__CUDA_ARCH__) && defined(TFHE_EMU_CHECKS)
#include <assert.h>
#define TFHE_CHECK(c) assert(c)
#else
#define TFHE_CHECK(c) ((void)0)
#endif

#ifndef TFHE_MULHI_WIDE
#define TFHE_MULHI_WIDE 0
#endif
#ifndef TFHE_ADD3
#define TFHE_ADD3 1
#endif

#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };  // host-only stand-in (tests/emu)
#endif

namespace tfhe {

// ---------------------------------------------------------------- primes and compile-time tables
// Two largest primes p == 1 (mod 4096) below 2^32/26 (so N <= 2048); Q0 < Q1.  psi4096 = primitive
// 4096-th root of unity (found offline; checked by tests: psi^2048 == -1).
constexpr uint32_t kQ0 = 165093377u;
constexpr uint32_t kQ1 = 165142529u;
constexpr uint32_t kPsi4096_0 = 18065781u;
constexpr uint32_t kPsi4096_1 = 152904259u;
constexpr int kMaxLogN = 11;

constexpr uint32_t prime_c(int pr) { return pr == 0 ? kQ0 : kQ1; }
constexpr uint32_t psi4096_c(int pr) { return pr == 0 ? kPsi4096_0 : kPsi4096_1; }
constexpr uint32_t mulmod_c(uint32_t a, uint32_t b, uint32_t q) { return (uint32_t)((uint64_t)a * b % q); }
constexpr uint32_t powmod_c(uint32_t a, uint64_t e, uint32_t q) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mulmod_c(r, a, q);
        a = mulmod_c(a, a, q);
        e >>= 1;
    }
    return r;
}
constexpr uint32_t invmod_c(uint32_t a, uint32_t q) { return powmod_c(a, q - 2, q); }
constexpr uint32_t shoup_c(uint32_t w, uint32_t q) { return (uint32_t)(((uint64_t)w << 32) / q); }
constexpr uint32_t brv_c(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}
// primitive 2N-th root for N = 2^logn
constexpr uint32_t psi2n_c(int pr, int logn) { return powmod_c(psi4096_c(pr), 1ull << (kMaxLogN - logn), prime_c(pr)); }
// Forward (Cooley-Tukey, natural in -> bit-reversed out) twiddle of stage s (0-based), block b in [0,2^s):
//   w(s,b) = psi^{(N >> (s+1)) * (2*brv_s(b)+1)}      (== psi_rev[2^s + b] of the usual merged NTT)
constexpr uint32_t fwd_tw_c(int pr, int logn, int s, uint32_t b) {
    return powmod_c(psi2n_c(pr, logn), (uint64_t)((1u << logn) >> (s + 1)) * (2 * brv_c(b, s) + 1), prime_c(pr));
}
constexpr uint32_t inv_tw_c(int pr, int logn, int s, uint32_t b) { return invmod_c(fwd_tw_c(pr, logn, s, b), prime_c(pr)); }

constexpr uint64_t kQ0Q1 = (uint64_t)kQ0 * kQ1;
constexpr uint64_t kHalfQ0Q1 = (kQ0Q1 - 1) / 2;
constexpr uint32_t kQ0InvModQ1 = invmod_c(kQ0 % kQ1, kQ1);
constexpr uint32_t kQ0InvModQ1_s = shoup_c(kQ0InvModQ1, kQ1);

template <int PR>
struct Prime {
    static constexpr uint32_t q = prime_c(PR);
    static constexpr uint32_t two_q = 2u * q;
    static constexpr uint32_t c32 = (uint32_t)((1ull << 32) % q);  // 2^32 mod q
    static constexpr uint32_t c32_s = shoup_c(c32, q);
    static constexpr uint32_t one_s = shoup_c(1u, q);               // floor(2^32/q)
};

// ---------------------------------------------------------------- scalar helpers
TFHE_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && TFHE_MULHI_WIDE
    // A/B switch: high word via IMAD.WIDE instead of IMAD.HI.  Measured on B200 both issue at HALF the plain
    // IMAD rate (ncu: fma-heavy pipe cycles; IMAD.HI 8.9 vs IMAD 18.5 T lane-ops/s) and perform the same,
    // so the default stays mul.hi (no 64-bit register pair).
    uint32_t lo, hi;
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    (void)lo;
    return hi;
#elif defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
TFHE_HD uint32_t umin_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// x*w mod q, lazily: result in [0, 2q) for ANY x < 2^32 (w < q, ws = floor(w*2^32/q)).
TFHE_HD uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t ws, uint32_t q) {
    uint32_t hi = mulhi_u32(x, ws);
    return x * w - hi * q;
}
// [0,2q) -> [0,q)
TFHE_HD uint32_t csub(uint32_t x, uint32_t q) { return umin_u32(x, x - q); }

// 32-bit add that ptxas must emit on the ALU pipe: a third addend `z` that is zero at run time but opaque at
// compile time (a kernel-parameter word) makes it a genuine 3-input IADD3; a plain 2-input add is often
// emitted as IMAD.IADD, which put ~15% more work on the binding FMA-heavy pipe (profiles/r01_v2_*).
TFHE_HD uint32_t add_alu(uint32_t a, uint32_t b, uint32_t z) {
#if TFHE_ADD3
    return a + b + z;
#else
    (void)z;
    return a + b;
#endif
}

// Cooley-Tukey butterfly, no correction: inputs < B*q  ->  outputs < (B+2)*q.
TFHE_HD void ct_bfly(uint32_t &X, uint32_t &Y, uint32_t w, uint32_t ws, uint32_t q, uint32_t z = 0) {
    uint32_t T = shoup_mul(Y, w, ws, q);
    TFHE_CHECK((uint64_t)X + T < (1ull << 32) && (uint64_t)X + 2ull * q < (1ull << 32));
    Y = X - T + 2u * q;
    X = add_alu(X, T, z);
}
// Gentleman-Sande butterfly, inputs and outputs in [0, 2q).
TFHE_HD void gs_bfly(uint32_t &X, uint32_t &Y, uint32_t w, uint32_t ws, uint32_t q, uint32_t z = 0) {
    TFHE_CHECK(X < 2u * q && Y < 2u * q);
    uint32_t S = add_alu(X, Y, z);
    uint32_t D = X - Y + 2u * q;
    X = umin_u32(S, S - 2u * q);
    Y = shoup_mul(D, w, ws, q);
}

template <int B, int E, class F>
TFHE_HD void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---------------------------------------------------------------- reference scalar semantics
// utils.rs:13-33 switch_modulus to 2N = 2^(logn+1): round-half-up(v / 2^(31-logn)) mod 2N
TFHE_HD uint32_t mod_switch(uint32_t v, int logn) {
    const int sh = 31 - logn;  // log2 divisor
    uint32_t r = (v >> sh) + ((v >> (sh - 1)) & 1u);
    return r & ((2u << logn) - 1u);
}

// decomposer.rs:27-80.  Returns the `L` most significant signed digits (out[0] = most significant)
// with the reference's quirk: window == B-1 plus an incoming carry yields digit +B and NO carry
// (SURVEY 9-B H3), so digits lie in {-B/2..B/2-1} u {B}.
template <int LOGB, int L>
TFHE_HD void decompose_signed(uint32_t v, int32_t *out) {
    constexpr int IG = 32 - LOGB * L;
    uint32_t r = v;
    if constexpr (IG > 0) r = ((v >> IG) + ((v >> (IG - 1)) & 1u)) << IG;  // round_value, wraps at the top
    constexpr uint32_t MASK = (1u << LOGB) - 1u, HALF = 1u << (LOGB - 1);
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < L; j++) {  // windows below IG are zero and produce no carry
        uint32_t t = ((r >> (IG + LOGB * j)) & MASK) + carry;
        uint32_t cm = t & HALF;
        out[L - 1 - j] = (int32_t)t - (int32_t)(cm << 1);
        carry = cm >> (LOGB - 1);
    }
}

// utils.rs:183-207: coefficient j of p * X^a for a in [0, 2N) (negacyclic), p has N = 2^logn entries
TFHE_HD uint32_t rot_coeff(const uint32_t *p, uint32_t j, uint32_t a, int logn) {
    const uint32_t N = 1u << logn;
    uint32_t idx = (j - a) & (2u * N - 1u);
    uint32_t v = p[idx & (N - 1u)];
    return (idx & N) ? 0u - v : v;
}

// CRT of (r0 mod Q0, r1 mod Q1), both lazily in [0,2q): centred lift, returned mod 2^32.
TFHE_HD uint32_t crt_to_u32(uint32_t r0, uint32_t r1) {
    r0 = csub(r0, kQ0);
    r1 = csub(r1, kQ1);
    uint32_t d = r1 - r0;            // r0 < Q0 < Q1
    d = umin_u32(d, d + kQ1);        // wrapped negative -> + Q1
    uint32_t t = csub(shoup_mul(d, kQ0InvModQ1, kQ0InvModQ1_s, kQ1), kQ1);
    uint64_t x = (uint64_t)r0 + (uint64_t)kQ0 * t;  // in [0, Q0*Q1)
    uint32_t lo = (uint32_t)x;
    return x > kHalfQ0Q1 ? lo - (uint32_t)kQ0Q1 : lo;
}

// 64-bit lazy accumulator -> [0, 2q)
template <int PR>
TFHE_HD uint32_t reduce_acc64(uint64_t a) {
    using P = Prime<PR>;
    uint32_t hi = (uint32_t)(a >> 32), lo = (uint32_t)a;
    uint32_t r = shoup_mul(hi, P::c32, P::c32_s, P::q) + shoup_mul(lo, 1u, P::one_s, P::q);  // < 4q
    return umin_u32(r, r - P::two_q);
}

// ---------------------------------------------------------------- team NTT geometry
// One polynomial of N = 2^LOGN coefficients is transformed by a TEAM of T = N/E threads, E = 2^LOGE
// coefficients per thread in registers.  Index bits of j (MSB first):  hA[LOGE] | mid[QB] | lo[LOGE].
//   layout A: regs <-> hA          thread t  <-> (mid|lo)            j = (e << LOGT) | t
//   layout B: regs <-> mid|loX     thread    <-> hA + loR            (loX = lowest XB bits of lo)
//   layout C: regs <-> lo          thread t  <-> (hA|mid)            j = (t << LOGE) | e
// Forward: pass A (stages 0..LOGE-1, compile-time twiddles) -> B (QB stages) -> C (LOGE stages);
// inverse (Gentleman-Sande) runs C -> B -> A.  Exchanges go through a padded shared-memory buffer:
// phys(j) = j + 4*(j >> 5)  (additive over disjoint bit sets => compile-time register offsets, and
// conflict-free for all three layouts with the thread-bit orders below; see tools/bank_check.py).
template <int LOGN_, int LOGE_>
struct NttCfg {
    static constexpr int LOGN = LOGN_, LOGE = LOGE_;
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, LOGT = LOGN - LOGE, T = 1 << LOGT;
    static constexpr int QB = LOGN - 2 * LOGE;
    static constexpr int XB = LOGE - QB;
    static constexpr int NB_TW = (1 << QB) - 1;  // per-thread twiddles of pass B
    static constexpr int NC_TW = E - 1;          // per-thread twiddles of pass C
    static constexpr int NPAD = N + 4 * (N >> 5);
    static_assert(QB >= 1 && QB <= LOGE, "unsupported (LOGN, LOGE)");
    static_assert(LOGT >= 5, "team must be at least one warp");
    static constexpr int VB = 1 << XB;  // contiguous words per access in layout B (1, 2 or 4)
    static_assert(XB <= 2, "layout B vector width");
};

TFHE_HD constexpr uint32_t phys_c(uint32_t j) { return j + ((j >> 5) << 2); }

// j-bit position taken by bit i of the team-thread id in layout B: loR bits then hA bits, ascending,
// except (LOGN,LOGE)=(10,4) where {2,3,7,6,8,9} keeps the 8-lane LDS.128 phases conflict-free.
template <class C>
TFHE_HD constexpr int bperm_c(int i) {
    if (C::LOGN == 10 && C::LOGE == 4) {
        constexpr int p[6] = {2, 3, 7, 6, 8, 9};
        return p[i];
    }
    return i < C::QB ? C::XB + i : (C::LOGN - C::LOGE) + (i - C::QB);
}
template <class C>
TFHE_HD uint32_t jbase_B(uint32_t t) {
    uint32_t j = 0;
#pragma unroll
    for (int i = 0; i < C::LOGT; i++) j |= ((t >> i) & 1u) << bperm_c<C>(i);
    return j;
}
// register index eB -> its j bits in layout B
template <class C>
TFHE_HD constexpr uint32_t jreg_B(uint32_t eB) {
    return ((eB >> C::XB) << C::LOGE) | (eB & ((1u << C::XB) - 1u));
}

// ---- exchanges (buf = team-private padded buffer of NPAD words) ----
template <class C>
TFHE_HD void store_A(const uint32_t *x, uint32_t *buf, uint32_t t) {
    uint32_t *b = buf + phys_c(t);
    static_for<0, C::E>([&](auto ei) {
        constexpr uint32_t e = decltype(ei)::value;
        b[phys_c(e << C::LOGT)] = x[e];
    });
}
template <class C>
TFHE_HD void load_A(uint32_t *x, const uint32_t *buf, uint32_t t) {
    const uint32_t *b = buf + phys_c(t);
    static_for<0, C::E>([&](auto ei) {
        constexpr uint32_t e = decltype(ei)::value;
        x[e] = b[phys_c(e << C::LOGT)];
    });
}

template <int V>
struct VecT;
template <>
struct VecT<1> { using type = uint32_t; };
template <>
struct VecT<2> { struct alignas(8) type { uint32_t v[2]; }; };
template <>
struct VecT<4> { struct alignas(16) type { uint32_t v[4]; }; };

template <int V>
TFHE_HD void ld_vec(uint32_t *dst, const uint32_t *src) {
    using VT = typename VecT<V>::type;
    VT tmp = *reinterpret_cast<const VT *>(src);
    const uint32_t *p = reinterpret_cast<const uint32_t *>(&tmp);
#pragma unroll
    for (int i = 0; i < V; i++) dst[i] = p[i];
}
template <int V>
TFHE_HD void st_vec(uint32_t *dst, const uint32_t *src) {
    using VT = typename VecT<V>::type;
    VT tmp;
    uint32_t *p = reinterpret_cast<uint32_t *>(&tmp);
#pragma unroll
    for (int i = 0; i < V; i++) p[i] = src[i];
    *reinterpret_cast<VT *>(dst) = tmp;
}

template <class C>
TFHE_HD void store_B(const uint32_t *x, uint32_t *buf, uint32_t jbB) {
    uint32_t *b = buf + phys_c(jbB);
    static_for<0, (C::E >> C::XB)>([&](auto gi) {
        constexpr uint32_t eB = decltype(gi)::value << C::XB;
        st_vec<C::VB>(b + phys_c(jreg_B<C>(eB)), x + eB);
    });
}
template <class C>
TFHE_HD void load_B(uint32_t *x, const uint32_t *buf, uint32_t jbB) {
    const uint32_t *b = buf + phys_c(jbB);
    static_for<0, (C::E >> C::XB)>([&](auto gi) {
        constexpr uint32_t eB = decltype(gi)::value << C::XB;
        ld_vec<C::VB>(x + eB, b + phys_c(jreg_B<C>(eB)));
    });
}
template <class C>
TFHE_HD void store_C(const uint32_t *x, uint32_t *buf, uint32_t t) {
    uint32_t *b = buf + phys_c(t << C::LOGE);
    static_for<0, C::E / 4>([&](auto ci) {
        constexpr uint32_t c = decltype(ci)::value;
        st_vec<4>(b + 4 * c, x + 4 * c);
    });
}
template <class C>
TFHE_HD void load_C(uint32_t *x, const uint32_t *buf, uint32_t t) {
    const uint32_t *b = buf + phys_c(t << C::LOGE);
    static_for<0, C::E / 4>([&](auto ci) {
        constexpr uint32_t c = decltype(ci)::value;
        ld_vec<4>(x + 4 * c, b + 4 * c);
    });
}

// ---- register passes ----
// All passes take the prime q and their twiddles as RUNTIME values so that both RNS primes execute
// the same instructions (one copy of the unrolled code in the instruction cache; the v1 kernel with
// compile-time primes was bound by instruction fetch, profiles/r01_v1_*).  tw[2^u-1+m] = (w, ws).
//
// Pass over register bits [LO, LO+Q): stage u pairs bit LO+Q-1-u, twiddle index = top u bits of the field.
template <int E, int LO, int Q>
TFHE_HD void fwd_pass_bits(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) {
    static_for<0, Q>([&](auto ui) {
        constexpr int u = decltype(ui)::value;
        constexpr int bit = 1 << (LO + Q - 1 - u);
        static_for<0, E>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) {
                constexpr int field = (e >> LO) & ((1 << Q) - 1);
                constexpr int m = field >> (Q - u);
                const uint2 w = tw[(1 << u) - 1 + m];
                ct_bfly(x[e], x[e + bit], w.x, w.y, q, z);
            }
        });
    });
}
template <int E, int LO, int Q>
TFHE_HD void inv_pass_bits(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) {
    static_for<0, Q>([&](auto ui) {
        constexpr int u = Q - 1 - decltype(ui)::value;
        constexpr int bit = 1 << (LO + Q - 1 - u);
        static_for<0, E>([&](auto ei) {
            constexpr int e = decltype(ei)::value;
            if constexpr ((e & bit) == 0) {
                constexpr int field = (e >> LO) & ((1 << Q) - 1);
                constexpr int m = field >> (Q - u);
                const uint2 w = tw[(1 << u) - 1 + m];
                gs_bfly(x[e], x[e + bit], w.x, w.y, q, z);
            }
        });
    });
}
// pass A: stages 0..LOGE-1 on the hA bits (all E register bits); twA[2^s-1+b] = w(s, b), the same for
//         every thread (kept in the kernel-parameter constant bank).
// pass B: stages LOGE..LOGE+QB-1 on the mid bits (register bits [XB, XB+QB)); twB = w(LOGE+u, (hA<<u)|m).
// pass C: stages LOGE+QB..LOGN-1 on the lo bits; twC = w(LOGE+QB+u, (t<<u)|m).
template <class C> TFHE_HD void fwd_pass_A(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { fwd_pass_bits<C::E, 0, C::LOGE>(x, tw, q, z); }
template <class C> TFHE_HD void fwd_pass_B(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { fwd_pass_bits<C::E, C::XB, C::QB>(x, tw, q, z); }
template <class C> TFHE_HD void fwd_pass_C(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { fwd_pass_bits<C::E, 0, C::LOGE>(x, tw, q, z); }
template <class C> TFHE_HD void inv_pass_A(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { inv_pass_bits<C::E, 0, C::LOGE>(x, tw, q, z); }
template <class C> TFHE_HD void inv_pass_B(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { inv_pass_bits<C::E, C::XB, C::QB>(x, tw, q, z); }
template <class C> TFHE_HD void inv_pass_C(uint32_t *x, const uint2 *tw, uint32_t q, uint32_t z) { inv_pass_bits<C::E, 0, C::LOGE>(x, tw, q, z); }

// Per-prime constants handed to the kernels by value (kernel-parameter constant bank; a warp only
// ever reads its own prime's entry, so every access is a uniform LDC).
constexpr int kMaxPassA = 15;  // E-1 for E = 16
struct PrimeTab {
    uint32_t q, c32, c32_s, one_s;  // prime, 2^32 mod q, shoup(c32), floor(2^32/q)
    uint32_t ninv, ninv_s, zero, pad1;  // N^-1 mod q (key transform only); zero: run-time 0 for add_alu
    uint2 fwdA[kMaxPassA + 1];
    uint2 invA[kMaxPassA + 1];
};
inline void fill_prime_tab(int pr, int logn, int loge, PrimeTab &t) {
    const uint32_t q = prime_c(pr);
    t.q = q;
    t.c32 = (uint32_t)((1ull << 32) % q);
    t.c32_s = shoup_c(t.c32, q);
    t.one_s = shoup_c(1u, q);
    t.ninv = invmod_c((1u << logn) % q, q);
    t.ninv_s = shoup_c(t.ninv, q);
    t.zero = t.pad1 = 0;
    for (int i = 0; i <= kMaxPassA; i++) t.fwdA[i] = t.invA[i] = uint2{0, 0};
    for (int s = 0; s < loge; s++)
        for (uint32_t b = 0; b < (1u << s); b++) {
            const uint32_t w = fwd_tw_c(pr, logn, s, b), wi = invmod_c(w, q);
            t.fwdA[(1 << s) - 1 + b] = uint2{w, shoup_c(w, q)};
            t.invA[(1 << s) - 1 + b] = uint2{wi, shoup_c(wi, q)};
        }
}

// 64-bit lazy accumulator -> [0, 2q), runtime prime
TFHE_HD uint32_t reduce_acc64_rt(uint64_t a, const PrimeTab &p) {
    uint32_t hi = (uint32_t)(a >> 32), lo = (uint32_t)a;
    uint32_t r = shoup_mul(hi, p.c32, p.c32_s, p.q) + shoup_mul(lo, 1u, p.one_s, p.q);  // < 4q
    return umin_u32(r, r - 2u * p.q);
}

// Position (in the transformed, bit-reversed domain) -> slot in the stored BSK row: thread t of
// layout C reads chunk c (4 words) at word (c*T + t)*4, i.e. consecutive lanes read consecutive 16 B.
template <class C>
TFHE_HD constexpr uint32_t bsk_slot(uint32_t t, uint32_t e) {
    return (((e >> 2) * C::T + t) << 2) | (e & 3u);
}

}  // namespace tfhe
