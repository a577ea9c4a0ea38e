// kernels_fft.cuh -- sm_100a kernels of the FP64-FFT arithmetic path (see fft_team.cuh for the arithmetic).
//
//   pbs_fft_kernel        blind rotation: ONE CTA = CTS ciphertexts; each ciphertext is a team of P = k+1 sub-teams
//                         of T threads (sub-team s: digits + forward transforms of polynomial s, multiply-accumulate +
//                         inverse transforms of result column s); the GLWE accumulator of every team is resident in
//                         shared memory for all n CMUX steps; thread 0 streams the FFT-domain bootstrapping key
//                         global -> shared with TMA bulk copies into a ring (full/empty mbarriers).  Every key byte
//                         fetched from L2 is used by all CTS ciphertexts of the CTA.  Also runs one external product /
//                         CMUX for the sub-operation entry points.
//   bsk_fft_transform_kernel   raw BSK -> limb-split FFT domain (one-off at key upload)
//
// Reference: bootstrapping.rs:58-105 (blind rotation), ggsw.rs:132-178 (external product, cmux).
#pragma once
#include "fft_team.cuh"
#include "kernels.cuh"

namespace tfhe {
namespace fft {

struct FftArgs {
    TwTablesF tw;
    const cplx *bsk_fft;       // [n][ROWS][2 limbs][P][M] slot order, pre-scaled by 1/M
    const uint32_t *lwe_in;    // mode 0: [B][n+1]
    const uint32_t *luts;      // [T][N] unencoded
    const uint32_t *lut_idx;   // [B] or nullptr
    const uint32_t *in0, *in1; // mode 1 (out = ExtProd(G, in0)) / mode 2 (out = ExtProd(G, in1-in0)+in0): [B][P][N]
    const uint32_t *ggsw_index;
    uint32_t *glwe_out;        // [B][P][N]
    uint32_t *err_flag;        // bit 0: test-vector entry >= 2^log_p (glwe.rs:144); bit 1: TMA wait timed out
    unsigned long long *margin;  // largest |x - rint(x)| before rounding (bits of a non-negative double), CHECK only
    uint32_t n, batch, mode, log_p, enc_shift;
};

// TFHE_FFT_ABLATE (bit mask, measurement only -- results are WRONG when set): 1 = constant digits instead of the
// decomposition, 2 = no multiply-accumulate, 4 = no team barriers, 16 = no inverse transforms.  Used to attribute the
// step time of the latency-bound kernel to its phases (profiles/r01_fft_ablation.log).
#ifndef TFHE_FFT_ABLATE
#define TFHE_FFT_ABLATE 0
#endif
__device__ __forceinline__ void team_bar_id(uint32_t id, int nthreads) {
#if (TFHE_FFT_ABLATE & 4)
    __syncwarp();
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {   // non-blocking
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- key relay through tensor memory (TMEM)
// The multiply-accumulate is bound by shared-memory bandwidth, and every team of a CTA reads the same key row from the
// TMA ring.  With the relay ONE team (rotating per row) reads a row from shared memory and also parks it in TMEM
// (tcgen05.st, 32x32b: lane = the thread's lane in its warp's TMEM quadrant, 8 columns per complex pair); the other teams
// fetch it with tcgen05.ld, which does not touch the shared-memory pipe.  A team is 4 warps, so warp w of every team owns
// the same TMEM quadrant w and the thread that stored a value and the threads that load it sit on the same TMEM lane.
// Shared-memory key reads drop to 1/CTS; the bits are the same (a copy), so results do not change.
#ifndef TFHE_FFT_TMEM
#define TFHE_FFT_TMEM 0
#endif
#ifndef TFHE_FFT_TMEM_SLOTS
#define TFHE_FFT_TMEM_SLOTS 4
#endif
#ifndef TFHE_FFT_TMEM_CH
#define TFHE_FFT_TMEM_CH 2     // points per tcgen05.ld (8 x 32-bit columns each)
#endif
template <class K, bool BMMP>
struct KeyRelay {
    static constexpr bool ON = (TFHE_FFT_TMEM != 0) && !BMMP && K::TEAM_THREADS == 128 && K::CTS > 1;
    static constexpr int NT = TFHE_FFT_TMEM_SLOTS, SLOT_COLS = K::EH * 8, CH = TFHE_FFT_TMEM_CH;
    static constexpr int need = NT * SLOT_COLS;
    static constexpr int COLS = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    static_assert(need <= 512, "TMEM has 512 columns");
    static_assert(K::EH % CH == 0 && (CH == 1 || CH == 2 || CH == 4), "chunk = 8, 16 or 32 columns");
    // after the ring's mbarriers: tfull[NT][4 quadrants], tempty[NT][4], TMEM base address
    static constexpr int BYTES = ON ? 2 * NT * 4 * 8 + 16 : 0;
};
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {   // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one point of a key row, both limbs: 8 columns
__device__ __forceinline__ void tmem_st_pair(uint32_t taddr, const cplx a, const cplx b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__double2loint(a.re)),
                 "r"(__double2hiint(a.re)), "r"(__double2loint(a.im)), "r"(__double2hiint(a.im)), "r"(__double2loint(b.re)),
                 "r"(__double2hiint(b.re)), "r"(__double2loint(b.im)), "r"(__double2hiint(b.im))
                 : "memory");
}
// load + wait in ONE statement: the outputs are defined only once the data has arrived
template <int NW>
__device__ __forceinline__ void tmem_ld_wait(uint32_t taddr, uint32_t (&w)[NW]) {
    if constexpr (NW == 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(taddr)
                     : "memory");
    if constexpr (NW == 16)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                     : "r"(taddr)
                     : "memory");
    if constexpr (NW == 32)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                     : "r"(taddr)
                     : "memory");
}
__device__ __forceinline__ cplx cplx_from_words(const uint32_t *w) {
    return cplx{__hiloint2double((int)w[1], (int)w[0]), __hiloint2double((int)w[3], (int)w[2])};
}
__device__ __forceinline__ void mac_point(cplx &acc0, cplx &acc1, const cplx xv, const cplx a, const cplx b) {
    acc0.re = fma_d(-xv.im, a.im, fma_d(xv.re, a.re, acc0.re));
    acc0.im = fma_d(xv.im, a.re, fma_d(xv.re, a.im, acc0.im));
    acc1.re = fma_d(-xv.im, b.im, fma_d(xv.re, b.re, acc1.re));
    acc1.im = fma_d(xv.im, b.re, fma_d(xv.re, b.im, acc1.im));
}
// relay team: key row from the shared-memory slot, parked in TMEM on the way (PUBLISH) and multiplied (DOMAC)
template <class K, bool OWN, bool PUBLISH, bool DOMAC>
__device__ __forceinline__ void mac_row_relay(FftRegs<K> &r, uint32_t t, uint32_t col, const cplx *slot, const cplx *xbuf, uint32_t h, uint32_t taddr) {
    const cplx *g0 = slot + col * K::MH + t, *g1 = g0 + K::P * K::MH;
    static_for<0, K::HALVES>([&](auto hi) {
        constexpr int hh = decltype(hi)::value;
        if (h == (uint32_t)hh) {
#pragma unroll
            for (int q = 0; q < K::EH; q++) {
                constexpr int e0 = hh * K::EH;
                const cplx a = g0[q * K::T], b = g1[q * K::T];
                if constexpr (PUBLISH) tmem_st_pair(taddr + q * 8, a, b);
                if constexpr (DOMAC) {
                    const cplx xv = OWN ? r.x[e0 + q] : xbuf[(e0 + q) * K::T + t];
                    mac_point(r.acc[0][e0 + q], r.acc[1][e0 + q], xv, a, b);
                }
            }
        }
    });
}
// the other teams: key row from TMEM
template <class K, bool OWN, int CH>
__device__ __forceinline__ void mac_row_tmem(FftRegs<K> &r, uint32_t t, const cplx *xbuf, uint32_t h, uint32_t taddr) {
    static_for<0, K::HALVES>([&](auto hi) {
        constexpr int hh = decltype(hi)::value;
        if (h == (uint32_t)hh) {
#pragma unroll
            for (int c = 0; c < K::EH / CH; c++) {
                constexpr int e0 = hh * K::EH;
                uint32_t w[CH * 8];
                tmem_ld_wait<CH * 8>(taddr + c * CH * 8, w);
#pragma unroll
                for (int u = 0; u < CH; u++) {
                    const int q = c * CH + u;
                    const cplx xv = OWN ? r.x[e0 + q] : xbuf[(e0 + q) * K::T + t];
                    mac_point(r.acc[0][e0 + q], r.acc[1][e0 + q], xv, cplx_from_words(w + u * 8), cplx_from_words(w + u * 8 + 4));
                }
            }
        }
    });
}

// BMMP = true: blind rotation unrolled by two (notes/BMMP Bootstrapping.md): n/2 steps, each consumes the three GGSWs of
// a key triple (3 slots per row), decomposes acc itself and adds ExtProd(bundle, acc); blind rotation mode only.
template <class K, bool BMMP = false>
__global__ void __launch_bounds__(K::THREADS, 1) pbs_fft_kernel(const __grid_constant__ FftArgs a) {
    using C = typename K::F;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t team = tid / K::TEAM_THREADS, tt = tid % K::TEAM_THREADS, sub = tt / K::T, t = tt % K::T;
    const uint32_t team_bytes = (uint32_t)K::team_bytes((int)a.n);
    uint8_t *ring = smem + K::CTS * team_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES), *empty = full + K::NSLOT;
    using RL = KeyRelay<K, BMMP>;   // key rows relayed through TMEM (see above)
#ifndef TFHE_FFT_OWNFIRST
#define TFHE_FFT_OWNFIRST 1
#endif
#ifndef TFHE_FFT_SELFREFILL
#define TFHE_FFT_SELFREFILL 1
#endif
    // OWN_FIRST: see the level loop.  SELF_REFILL: no producer thread -- the warp whose arrival frees a ring slot issues
    // the TMA copy of the row that reuses it, so a refill starts the moment the slot's last reader is done.
    constexpr bool OWN_FIRST = (TFHE_FFT_OWNFIRST != 0) && !RL::ON && !BMMP && !K::SINGLE_BUF && K::HALVES == 1 && K::NSLOT >= K::P && !(TFHE_FFT_ABLATE);
    constexpr bool SELF_REFILL = OWN_FIRST && (TFHE_FFT_SELFREFILL != 0);
#ifndef TFHE_FFT_P2PG
#define TFHE_FFT_P2PG 0
#endif
    constexpr bool P2PG = (TFHE_FFT_P2PG != 0) && !OWN_FIRST && !RL::ON && !(TFHE_FFT_ABLATE);
    // point-to-point row barriers of every team (OWN_FIRST / P2P below): pub[P], rd[P]
    uint64_t *p2p_base = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES + 2 * K::NSLOT * 8 + RL::BYTES);
    uint64_t *pub = p2p_base + team * 2 * K::P, *rd = pub + K::P;
    uint32_t *claimed = reinterpret_cast<uint32_t *>(p2p_base + K::CTS * 2 * K::P);   // [NSLOT] refills issued per ring slot
    uint64_t *tfull = empty + K::NSLOT, *tempty = tfull + RL::NT * 4;   // [TMEM slot][quadrant]
    uint32_t *tmem_word = reinterpret_cast<uint32_t *>(tempty + RL::NT * 4);

    const bool single = a.mode != 0;   // sub-operation entry points: one ciphertext (team 0) per CTA, its own GGSW
    // blind rotation: the batch is split over the grid as evenly as possible (CTA b gets base or base+1 ciphertexts,
    // base+1 <= CTS), so a partially filled last wave shortens every CTA instead of leaving SMs idle
    const uint32_t base = a.batch / gridDim.x, rem = a.batch % gridDim.x;
    const uint32_t ct0 = single ? blockIdx.x : blockIdx.x * base + min(blockIdx.x, rem);
    const uint32_t active = single ? 1u : base + (blockIdx.x < rem ? 1u : 0u);
    constexpr uint32_t KEYS = BMMP ? 3u : 1u;                 // GGSWs per step
    constexpr uint32_t STEP_SLOTS = K::SLOTS_PER_STEP * KEYS;
    const uint32_t n_steps = single ? 1u : (BMMP ? a.n / 2u : a.n);
    const uint32_t total_slots = n_steps * STEP_SLOTS;

    if (tid == 0) {
        for (int s = 0; s < K::NSLOT; s++) {
            mbar_init(full + s, 1);
            // every warp of every active team consumes every slot; with the relay only the relaying team's warps do
            mbar_init(empty + s, (RL::ON ? 1u : active) * K::P * K::WARPS_PER_SUB);
        }
        for (int s = 0; s < K::NSLOT; s++) claimed[s] = 0;
        for (int s = 0; s < K::CTS * 2 * K::P; s++)   // [team][pub[P], rd[P]]
            mbar_init(p2p_base + s, (uint32_t)K::WARPS_PER_SUB * ((s / K::P) % 2 == 0 ? 1u : (uint32_t)(K::P - 1)));
        if constexpr (RL::ON) {
            for (int s = 0; s < RL::NT * 4; s++) {
                mbar_init(tfull + s, 1);                               // the relaying warp of the quadrant
                mbar_init(tempty + s, active > 1u ? active - 1u : 1u);  // the same warp of every other active team
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tmem_base = 0;
    if constexpr (RL::ON) {
        if (tid < 32) {
            __syncwarp();
            tmem_alloc(tmem_word, RL::COLS);
        }
        tc_fence_before();
    }
    __syncthreads();   // the only CTA-wide barrier: teams are independent below
    if constexpr (RL::ON) {
        tc_fence_after();
        tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_word);
    }

    if (team >= active) return;
    const uint32_t ct = ct0 + team;
    uint8_t *tm = smem + team * team_bytes;
    uint32_t *acc = reinterpret_cast<uint32_t *>(tm + K::TM_ACC);
    uint8_t *sb = tm + K::TM_SUB + sub * K::SUB_BYTES;
    typename K::stash_t *stash = reinterpret_cast<typename K::stash_t *>(sb);
    cplx *buf0 = reinterpret_cast<cplx *>(sb + K::STASH_BYTES), *buf1 = K::SINGLE_BUF ? buf0 : buf0 + C::MPAD;
    uint16_t *at = reinterpret_cast<uint16_t *>(tm + K::TM_AT);
    // named barriers: one per team, plus one per sub-team unless a sub-team is a single warp (then __syncwarp)
    constexpr bool WARP_SUB = K::T == 32;
    const uint32_t team_bar = 1 + team * (WARP_SUB ? 1 : K::P + 1), sub_bar = team_bar + 1 + sub;
    static_assert(K::CTS * (WARP_SUB ? 1 : K::P + 1) <= 15, "named barrier ids");
    auto sub_sync = [&]() {
        if constexpr (WARP_SUB) __syncwarp();
        else team_bar_id(sub_bar, K::T);
    };
    const uint32_t jbB = jbase_B<C>(t);
    const cplx *twB = a.tw.twB + (t >> C::QB) * C::NB_TW;
    const cplx *twC = a.tw.twC;
#ifndef TFHE_FFT_TWREG
#define TFHE_FFT_TWREG 1
#endif
    // a thread's two twiddle-table entries (passes B and C) never change: fetched once, not after every barrier.  Measured
    // to pay only with the OWN_FIRST level loop (P1: 71.4 -> 69.8 ms); with the ring-order loop the 8 registers cost more
    // than the loads (P1: 74.8 -> 79.9 ms, and P0 / the BMMP variant alike), so those keep loading the entries per pass.
    constexpr bool TW_REG = (TFHE_FFT_TWREG != 0) && (TFHE_FFT_OWNFIRST != 0) && !BMMP && !K::SINGLE_BUF && K::HALVES == 1 && K::NSLOT >= K::P && !(TFHE_FFT_ABLATE) && !(TFHE_FFT_TMEM);
    cplx twB_base = {}, twC_base = {};
    if constexpr (TW_REG) {
        twB_base = pass_tw_base<C::QB>(twB, 1);
        twC_base = pass_tw_base<C::LOGE>(twC + t, C::T);
    }
    // operands of the decomposed difference  minuend(p, (j - rot)) - subtrahend(p, j)
    const uint32_t *mbase = acc, *sbase = acc;

    if (!single) {
        // utils.rs:23-33 mod switch of (a_0..a_{n-1}, b) to 2N
        const uint32_t *lwe = a.lwe_in + (size_t)ct * (a.n + 1);
        for (uint32_t i = tt; i <= a.n; i += K::TEAM_THREADS) at[i] = (uint16_t)mod_switch(__ldg(lwe + i), K::LOGN);
        team_bar_id(team_bar, K::TEAM_THREADS);
        // acc = trivial GLWE of the encoded test vector times X^{-b~}  (bootstrapping.rs:79-86)
        const uint32_t b = at[a.n];
        const uint32_t *lut = a.luts + (size_t)(a.lut_idx ? __ldg(a.lut_idx + ct) : 0u) * K::N;
        for (uint32_t idx = tt; idx < (uint32_t)(K::P * K::N); idx += K::TEAM_THREADS) {
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
    } else {
        // single-ciphertext modes borrow idle shared memory of the other teams (FftPbsCfg::SPARE_*) for the polynomial to
        // decompose and for a zero subtrahend, so the step below is the same code as the blind rotation with rot = 0
        uint32_t *din = reinterpret_cast<uint32_t *>(smem + 1 * team_bytes + K::SPARE_DIN);
        uint32_t *zero = reinterpret_cast<uint32_t *>(smem + K::SPARE_ZERO_TEAM * team_bytes + K::SPARE_ZERO);
        const uint32_t *x0 = a.in0 + (size_t)ct * K::P * K::N, *x1 = a.in1 + (size_t)ct * K::P * K::N;
        for (uint32_t idx = tt; idx < (uint32_t)(K::P * K::N); idx += K::TEAM_THREADS) {
            const uint32_t v0 = __ldg(x0 + idx);
            acc[idx] = a.mode == 2 ? v0 : 0u;
            din[idx] = a.mode == 2 ? __ldg(x1 + idx) - v0 : v0;
            zero[idx] = 0u;
        }
        mbase = din;
        sbase = zero;
    }
    team_bar_id(team_bar, K::TEAM_THREADS);

    FftRegs<K> R;
    double maxfrac = 0.0;
#ifndef TFHE_FFT_ACCREG
#define TFHE_FFT_ACCREG 0
#endif
    // this thread's 2E words of acc[sub] (it decomposes the same ones it updates) cross the step boundary in registers
    constexpr bool ACC_REG = (TFHE_FFT_ACCREG != 0) && !K::SINGLE_BUF;
    uint32_t accv[2 * K::E];
    if constexpr (ACC_REG) {
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            const uint32_t j = ((uint32_t)e << C::LOGT) | t;
            accv[2 * e] = sbase[sub * K::N + j];
            accv[2 * e + 1] = sbase[sub * K::N + j + K::M];
        }
    }
    uint32_t it = 0;   // position in the key stream (the same sequence in every warp)
    uint32_t lv = 0;   // levels this team has run (phase of its point-to-point row barriers)

    // ---- key stream producer: thread 0 issues the TMA bulk copies (SASS UBLKCP) in consumption order, up to NSLOT
    // slots ahead of its own position.  pump(need) returns with slots [0, need) issued (blocking on the ring's `empty`
    // barriers if it must) and opportunistically issues further slots whose ring entry is already free.
    const bool producer = tid == 0;
    const uint8_t *ksrc = reinterpret_cast<const uint8_t *>(a.bsk_fft) + (single ? (size_t)__ldg(a.ggsw_index + ct0) * K::GGSW_BYTES : 0);
    uint32_t issued = 0;
    auto pump = [&](uint32_t need) {
        while (issued < total_slots && issued < it + (uint32_t)K::NSLOT) {
            const uint32_t s = issued % K::NSLOT;
            if (issued >= (uint32_t)K::NSLOT) {
                const uint32_t par = ((issued / K::NSLOT) - 1u) & 1u;
                if (issued < need) mbar_wait(empty + s, par, a.err_flag);
                else if (!mbar_try(empty + s, par)) break;
            }
            mbar_expect_tx(full + s, K::SLOT_BYTES);
            bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + (size_t)issued * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
            bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + (size_t)issued * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
            issued++;
        }
    };
    // lane 0 of a warp that is done with ring row r
    auto release_slot = [&](uint32_t r) {
        const uint32_t s = r % K::NSLOT;
        mbar_arrive(empty + s);
        if constexpr (SELF_REFILL) {
            const uint32_t k = r + (uint32_t)K::NSLOT, u = r / K::NSLOT;
            if (k < total_slots && mbar_test(empty + s, u & 1u)) {          // this arrival (or a later one) completed the slot's use u
                if (atomicCAS(claimed + s, u, u + 1u) == u) {               // one issuer per use
                    mbar_expect_tx(full + s, K::SLOT_BYTES);
                    bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + (size_t)k * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
                    bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + (size_t)k * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
                }
            }
        }
    };
    if constexpr (SELF_REFILL) { if (producer) pump(0); }   // the first NSLOT rows; every later row is issued by release_slot
    auto diff = [&](uint32_t pp, uint32_t j, uint32_t rot) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - sbase[pp * K::N + j]; };
    // one key row with the TMEM relay: the relaying team (it mod active) reads the shared-memory slot and parks the row in
    // TMEM slot it mod NT; the other teams wait for it there.  Every warp walks the same sequence `it`.
    const uint32_t quad = (tid >> 5) & 3u, tq = tmem_base + ((quad * 32u) << 16);
    uint32_t rteam = 0;   // = it mod active
    auto key_row = [&](auto domac_c, uint32_t p, uint32_t half) {
        constexpr bool DOMAC = decltype(domac_c)::value;
        const uint32_t s = it % K::NSLOT, ts = it % (uint32_t)RL::NT, tuse = it / (uint32_t)RL::NT;
        const uint32_t ta = tq + ts * (uint32_t)RL::SLOT_COLS;
        uint64_t *tf = tfull + ts * 4 + quad, *te = tempty + ts * 4 + quad;
        const cplx *peer = reinterpret_cast<const cplx *>(tm + K::TM_SUB + p * K::SUB_BYTES + K::STASH_BYTES);
        if (producer) pump(it + 1);
        if (active == 1u || team == rteam) {
            mbar_wait(full + s, (it / K::NSLOT) & 1u, a.err_flag);
            const cplx *slot = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES);
            if (active == 1u) {
                if constexpr (DOMAC) {
                    if (p == sub) mac_row_relay<K, true, false, true>(R, t, sub, slot, nullptr, half, 0u);
                    else mac_row_relay<K, false, false, true>(R, t, sub, slot, peer, half, 0u);
                }
            } else {
                if (tuse > 0u) mbar_wait(te, (tuse - 1u) & 1u, a.err_flag);   // the slot's previous row has been read
                tc_fence_after();
                if (p == sub) mac_row_relay<K, true, true, DOMAC>(R, t, sub, slot, nullptr, half, ta);
                else mac_row_relay<K, false, true, DOMAC>(R, t, sub, slot, peer, half, ta);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tf);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        } else {
            mbar_wait(tf, tuse & 1u, a.err_flag);
            tc_fence_after();
            if constexpr (DOMAC) {
                if (p == sub) mac_row_tmem<K, true, RL::CH>(R, t, nullptr, half, ta);
                else mac_row_tmem<K, false, RL::CH>(R, t, peer, half, ta);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(te);
        }
        it++;
        rteam = rteam + 1u == active ? 0u : rteam + 1u;
    };
#pragma unroll 1
    for (uint32_t i = 0; i < n_steps; i++) {
        const uint32_t rot = single ? 0u : (BMMP ? at[2 * i] : at[i]);
        const uint32_t rot1 = BMMP ? at[2 * i + 1] : 0u;
        // monomial exponents of the BMMP bundle: X^(a+a') - 1, X^a - 1, X^a' - 1
        const uint32_t ex0 = (rot + rot1) & (2u * K::N - 1u);
        cplx zb[3] = {};   // per-thread part of the three monomial factors of this step
        if constexpr (BMMP) {
            zb[0] = bmmp_base<K>(a.tw.ztab, t, ex0);
            zb[1] = bmmp_base<K>(a.tw.ztab, t, rot);
            zb[2] = bmmp_base<K>(a.tw.ztab, t, rot1);
        }
        if (!single && rot == 0 && rot1 == 0) {
            // diff == 0 (bundle == 0) => external product == 0 exactly: consume this step's slots without using them
#pragma unroll 1
            for (uint32_t s = 0; s < STEP_SLOTS; s++) {
                if constexpr (RL::ON) {
                    key_row(std::false_type{}, 0u, 0u);   // a relaying team still parks its rows for the others
                } else {
                    if constexpr (!SELF_REFILL) { if (producer) pump(it + 1); }
                    mbar_wait(full + (it % K::NSLOT), (it / K::NSLOT) & 1u, a.err_flag);
                    __syncwarp();
                    if (lane == 0) release_slot(it);
                    it++;
                }
            }
            continue;
        }
        zero_acc<K>(R);
        // OWN_FIRST (the ring holds all P rows of a level): a sub-team multiplies its OWN transformed row (still in its
        // registers) before the team barrier that publishes the rows, and runs the register part of the next level's F1
        // before the barrier that lets it overwrite its published row -- skew between sub-teams is absorbed by useful
        // work instead of barrier waits.  Rows are consumed out of ring order, so the producer issues a whole level up
        // front (it would otherwise wait at the team barrier for a sub-team that waits for a row it has not issued).
        if constexpr (OWN_FIRST) {
            auto mac_slot = [&](uint32_t p) {
                const uint32_t ir = it + p, s = ir % K::NSLOT;
                mbar_wait(full + s, (ir / K::NSLOT) & 1u, a.err_flag);
                const cplx *slot = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES);
                const cplx *peer = reinterpret_cast<const cplx *>(tm + K::TM_SUB + p * K::SUB_BYTES + K::STASH_BYTES);
                if (p == sub) phase_mac<K, true>(R, t, sub, slot, nullptr, 0u);
                else phase_mac<K, false>(R, t, sub, slot, peer, 0u);
                __syncwarp();
                if (lane == 0) release_slot(ir);
            };
#ifndef TFHE_FFT_P2P
#define TFHE_FFT_P2P 0
#endif
            // P2P: the two team-wide barriers of a level become point-to-point mbarriers -- pub[p] "sub-team p has published
            // its row of this level", rd[p] "every other sub-team has read it" -- so a sub-team waits only for the event it
            // depends on, and as late as possible (after the register part of the next pass).
            constexpr bool P2P = TFHE_FFT_P2P != 0;
            auto wait_row_read = [&]() {   // my previously published row (buf0) may be overwritten
                if constexpr (P2P) { if (lv > 0u) mbar_wait(rd + sub, (lv - 1u) & 1u, a.err_flag); }
                else team_bar_id(team_bar, K::TEAM_THREADS);
            };
#pragma unroll 1
            for (uint32_t lev = 0; lev < (uint32_t)K::L; lev++) {
                if constexpr (ACC_REG) phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - accv[k]; });
                else phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return diff(pp, j, rot); });
                if (P2P || lev > 0) wait_row_read();
                store_A<C>(R.x, buf0, t);
                if constexpr (!SELF_REFILL) { if (producer) pump(it + (uint32_t)K::P); }
                sub_sync();
                if constexpr (TW_REG) {
                    phase_F2v<K>(R, jbB, twB_base, buf0, buf1);
                    sub_sync();
                    phase_F3v<K>(R, t, twC_base, buf1);
                } else {
                    phase_F2<K>(R, jbB, twB, buf0, buf1);
                    sub_sync();
                    phase_F3<K>(R, t, twC, buf1);
                }
                phase_xstore<K>(R, t, buf0);
                if constexpr (P2P) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pub + sub);
                }
                mac_slot(sub);
                if constexpr (!P2P) team_bar_id(team_bar, K::TEAM_THREADS);   // all P transformed rows of this level are published
#pragma unroll 1
                for (uint32_t p = 0; p < (uint32_t)K::P; p++) {
                    if (p == sub) continue;
                    if constexpr (P2P) mbar_wait(pub + p, lv & 1u, a.err_flag);
                    mac_slot(p);
                    if constexpr (P2P) {   // mac_slot ends with __syncwarp: every lane is past its reads of row p
                        if (lane == 0) mbar_arrive(rd + p);
                    }
                }
                it += (uint32_t)K::P;
                lv++;
            }
            if constexpr (!P2P) team_bar_id(team_bar, K::TEAM_THREADS);   // the last published rows have been read: buf0 may be overwritten
        } else
#pragma unroll 1
        for (uint32_t lev = 0; lev < (uint32_t)K::L; lev++) {
            // forward transform of this sub-team's digit row (polynomial `sub`, level `lev`)
#if (TFHE_FFT_ABLATE & 1)
            phase_F1a<K>(R, t, sub, 1u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return 0u; });
#else
            if constexpr (BMMP && ACC_REG) phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return accv[k]; });
            else if constexpr (BMMP) phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return acc[pp * K::N + j]; });
            else if constexpr (ACC_REG) phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - accv[k]; });
            else phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return diff(pp, j, rot); });
#endif
            // rows in ring order; P2PG: point-to-point row barriers instead of the two team-wide barriers of a level (a
            // sub-team waits for the publisher of the row it is about to read, and for its readers only before it overwrites)
            if constexpr (P2PG) { if (lv > 0u) mbar_wait(rd + sub, (lv - 1u) & 1u, a.err_flag); }
            store_A<C>(R.x, buf0, t);
            sub_sync();
            if constexpr (K::SINGLE_BUF) {           // one buffer: a barrier between every load and the next store
                phase_F2a<K>(R, jbB, twB, buf0);
                sub_sync();
                phase_F2b<K>(R, jbB, buf0);
                sub_sync();
                phase_F3<K>(R, t, twC, buf0);
                sub_sync();
            } else {
                phase_F2<K>(R, jbB, twB, buf0, buf1);
                sub_sync();
                phase_F3<K>(R, t, twC, buf1);
            }
            phase_xstore<K>(R, t, buf0);             // buf0 is free: every thread of the sub-team is past its loads from it
            if constexpr (P2PG) {
                __syncwarp();
                if (lane == 0) mbar_arrive(pub + sub);
            } else {
                team_bar_id(team_bar, K::TEAM_THREADS);  // all P transformed rows of this level are published
            }
#pragma unroll 1
            for (uint32_t pk = 0; pk < (uint32_t)K::P * KEYS * K::HALVES; pk++) {
                const uint32_t p = pk / (KEYS * K::HALVES), which = (pk / K::HALVES) % KEYS, half = pk % K::HALVES;
                if constexpr (RL::ON) {
                    key_row(std::true_type{}, p, half);
                    continue;
                }
                const uint32_t s = it % K::NSLOT;
                if (producer) pump(it + 1);
                if constexpr (P2PG) { if (p != sub && pk % (KEYS * K::HALVES) == 0u) mbar_wait(pub + p, lv & 1u, a.err_flag); }
                mbar_wait(full + s, (it / K::NSLOT) & 1u, a.err_flag);
#if !(TFHE_FFT_ABLATE & 2)
                const cplx *slot = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES);
                const cplx *peer = reinterpret_cast<const cplx *>(tm + K::TM_SUB + p * K::SUB_BYTES + K::STASH_BYTES);
                if constexpr (BMMP) {
                    if (p == sub) phase_mac_bmmp<K, true>(R, t, sub, slot, nullptr, a.tw.ztab, which == 0 ? ex0 : which == 1 ? rot : rot1, which == 0 ? zb[0] : which == 1 ? zb[1] : zb[2]);
                    else phase_mac_bmmp<K, false>(R, t, sub, slot, peer, a.tw.ztab, which == 0 ? ex0 : which == 1 ? rot : rot1, which == 0 ? zb[0] : which == 1 ? zb[1] : zb[2]);
                } else {
                    (void)which;
                    if (p == sub) phase_mac<K, true>(R, t, sub, slot, nullptr, half);
                    else phase_mac<K, false>(R, t, sub, slot, peer, half);
                }
#endif
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(empty + s);
                    if constexpr (P2PG) { if (p != sub && pk % (KEYS * K::HALVES) == KEYS * K::HALVES - 1u) mbar_arrive(rd + p); }
                }
                it++;
            }
            if constexpr (P2PG) lv++;
            else team_bar_id(team_bar, K::TEAM_THREADS);  // the published rows have been read: buf0 may be overwritten
        }
        if constexpr ((OWN_FIRST && (TFHE_FFT_P2P != 0)) || P2PG) {
            if (lv > 0u) mbar_wait(rd + sub, (lv - 1u) & 1u, a.err_flag);   // the last published row has been read
        }
#if !(TFHE_FFT_ABLATE & 16)
        // inverse transforms of this sub-team's column: the low- and high-limb products together (fft_team.cuh phase_J*)
        if constexpr (K::SINGLE_BUF) {   // one limb after the other through the single buffer
            uint32_t lo[2 * K::E];
            phase_K1<K, 0>(R, t, twC, buf0);
            sub_sync();
            phase_K2a<K, 0>(R, jbB, twB, buf0);
            if (producer) pump(0);
            sub_sync();
            phase_K2b<K, 0>(R, jbB, buf0);
            sub_sync();
            phase_K3_lo<K>(R, t, a.tw.twA, buf0, lo, maxfrac);
            sub_sync();
            phase_K1<K, 1>(R, t, twC, buf0);
            sub_sync();
            phase_K2a<K, 1>(R, jbB, twB, buf0);
            sub_sync();
            phase_K2b<K, 1>(R, jbB, buf0);
            sub_sync();
            phase_K3_hi<K>(R, t, a.tw.twA, buf0, lo, acc + sub * K::N, maxfrac);
        } else {
            if constexpr (TW_REG) {
                phase_J1v<K>(R, t, twC_base, buf0, buf1);
                sub_sync();
                phase_J2av<K>(R, jbB, twB_base, buf0, buf1);
            } else {
                phase_J1<K>(R, t, twC, buf0, buf1);
                sub_sync();
                phase_J2a<K>(R, jbB, twB, buf0, buf1);
            }
            if constexpr (!SELF_REFILL) { if (producer) pump(0); }   // ring entries freed by slower teams: refill them while this team inverts
            sub_sync();
            phase_J2b<K>(R, jbB, buf0, buf1);
            sub_sync();
            if constexpr (ACC_REG) {
                phase_J3r<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, accv, maxfrac);
            } else {
                phase_J3<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, maxfrac);
            }
        }
#endif
        sub_sync();   // acc[sub] (read only by this sub-team) is up to date before the next step's digits
    }
    uint32_t *out = a.glwe_out + ((size_t)ct * K::P + sub) * K::N;
    for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[sub * K::N + idx];
    if constexpr (RL::ON) {
        static_assert(!RL::ON || K::CTS * (K::T == 32 ? 1 : K::P + 1) < 15, "named barrier 15 is the TMEM release barrier");
        tc_fence_before();
        team_bar_id(15, (int)(active * K::TEAM_THREADS));   // every active team is done with TMEM
        if (tid < 32) {
            tc_fence_after();
            tmem_dealloc(tmem_base, RL::COLS);
        }
    }
    if constexpr (K::CHECK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxfrac = fmax(maxfrac, __shfl_xor_sync(0xFFFFFFFFu, maxfrac, o));
        if (lane == 0) atomicMax(a.margin, (unsigned long long)__double_as_longlong(maxfrac));
    }
}

// ------------------------------------------------------------------------------------------ key transform
// grid = number of polynomials (n*ROWS*P); block = 2T (one team per limb); in natural [n][ROWS][P][N] u32,
// out [n][ROWS (level-major)][2][P][M] complex.
struct FftTransformArgs {
    TwTablesF tw;
    const uint32_t *raw;
    cplx *out;
    uint32_t keys_per_step;   // 1: standard key; 3: BMMP key triples, stored [step][row (level-major)][which][limb][P][M]
};
template <class K>
__global__ void __launch_bounds__(2 * K::T) bsk_fft_transform_kernel(const __grid_constant__ FftTransformArgs a) {
    using C = typename K::F;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, limb = tid / K::T, t = tid % K::T;
    cplx *buf0 = reinterpret_cast<cplx *>(smem) + limb * 2 * C::MPAD, *buf1 = buf0 + C::MPAD;
    const size_t poly = blockIdx.x;  // = (i*ROWS + r)*P + c, r = p*L + lev (ggsw.rs:37-41)
    const size_t ir = poly / K::P, c = poly % K::P, i = ir / K::ROWS, r = ir % K::ROWS;
    const uint32_t *g = a.raw + poly * K::N;
    const size_t kps = a.keys_per_step, step = i / kps, which = i % kps;
    const size_t row = (step * K::ROWS + key_row_index<K>((uint32_t)(r / K::L), (uint32_t)(r % K::L))) * kps + which;   // consumption order
    cplx *o = a.out + ((row * K::HALVES * 2 + limb) * K::P + c) * K::MH;   // half 0; phase_T3 adds the half stride
    FftRegs<K> R;
    phase_T1<K>(R, t, (int)limb, g, a.tw.twA, buf0);
    team_bar_id(limb + 1, K::T);
    phase_F2<K>(R, jbase_B<C>(t), a.tw.twB + (t >> C::QB) * C::NB_TW, buf0, buf1);
    team_bar_id(limb + 1, K::T);
    phase_T3<K>(R, t, a.tw.twC, buf1, o);
}

// ------------------------------------------------------------------------------------------ FP64 pipe peaks
// Dependent-free DFMA loops: KIND 0 = two register operands + a constant (the pipe's issue rate, 64 lanes/clk/SM);
// KIND 1 = three distinct register operands, as in the multiply-accumulate and the pass-B/C butterflies (measured
// 2/3 of KIND 0 on B200: the third 64-bit register operand costs a cycle); KIND 2 = the forward butterfly stream
// (4 three-register + 2 two-register DFMA per butterfly) from registers.
template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *sink, double a, double b, int iters) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = a + threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (threadIdx.x * 3 + i); }
    for (int it = 0; it < iters; it++) {
        if (KIND == 2) {
            cplx *c = reinterpret_cast<cplx *>(x);   // 4 complex points, 2 stages
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ct_bfly(c[0], c[2], cplx{y[u], z[u]});
                ct_bfly(c[1], c[3], cplx{y[u + 1], z[u + 1]});
                ct_bfly(c[0], c[1], cplx{y[u + 2], z[u + 2]});
                ct_bfly(c[2], c[3], cplx{y[u + 3], z[u + 3]});
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = KIND == 0 ? __fma_rn(x[i], y[(i + u) & 7], b) : __fma_rn(x[i], y[(i + u) & 7], z[(i + 2 * u + 1) & 7]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 1.2345) sink[0] = s;
}

}  // namespace fft
}  // namespace tfhe
