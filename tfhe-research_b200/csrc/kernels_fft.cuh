// kernels_fft.cuh -- sm_100a kernels of the FP64-FFT arithmetic path (see fft_team.cuh for the arithmetic).
//
//   pbs_fft_kernel        blind rotation: ONE CTA = CTS ciphertexts; each ciphertext is a team of P = k+1 sub-teams
//                         of T threads (sub-team s: digits + forward transforms of polynomial s, multiply-accumulate +
//                         inverse transforms of result column s); the GLWE accumulator of every team is resident in
//                         shared memory for all n CMUX steps; the FFT-domain bootstrapping key streams global -> shared
//                         with TMA bulk copies into a ring (full/empty mbarriers), issued by thread 0 or -- in the
//                         own-row-first level loop -- by whichever warp frees a ring slot.  Every key byte fetched from
//                         L2 is used by all CTS ciphertexts of the CTA.  Also runs one external product / CMUX for the
//                         sub-operation entry points.
//   bsk_fft_transform_kernel   raw BSK -> limb-split FFT domain (one-off at key upload)
//
// Reference: bootstrapping.rs:58-105 (blind rotation), ggsw.rs:132-178 (external product, cmux).
#pragma once
#include "fft_team.cuh"
#include "fft_tmem.cuh"
#include "kernels.cuh"

namespace tfhe {
namespace fft {

struct FftArgs {
    TwTablesF tw;
    const cplx *bsk_fft;       // [n][L][P slots][2 limbs][P][M], diagonal-major (fft_team.cuh key_slot_index), slot order, pre-scaled by 1/M
    const uint32_t *lwe_in;    // mode 0: [B][n+1]
    const uint32_t *luts;      // [T][N] unencoded
    const uint32_t *lut_idx;   // [B] or nullptr
    const uint32_t *in0, *in1; // mode 1 (out = ExtProd(G, in0)) / mode 2 (out = ExtProd(G, in1-in0)+in0): [B][P][N]
    const uint32_t *ggsw_index;
    uint32_t *glwe_out;        // [B][P][N]
    uint32_t *err_flag;        // bit 0: test-vector entry >= 2^log_p (glwe.rs:144); bit 1: TMA wait timed out; bit 2: lut_idx out of range
    unsigned long long *margin;  // largest |x - rint(x)| before rounding (bits of a non-negative double), CHECK only
    unsigned long long *prof;    // optional phase-cycle counters of the latency kernel (measurement only; nullptr in production)
    uint32_t n, batch, mode, log_p, enc_shift;
    uint32_t n_luts;           // lut_idx[b] >= n_luts sets err_flag bit 2 and selects test vector 0
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

// BMMP = true: blind rotation unrolled by two (notes/BMMP Bootstrapping.md): n/2 steps, each consumes the three GGSWs of
// a key triple (3 slots per row), decomposes acc itself and adds ExtProd(bundle, acc); blind rotation mode only.
//
// Key stream.  The key is stored diagonal-major (fft_team.cuh key_slot_index): slot d of a level holds, at column position c,
// the polynomial (row of polynomial (c + d) mod P, column c).  Sub-team s always reads column position s and multiplies, at
// slot d, the transformed digits of polynomial (s + d) mod P: its OWN row (still in its registers) at d = 0 -- for every s --
// and the published rows of the other sub-teams after that.  So the slots stream in one fixed order through a two-slot ring,
// and every sub-team starts a level without waiting for anyone.
// Two level loops (what they compute is the same):
//  * OWN_FIRST (P0, P1) -- the own row is multiplied before the team barrier that publishes the rows, and the register part
//    of the next level's F1 runs before the barrier that lets a sub-team overwrite its published row: skew between sub-teams
//    is absorbed by useful work instead of barrier waits (P1: 74.8 -> 70.8 ms per batch of 4096).  There is no producer
//    thread: the warp whose arrival frees a ring slot issues the TMA copy of the row that reuses it (release_slot), i.e. a
//    refill starts the moment the slot's last reader is done (P1: -> 69.4 ms; with thread 0 as producer 68.5 vs 65.2 ms in v11).
//  * general (P2: one exchange buffer, half-row slots; the BMMP variant: three keys per row) -- same order, own row before
//    the publish barrier; thread 0 is the producer (BMMP) or the last reader of a slot refills it (P2: 196.6 -> 183.5 ms).
// Variants measured and dropped are listed in profiles/r01_fft_v8_variants.README.
#ifndef TFHE_FFT_OWNFIRST
#define TFHE_FFT_OWNFIRST 1
#endif
#ifndef TFHE_FFT_ACCREG
#define TFHE_FFT_ACCREG 1
#endif
#ifndef TFHE_FFT_PRETEST
#define TFHE_FFT_PRETEST 1
#endif
template <class K, bool BMMP = false>
__global__ void __launch_bounds__(K::THREADS, 1) pbs_fft_kernel(const __grid_constant__ FftArgs a) {
    using C = typename K::F;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    // tensor-memory exchanges: a warp reaches only the 32 lanes of its quarter (warp % 4), and the sub-teams of a ciphertext publish
    // their rows to each other through those lanes -- so team = lane quarter, sub-team = warp / 4
    const uint32_t team = K::XCHG ? (tid >> 5) & 3u : tid / K::TEAM_THREADS, sub = K::XCHG ? tid >> 7 : (tid % K::TEAM_THREADS) / K::T;
    const uint32_t t = K::XCHG ? lane : tid % K::T, tt = sub * K::T + t;
    const uint32_t team_bytes = (uint32_t)K::team_bytes((int)a.n);
    uint8_t *ring = smem + K::CTS * team_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES), *empty = full + K::NSLOT;
    uint32_t *claimed = reinterpret_cast<uint32_t *>(empty + K::NSLOT);   // [NSLOT] refills issued per ring slot (SELF_REFILL)
    constexpr bool OWN_FIRST = (TFHE_FFT_OWNFIRST != 0) && !BMMP && !K::SINGLE_BUF && K::HALVES == 1 && !(TFHE_FFT_ABLATE);
    // ring slots are refilled by their last reader (release_slot) instead of by thread 0: measured per configuration -- P1 68.5 ->
    // 65.2 ms, N = 2048 (half-row slots) 196.6 -> 183.5 ms, P0 neutral (116.7 / 117.2 ms); the BMMP variant was slower with it
    // (70.7 -> 72.6 ms) and keeps the producer thread.
#ifndef TFHE_FFT_SELFREFILL
#define TFHE_FFT_SELFREFILL 1
#endif
    constexpr bool SELF_REFILL = (TFHE_FFT_SELFREFILL != 0) && (OWN_FIRST || (K::HALVES > 1 && !BMMP));

    const bool single = K::HAS_SINGLE_MODES && a.mode != 0;   // sub-operation entry points: one ciphertext (team 0) per CTA, its own GGSW
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
            mbar_init(empty + s, active * K::P * K::WARPS_PER_SUB);   // every warp of every active team consumes every slot
            claimed[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tensor-memory exchanges (fft_tmem.cuh): warp 0 allocates all 512 columns (one CTA per SM) -- every warp owns 96 columns
    // (warp / 4) of its 32 lanes (warp % 4): 32 for the exchanges of its transforms, 2 x 32 for the rows it publishes, and 48 more
    // behind those of all sub-teams for its derived twiddles
    uint32_t *tmem_ptr = claimed + K::NSLOT;
    // TWT (P1): 256 columns, 64 per warp (warp / 4) of its lanes, for the derived twiddles of passes B and C
    constexpr bool TWT = K::TWT && OWN_FIRST;
    // TAIL (P1): per warp 64 columns for the tensor-memory swaps of the last three stages of every transform (2 x 32 for the paired
    // inverse) + 32 for the parked twiddles of those stages + 32 for pass B's: 3 warps per lane quarter = 384 of the 512 columns
    constexpr bool TAIL = K::TAIL && !K::TAIL16 && OWN_FIRST;
    // TAIL16 (N = 2048): ring-order loop, single exchange buffer; per warp 64 swap columns + 32 + 64 of parked twiddles, 2 warps per quarter
    constexpr bool TAIL16 = K::TAIL16 && !OWN_FIRST && !BMMP;
#ifndef TFHE_TMEM_TAIL16_LATEBAR
#define TFHE_TMEM_TAIL16_LATEBAR 1
#endif
#ifndef TFHE_TMEM_TAIL16_TWB
#define TFHE_TMEM_TAIL16_TWB 1   // the 15 derived twiddles of pass B'' wait in tensor memory too (two 32-column blocks per warp)
#endif
    static_assert(!TAIL16 || ((K::THREADS / 32 + 3) / 4) * (TMEM_TAIL16_COLS + TMEM_TAIL16_TW_COLS + 64u) <= 512, "tensor-memory columns");
    static_assert(!TAIL || ((K::THREADS / 32 + 3) / 4) * (TMEM_TAIL_COLS + TMEM_TAIL_TW_COLS) <= 512, "tensor-memory columns");
    constexpr bool USE_TMEM = K::XCHG || TWT || TAIL || TAIL16;
    constexpr uint32_t TMEM_ALLOC = (K::XCHG || TAIL || TAIL16) ? 512u : 256u;
    static_assert(!TWT || ((K::THREADS / 32 + 3) / 4) * TMEM_TWT_COLS <= TMEM_ALLOC, "tensor-memory columns");
    static_assert(!TWT || (K::F::NB_TW <= 8 && K::F::NC_TW <= 8), "one 32-column block per pass");
    if constexpr (USE_TMEM) {
        static_assert(!K::XCHG || K::P * TMEM_SUB_COLS <= 512, "tensor-memory columns");
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_ALLOC) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();   // the only CTA-wide barrier: teams are independent below
    uint32_t taddr = 0, tquarter = 0, pubsel = 0;   // own columns / first column of the team's lanes / publish buffer of the next level
    TmemTw twx = {};
    uint32_t twt_cols = 0;
    if constexpr (TWT || TAIL) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        twt_cols = *tmem_ptr + ((((tid >> 5) & 3u) * 32u) << 16) + (tid >> 7) * TMEM_TWT_COLS;
        static_assert(TMEM_TWT_COLS == TMEM_TAIL_COLS, "same column budget per warp");
    }
    TailTw tailtw = {};
    if constexpr (TAIL) {
        tailtw = load_tail_tw(a.tw.twX, t);
#if TFHE_TMEM_TAIL_TWSTORE
        if (team < active)
            tmem_tail_tw_setup(tailtw, *tmem_ptr + ((((tid >> 5) & 3u) * 32u) << 16) + ((K::THREADS / 32 + 3) / 4) * TMEM_TAIL_COLS + (tid >> 7) * TMEM_TAIL_TW_COLS);
#endif
    }
    if constexpr (K::XCHG) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tquarter = *tmem_ptr + ((team * 32u) << 16);
        taddr = tquarter + sub * TMEM_SUB_COLS;
        twx = load_tmem_tw(a.tw.twX, lane);
#if TFHE_TMEM_TWSTORE
        static_assert(K::P * TMEM_SUB_COLS + K::P * TMEM_TW_COLS <= 512, "tensor-memory columns");
        if (team < active) tmem_tw_setup(twx, tquarter + (uint32_t)K::P * TMEM_SUB_COLS + sub * TMEM_TW_COLS);
#endif
    }

    if (team >= active) {
        if constexpr (USE_TMEM) {   // warp 0 releases tensor memory once every warp has arrived here or at the end of the kernel
            if (active == 0) {
                if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr), "n"(TMEM_ALLOC) : "memory");
            } else {
                asm volatile("bar.arrive 15, %0;" ::"n"(K::THREADS) : "memory");
            }
        }
        return;
    }
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
    // A thread's two twiddle-table entries (passes B and C) never change.  The OWN_FIRST loop keeps them in registers
    // (P1: 71.4 -> 69.8 ms); with the ring-order loop the 8 registers cost more than the loads after every barrier
    // (P1: 74.8 -> 79.9 ms, P0 and the BMMP variant alike), so that loop fetches them per pass.
    cplx twB_base = {}, twC_base = {};
    uint32_t jb_swz = 0;
    Tail16Tw t16 = {};
    uint32_t t16_swap = 0, t16_twb = 0;
    cplx *b16_even = nullptr, *b16_odd = nullptr;   // this thread's even / odd registers in layout B''16 (fft_team.cuh store_Bsw16)
    if constexpr (TAIL16) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t jb = jbase_Bsw16(t), x3 = (jb >> 6) & 7u, low = jb ^ (x3 & 3u), j8 = (x3 >> 2) & 1u;
        b16_even = buf0 + low + (j8 ? 4u : 0u);
        b16_odd = buf0 + low + (j8 ? 0u : 4u);
        const uint32_t q0 = *tmem_ptr + ((((tid >> 5) & 3u) * 32u) << 16);
        t16_swap = q0 + (tid >> 7) * TMEM_TAIL16_COLS;
        t16 = load_tail16_tw(a.tw.twX, t, jb >> 6);
        tmem_tail16_tw_setup(t16, q0 + ((K::THREADS / 32 + 3) / 4) * TMEM_TAIL16_COLS + (tid >> 7) * TMEM_TAIL16_TW_COLS);
#if TFHE_TMEM_TAIL16_TWB
        {
            t16_twb = q0 + ((K::THREADS / 32 + 3) / 4) * (TMEM_TAIL16_COLS + TMEM_TAIL16_TW_COLS) + (tid >> 7) * 64u;
            cplx tw[16];
            derive_pass_tw<4>(tw, t16.pB);
            tmem_tw_block_store<7>(t16_twb, tw);            // stages 4, 5, 6
            tmem_tw_block_store<8>(t16_twb + 32u, tw + 7);  // stage 7
            tmem_wait_st();
        }
#endif
    }
#ifndef TFHE_TMEM_TAIL_TWB
#define TFHE_TMEM_TAIL_TWB 1
#endif
    uint32_t tail_twb_cols = 0;
    if constexpr (TAIL) {   // pass B of layout B'': the table row of this thread's (j8 j7 j6); no pass C
        twB_base = pass_tw_base<C::QB>(a.tw.twB + hA_Bsw(t) * C::NB_TW, 1);
        jb_swz = swz9(jbase_Bsw(t));
#if TFHE_TMEM_TAIL_TWB
        {   // its derived twiddles wait in tensor memory as well (one more 32-column block per warp)
            static_assert(((K::THREADS / 32 + 3) / 4) * (TMEM_TAIL_COLS + TMEM_TAIL_TW_COLS + 32u) <= 512, "tensor-memory columns");
            tail_twb_cols = *tmem_ptr + ((((tid >> 5) & 3u) * 32u) << 16) + ((K::THREADS / 32 + 3) / 4) * (TMEM_TAIL_COLS + TMEM_TAIL_TW_COLS) + (tid >> 7) * 32u;
            cplx tw[8];
            derive_pass_tw<C::QB>(tw, twB_base);
            tmem_tw_block_store<C::NB_TW>(tail_twb_cols, tw);
            tmem_wait_st();
        }
#endif
    } else if constexpr (OWN_FIRST) {
        twB_base = pass_tw_base<C::QB>(twB, 1);
        twC_base = pass_tw_base<C::LOGE>(twC + t, C::T);
        if constexpr (TWT) {   // derive once, keep in tensor memory: 14 FP64 operations per pass and 8 registers less
            cplx tw[8];
            derive_pass_tw<C::QB>(tw, twB_base);
            tmem_tw_block_store<C::NB_TW>(twt_cols, tw);
            derive_pass_tw<C::LOGE>(tw, twC_base);
            tmem_tw_block_store<C::NC_TW>(twt_cols + 32u, tw);
            tmem_wait_st();
        }
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
        uint32_t li = a.lut_idx ? __ldg(a.lut_idx + ct) : 0u;
        if (li >= a.n_luts) {   // never read a test vector out of bounds: flag it (host returns TFHE_E_PARAM) and fall back to 0
            atomicOr(a.err_flag, 4u);
            li = 0u;
        }
        const uint32_t *lut = a.luts + (size_t)li * K::N;
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
    uint32_t it = 0;   // position in the key stream (the same sequence in every warp)
    uint32_t pre_row = 0xFFFFFFFFu;   // ring row whose `full` barrier was tested ahead of time, and the outcome
    bool pre_ok = false;
    // the 2E words of acc[sub] this thread decomposes are the ones it updates: they cross the step boundary in registers
    // (not with the single exchange buffer of N = 2048, which is out of registers)
#ifndef TFHE_FFT_ACCREG_G
#define TFHE_FFT_ACCREG_G 1
#endif
    constexpr bool ACC_REG = (TFHE_FFT_ACCREG != 0) && (OWN_FIRST || ((TFHE_FFT_ACCREG_G != 0) && !K::SINGLE_BUF && !(TFHE_FFT_ABLATE)));
    uint32_t accv[2 * K::E];
    if constexpr (ACC_REG) {
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            const uint32_t j = ((uint32_t)e << C::LOGT) | t;
            accv[2 * e] = sbase[sub * K::N + j];
            accv[2 * e + 1] = sbase[sub * K::N + j + K::M];
        }
    }

    // ---- key stream.  Ring-order loop: thread 0 issues the TMA bulk copies (SASS UBLKCP) in consumption order, up to
    // NSLOT slots ahead of its own position; pump(need) returns with slots [0, need) issued (blocking on the ring's
    // `empty` barriers if it must) and opportunistically issues further slots whose ring entry is already free.
    // SELF_REFILL: thread 0 only issues the first NSLOT rows; every later row is issued by release_slot.
    const bool producer = tid == 0;
    const uint8_t *ksrc = reinterpret_cast<const uint8_t *>(a.bsk_fft) + (single ? (size_t)__ldg(a.ggsw_index + ct0) * K::GGSW_BYTES : 0);
    auto issue_row = [&](uint32_t row) {
        const uint32_t s = row % K::NSLOT;
        mbar_expect_tx(full + s, K::SLOT_BYTES);
        bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + (size_t)row * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
        bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + (size_t)row * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
    };
    uint32_t issued = 0;
    auto pump = [&](uint32_t need, uint32_t pos) {   // pos = the caller's position in the key stream
        while (issued < total_slots && issued < pos + (uint32_t)K::NSLOT) {
            const uint32_t s = issued % K::NSLOT;
            if (issued >= (uint32_t)K::NSLOT) {
                const uint32_t par = ((issued / K::NSLOT) - 1u) & 1u;
                if (issued < need) mbar_wait(empty + s, par, a.err_flag);
                else if (!mbar_try(empty + s, par)) break;
            }
            issue_row(issued);
            issued++;
        }
    };
    // lane 0 of a warp that is done with ring row r.  SELF_REFILL: the slot's use u = r / NSLOT is complete once every warp
    // has arrived; whoever sees that first (at the latest the last arriver, right after its own arrival) claims the refill.
    // (looking at the outcome of the test one row later, to hide its ~100 cycles behind the next multiply-accumulate, was measured with
    // the 6-row ring of the tensor-memory configuration: 97.3 -> 105.0 ms -- the refill that waits starves the ring)
    // (testing the rows of a level together at the end of the level, one round trip per level: 95.0 -> 95.7 ms, not kept either)
    auto release_slot = [&](uint32_t r) {
        const uint32_t s = r % K::NSLOT;
        mbar_arrive(empty + s);
        if constexpr (SELF_REFILL) {
            const uint32_t k = r + (uint32_t)K::NSLOT, u = r / K::NSLOT;
            if (k < total_slots && mbar_test(empty + s, u & 1u)) {
                if (atomicCAS(claimed + s, u, u + 1u) == u) issue_row(k);   // one issuer per use
            }
        }
    };
    if constexpr (SELF_REFILL) { if (producer) pump(0, it); }
    auto diff = [&](uint32_t pp, uint32_t j, uint32_t rot) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - sbase[pp * K::N + j]; };
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
            for (uint32_t s = 0; s < STEP_SLOTS; s++, it++) {
                if constexpr (!SELF_REFILL) { if (producer) pump(it + 1, it); }
                mbar_wait(full + (it % K::NSLOT), (it / K::NSLOT) & 1u, a.err_flag);
                __syncwarp();
                if (lane == 0) release_slot(it);
            }
            continue;
        }
        zero_acc<K>(R);
        if constexpr (OWN_FIRST) {
            auto mac_slot = [&](auto own_c, uint32_t d) {   // ring row it + d: slot d of this level = row of polynomial (sub + d) mod P
                constexpr bool OWN = decltype(own_c)::value;
                const uint32_t ir = it + d, s = ir % K::NSLOT, p = (sub + d) % (uint32_t)K::P;
                if constexpr (!SELF_REFILL) { if (producer) pump(ir + 1, ir); }
                if constexpr (K::XCHG && !OWN && TFHE_TMEM_LOADFIRST) tmem_load_row(R.x, tquarter + p * TMEM_SUB_COLS + TMEM_PUB_COL + 32u * pubsel);
#if TFHE_FFT_PRETEST
                // the barrier of this row was tested while the previous row was being multiplied (the test has a latency of its own,
                // ~100 cycles even when the row has long arrived: 18 of them per step were 5 % of the P0 kernel's time)
                if (!(pre_row == ir && pre_ok)) mbar_wait(full + s, (ir / K::NSLOT) & 1u, a.err_flag);
                pre_row = ir + 1u;
                pre_ok = mbar_test(full + pre_row % K::NSLOT, (pre_row / K::NSLOT) & 1u);
#else
                mbar_wait(full + s, (ir / K::NSLOT) & 1u, a.err_flag);
#endif
                if constexpr (K::XCHG && !OWN && !TFHE_TMEM_LOADFIRST) tmem_load_row(R.x, tquarter + p * TMEM_SUB_COLS + TMEM_PUB_COL + 32u * pubsel);   // the peer's row: same lane, its columns
                const cplx *slot = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES);
                const cplx *peer = reinterpret_cast<const cplx *>(tm + K::TM_SUB + p * K::SUB_BYTES + K::STASH_BYTES);
                phase_mac<K, OWN || K::XCHG>(R, t, sub, slot, peer, 0u);
                __syncwarp();
                if (lane == 0) release_slot(ir);
            };
            auto level = [&](auto first_c, uint32_t lev) {
                constexpr bool FIRST = decltype(first_c)::value;
#if TFHE_FFT_ACCREG
                // level 0 is peeled off the loop: the subtrahend of the decomposition comes from accv (registers, handed over
                // by the previous step's J3), and accv is dead from here to the end of the step
                if constexpr (FIRST) phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - accv[k]; });
                else phase_F1a<K, 2>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t, uint32_t) { return 0u; });
#else
                phase_F1a<K>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return diff(pp, j, rot); });
#endif
                if constexpr (K::XCHG) {
                    tmem_fwd_rest(R.x, taddr, twx);          // the other passes, exchanged through tensor memory: layout F
                    // published in tensor memory, double buffered: the buffer written here was last read two levels ago, and
                    // every peer finished that level before it arrived at the barrier of the level between
                    tmem_store_row(R.x, taddr + TMEM_PUB_COL + 32u * pubsel);
                    mac_slot(std::true_type{}, 0u);
                    tmem_wait_st();                          // the published row has landed
                    tmem_fence_before_sync();
                    team_bar_id(team_bar, K::TEAM_THREADS);  // all P transformed rows of this level are published
                    tmem_fence_after_sync();
#pragma unroll 1   // (unrolled: 112 bytes of spills, 95.3 -> 109.3 ms)
                    for (uint32_t d = 1; d < (uint32_t)K::P; d++) mac_slot(std::false_type{}, d);
                    it += (uint32_t)K::P;
                    pubsel ^= 1u;
                    return;
                } else if constexpr (TAIL) {
                    // (exchange and publish through ONE buffer alternating by level, a sub-team barrier instead of the team barrier below, was
                    // measured slower: 61.3 -> 66.9 ms)
                    store_Asw<C>(R.x, buf1, t);              // the one shared-memory exchange goes through buf1: buf0 still holds the published row
#if TFHE_TMEM_TAIL_TWB
                    {
                        TwRaw32 raw;
                        cplx tw[8];
                        tmem_tw_block_request(raw, tail_twb_cols);   // R.x is dead here
                        sub_sync();
                        tmem_tw_block_claim<C::NB_TW>(raw, tw);
                        load_Bsw<C>(R.x, buf1, jb_swz);
                        fwd_pass<C::LOGE, C::QB>(R.x, tw);
                    }
#else
                    sub_sync();
                    {
                        cplx tw[C::NB_TW];
                        derive_pass_tw<C::QB>(tw, twB_base);
                        load_Bsw<C>(R.x, buf1, jb_swz);
                        fwd_pass<C::LOGE, C::QB>(R.x, tw);
                    }
#endif
                    tmem_fwd_tail9(R.x, twt_cols, tailtw);   // stages 6..8 inside the warp: layout F9
                    if (!FIRST) team_bar_id(team_bar, K::TEAM_THREADS);   // the row published at the previous level has been read
                } else {
                    if (!FIRST) team_bar_id(team_bar, K::TEAM_THREADS);   // the row published at the previous level has been read
                    store_A<C>(R.x, buf0, t);
                    if constexpr (TWT) {
                        TwRaw32 raw;
                        cplx tw[8];
                        tmem_tw_block_request(raw, twt_cols);         // R.x is dead here: the request costs no registers
                        sub_sync();
                        tmem_tw_block_claim<C::NB_TW>(raw, tw);
                        phase_F2w<K>(R, jbB, tw, buf0, buf1);
                        tmem_tw_block_request(raw, twt_cols + 32u);
                        sub_sync();
                        tmem_tw_block_claim<C::NC_TW>(raw, tw);
                        phase_F3w<K>(R, t, tw, buf1);
                    } else {
                        sub_sync();
                        phase_F2v<K>(R, jbB, twB_base, buf0, buf1);
                        sub_sync();
                        phase_F3v<K>(R, t, twC_base, buf1);
                    }
                }
                phase_xstore<K>(R, t, buf0);             // buf0 is free: every thread of the sub-team is past its loads from it
                mac_slot(std::true_type{}, 0u);
                team_bar_id(team_bar, K::TEAM_THREADS);  // all P transformed rows of this level are published
#pragma unroll 1
                for (uint32_t d = 1; d < (uint32_t)K::P; d++) mac_slot(std::false_type{}, d);
                it += (uint32_t)K::P;
            };
#if TFHE_FFT_ACCREG
            level(std::true_type{}, 0u);
#pragma unroll 1
            for (uint32_t lev = 1; lev < (uint32_t)K::L; lev++) level(std::false_type{}, lev);
#else
#pragma unroll 1
            for (uint32_t lev = 0; lev < (uint32_t)K::L; lev++) level(std::false_type{}, lev);
#endif
            if constexpr (!K::XCHG) team_bar_id(team_bar, K::TEAM_THREADS);      // the last published rows have been read: buf0 may be overwritten
        } else {
            auto level = [&](auto first_c, uint32_t lev) {
                constexpr bool FIRST = decltype(first_c)::value;
                // forward transform of this sub-team's digit row (polynomial `sub`, level `lev`); level 0 is peeled off the
                // loop so that its operands (accv, with ACC_REG) are dead during the later levels
#if (TFHE_FFT_ABLATE & 1)
                phase_F1a<K, 2>(R, t, sub, 1u, stash, a.tw.twA, [&](uint32_t, uint32_t) { return 0u; });
#else
                if constexpr (!FIRST) phase_F1a<K, 2>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t, uint32_t) { return 0u; });
                else if constexpr (BMMP && ACC_REG) phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t, uint32_t, int k) { return accv[k]; });
                else if constexpr (BMMP) phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return acc[pp * K::N + j]; });
                else if constexpr (ACC_REG) phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(mbase + pp * K::N, j, rot, K::LOGN) - accv[k]; });
                else phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j) { return diff(pp, j, rot); });
#endif
                if constexpr (TAIL16) {
#if TFHE_TMEM_TAIL16_LATEBAR
                    if constexpr (!FIRST) team_bar_id(team_bar, K::TEAM_THREADS);   // the rows published at the previous level have been read (after this level's digits and pass A)
#endif
                    store_Asw16<C>(R.x, buf0, t);        // the one shared-memory exchange
#if TFHE_TMEM_TAIL16_TWB
                    {
                        TwRaw32 r0;
                        cplx tw[8];
                        tmem_tw_block_request(r0, t16_twb);          // R.x is dead here
                        sub_sync();
                        tmem_tw_block_claim<7>(r0, tw);
                        load_Bsw16<C>(R.x, b16_even, b16_odd);
                        fwd_pass_range<4, 0, 3>(R.x, tw);            // stages 4, 5, 6
                        tmem_tw_block_request(r0, t16_twb + 32u);
                        tmem_tw_block_claim<8>(r0, tw);
                        fwd_pass_range<4, 3, 4>(R.x, tw - 7);        // stage 7
                    }
#else
                    sub_sync();
                    {
                        cplx tw[15];
                        derive_pass_tw<4>(tw, t16.pB);
                        load_Bsw16<C>(R.x, b16_even, b16_odd);
                        fwd_pass<4, 4>(R.x, tw);         // stages 4..7
                    }
#endif
                    tmem_fwd_tail16(R.x, t16_swap, t16);  // stages 8, 9 inside the warp: layout F10
                    sub_sync();                          // every thread of the sub-team is past its loads from buf0
                } else {
                store_A<C>(R.x, buf0, t);
                sub_sync();
                }
                if constexpr (TAIL16) {
                } else if constexpr (K::SINGLE_BUF) {           // one buffer: a barrier between every load and the next store
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
                // slot group d of this level (KEYS * HALVES ring slots) = row of polynomial (sub + d) mod P at this sub-team's
                // column position; OWN (d = 0): the transformed digit row is this thread's R.x
                auto row_slots = [&](auto own_c, uint32_t d) {
                    constexpr bool OWN = decltype(own_c)::value;
                    const uint32_t p = (sub + d) % (uint32_t)K::P;
                    const cplx *peer = reinterpret_cast<const cplx *>(tm + K::TM_SUB + p * K::SUB_BYTES + K::STASH_BYTES);
#pragma unroll 1
                    for (uint32_t kh = 0; kh < KEYS * K::HALVES; kh++, it++) {
                        const uint32_t which = kh / K::HALVES, half = kh % K::HALVES, s = it % K::NSLOT;
                        if constexpr (!SELF_REFILL) { if (producer) pump(it + 1, it); }
                        mbar_wait(full + s, (it / K::NSLOT) & 1u, a.err_flag);   // (testing one slot ahead as in mac_slot: P2 103.6 -> 105.9 ms, not kept here)
#if !(TFHE_FFT_ABLATE & 2)
                        const cplx *slot = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES);
                        if constexpr (BMMP) phase_mac_bmmp<K, OWN>(R, t, sub, slot, peer, a.tw.ztab, which == 0 ? ex0 : which == 1 ? rot : rot1, which == 0 ? zb[0] : which == 1 ? zb[1] : zb[2]);
                        else phase_mac<K, OWN>(R, t, sub, slot, peer, half);
#endif
                        __syncwarp();
                        if (lane == 0) release_slot(it);
                    }
                };
#ifndef TFHE_FFT_GEN_OWNFIRST
#define TFHE_FFT_GEN_OWNFIRST 1
#endif
#if TFHE_FFT_GEN_OWNFIRST
                row_slots(std::true_type{}, 0u);         // the own row needs no barrier
                team_bar_id(team_bar, K::TEAM_THREADS);  // all P transformed rows of this level are published
#else
                team_bar_id(team_bar, K::TEAM_THREADS);  // all P transformed rows of this level are published
                row_slots(std::true_type{}, 0u);
#endif
#pragma unroll 1
                for (uint32_t d = 1; d < (uint32_t)K::P; d++) row_slots(std::false_type{}, d);
                if constexpr (!(TAIL16 && TFHE_TMEM_TAIL16_LATEBAR)) team_bar_id(team_bar, K::TEAM_THREADS);  // the published rows have been read: buf0 may be overwritten
            };
            level(std::true_type{}, 0u);
#pragma unroll 1
            for (uint32_t lev = 1; lev < (uint32_t)K::L; lev++) level(std::false_type{}, lev);
        }
#if !(TFHE_FFT_ABLATE & 16)
        // inverse transforms of this sub-team's column: the low- and high-limb products together (fft_team.cuh phase_J*)
        if constexpr (TAIL16) {          // one limb after the other: tail in tensor memory, pass B'', the one exchange, pass A
            uint32_t lo[2 * K::E];
            static_for<0, 2>([&](auto li) {
                constexpr int LIMB = decltype(li)::value;
                tmem_inv_tail16(R.acc[LIMB], t16_swap, t16);
#if TFHE_TMEM_TAIL16_LATEBAR
                if constexpr (LIMB == 0) team_bar_id(team_bar, K::TEAM_THREADS);   // the rows published at the last level have been read: buf0 may be overwritten
#endif
#if TFHE_TMEM_TAIL16_TWB
                {
                    TwRaw32 r0;
                    cplx tw[8];
                    tmem_tw_block_request(r0, t16_twb + 32u);
                    tmem_tw_block_claim<8>(r0, tw);
                    inv_pass_range<4, 3, 4>(R.acc[LIMB], tw - 7);    // stage 7
                    tmem_tw_block_request(r0, t16_twb);
                    tmem_tw_block_claim<7>(r0, tw);
                    inv_pass_range<4, 0, 3>(R.acc[LIMB], tw);        // stages 6, 5, 4
                }
#else
                {
                    cplx tw[15];
                    derive_pass_tw<4>(tw, t16.pB);
                    inv_pass<4, 4>(R.acc[LIMB], tw);
                }
#endif
                if constexpr (LIMB == 1) sub_sync();   // the low limb's loads from buf0 are done
                store_Bsw16<C>(R.acc[LIMB], b16_even, b16_odd);
                sub_sync();
                load_Asw16<C>(R.acc[LIMB], buf0, t);
                inv_pass<C::LOGE, C::LOGE>(R.acc[LIMB], a.tw.twA);
                if constexpr (LIMB == 0) {
#pragma unroll
                    for (int e = 0; e < K::E; e++) {
                        lo[2 * e] = round_u32<K::CHECK>(R.acc[0][e].re, maxfrac);
                        lo[2 * e + 1] = round_u32<K::CHECK>(R.acc[0][e].im, maxfrac);
                    }
                } else {
                    uint32_t *acc_c = acc + sub * K::N;
#pragma unroll
                    for (int e = 0; e < K::E; e++) {
                        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
                        acc_c[j] += lo[2 * e] + (round_u32<K::CHECK>(R.acc[1][e].re, maxfrac) << 16);
                        acc_c[j + K::M] += lo[2 * e + 1] + (round_u32<K::CHECK>(R.acc[1][e].im, maxfrac) << 16);
                    }
                }
            });
            if constexpr (!SELF_REFILL) { if (producer) pump(0, it); }
        } else if constexpr (K::SINGLE_BUF) {   // one limb after the other through the single buffer
            uint32_t lo[2 * K::E];
            phase_K1<K, 0>(R, t, twC, buf0);
            sub_sync();
            phase_K2a<K, 0>(R, jbB, twB, buf0);
            if constexpr (!SELF_REFILL) { if (producer) pump(0, it); }
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
            if constexpr (K::XCHG) {
                tmem_inv_rest(R.acc[0], R.acc[1], taddr, taddr + TMEM_PUB_COL + 32u * pubsel, twx);   // stages 7..3 of both limbs, exchanged through tensor memory: layout A
            } else if constexpr (TAIL) {
                tmem_inv_tail9(R.acc[0], R.acc[1], twt_cols, tailtw);   // stages 8..6 of both limbs inside the warp: layout B''
#if TFHE_TMEM_TAIL_TWB
                TwRaw32 raw;
                cplx tw[8];
                tmem_tw_block_request(raw, tail_twb_cols);
                tmem_tw_block_claim<C::NB_TW>(raw, tw);
#else
                cplx tw[C::NB_TW];
                derive_pass_tw<C::QB>(tw, twB_base);
#endif
                inv_pass<C::LOGE, C::QB>(R.acc[0], tw);
                inv_pass<C::LOGE, C::QB>(R.acc[1], tw);
            } else if constexpr (TWT) {
                TwRaw32 raw;
                cplx tw[8];
                tmem_tw_block_request(raw, twt_cols + 32u);
                tmem_tw_block_claim<C::NC_TW>(raw, tw);
                phase_J1w<K>(R, t, tw, buf0, buf1);
                tmem_tw_block_request(raw, twt_cols);             // the accumulators are stored: dead registers
                sub_sync();
                tmem_tw_block_claim<C::NB_TW>(raw, tw);
                phase_J2aw<K>(R, jbB, tw, buf0, buf1);
            } else if constexpr (OWN_FIRST) {
                phase_J1v<K>(R, t, twC_base, buf0, buf1);
                sub_sync();
                phase_J2av<K>(R, jbB, twB_base, buf0, buf1);
            } else {
                phase_J1<K>(R, t, twC, buf0, buf1);
                sub_sync();
                phase_J2a<K>(R, jbB, twB, buf0, buf1);
                if constexpr (!SELF_REFILL) { if (producer) pump(0, it); }   // ring entries freed by slower teams: refill them while this team inverts
            }
            if constexpr (K::XCHG) {
                phase_J3r_regs<K>(R, t, a.tw.twA, acc + sub * K::N, accv, maxfrac);
            } else if constexpr (TAIL) {
                // (moving the barrier that frees buf0 from the end of the level loop to this point, the inverse's first shared-memory
                // access, was measured neutral: 61.38 vs 61.40 ms)
                store_Bsw<C>(R.acc[0], buf0, jb_swz);
                store_Bsw<C>(R.acc[1], buf1, jb_swz);
                sub_sync();
                phase_J3r_sw<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, accv, maxfrac);
            } else {
                sub_sync();
                phase_J2b<K>(R, jbB, buf0, buf1);
                sub_sync();
                if constexpr (ACC_REG) phase_J3r<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, accv, maxfrac);
                else phase_J3<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, maxfrac);
            }
        }
#endif
        sub_sync();   // acc[sub] (read only by this sub-team) is up to date before the next step's digits
    }
    uint32_t *out = a.glwe_out + ((size_t)ct * K::P + sub) * K::N;
    for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[sub * K::N + idx];
    if constexpr (K::CHECK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxfrac = fmax(maxfrac, __shfl_xor_sync(0xFFFFFFFFu, maxfrac, o));
        if (lane == 0) atomicMax(a.margin, (unsigned long long)__double_as_longlong(maxfrac));
    }
    if constexpr (USE_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 15, %0;" ::"n"(K::THREADS) : "memory");   // every warp of the CTA (idle teams arrive before they exit)
        if (tid < 32) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr), "n"(TMEM_ALLOC) : "memory");
        }
    }
}

// the key of the tensor-memory-exchange path: the same numbers as the shared-memory path's key, every polynomial re-ordered from
// slot order (point (t << 3) | e at e * 32 + t) to the spectral layout F of fft_tmem.cuh.  grid = polynomials (M points each).
__global__ void bsk_fft_reslot_kernel(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t poly = blockIdx.x;
    if (poly >= polys) return;
    const uint32_t idx = threadIdx.x;                    // position in slot order: e * 32 + t
    const uint32_t j = ((idx & 31u) << 3) | (idx >> 5);   // the point it holds
    out[poly * 256 + tmem_slot_of_index(j)] = in[poly * 256 + idx];
}

// the same for the tensor-memory-tail path (M = 512): slot order (point (t << 3) | e at e * 64 + t) -> layout F9
__global__ void bsk_fft_reslot9_kernel(const cplx *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t poly = blockIdx.x;
    if (poly >= polys) return;
    const uint32_t idx = threadIdx.x;                    // position in slot order: e * 64 + t
    const uint32_t j = ((idx & 63u) << 3) | (idx >> 6);   // the point it holds
    out[poly * 512 + tail9_slot_of_index(j)] = in[poly * 512 + idx];
}

// the same for N = 2048 (M = 1024, 16 points per thread, key slots of half a row): the polynomial of (row, limb, column) lives in two
// slots, point (t << 4) | e at slot e / 8, position (e % 8) * 64 + t; layout F10 puts point j into slot j5, position (j1 j0 j4) * 64 +
// thread(j).  grid = (rows, 2 * P polynomials per row), 1024 threads = one per point.
__global__ void bsk_fft_reslot10_kernel(const cplx *__restrict__ in, cplx *__restrict__ out, uint32_t polys_per_slot) {
    const size_t row = blockIdx.x, poly = blockIdx.y;                  // poly = limb * P + column
    const size_t slot_elems = (size_t)polys_per_slot * 512;
    const uint32_t idx = threadIdx.x, e = ((idx >> 9) << 3) | ((idx >> 6) & 7u), t = idx & 63u, j = (t << 4) | e;
    const uint32_t r = tail10_reg_of_index(j);
    const size_t src = (row * 2 + (idx >> 9)) * slot_elems + poly * 512 + (idx & 511u);
    const size_t dst = (row * 2 + (r >> 3)) * slot_elems + poly * 512 + (r & 7u) * 64u + tail10_thread_of_index(j);
    out[dst] = in[src];
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
    const size_t row = (step * K::ROWS + key_slot_index<K>((uint32_t)(r / K::L), (uint32_t)c, (uint32_t)(r % K::L))) * kps + which;   // consumption order, diagonal-major
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
