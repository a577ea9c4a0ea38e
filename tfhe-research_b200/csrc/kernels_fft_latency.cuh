// kernels_fft_latency.cuh -- blind rotation of ONE ciphertext per CTA with all teams of the CTA working on it
// (FFT path; batches of at most one ciphertext per SM: the "per-PBS latency" metric).
//
// The throughput kernel (kernels_fft.cuh) gives every ciphertext one team of P sub-teams and keeps 2-4 ciphertexts per SM
// busy; a lone ciphertext then walks its n CMUX steps as one dependent chain of L levels + one inverse per step on a mostly
// idle SM.  Here the L levels of a step -- L independent forward transforms per polynomial and their multiply-accumulates --
// are spread over the CTS teams of the CTA:
//   * team 0 (the owner) holds the accumulator, decomposes it (decomposer.rs:42-80) and stashes the digits of all levels;
//   * team h transforms the digit rows of levels h, h + CTS, ... (its sub-team s: polynomial s) and multiplies them with
//     column s of those levels' GGSW rows into its own register accumulators (same phases as the throughput kernel:
//     own row first, then the rows published by the other sub-teams of the team);
//   * the helpers hand their partial sums to the owner through their (now idle) exchange buffers; the owner adds them,
//     runs the paired inverse transforms, rounds and updates the accumulator (ggsw.rs:132-178, bootstrapping.rs:90-105).
// Per step: two CTA barriers (digits ready / partial sums ready); the critical path is ceil(L / CTS) levels instead of L.
// The partial sums change the order of the floating-point additions, not the result: every rounded value stays within the
// a-priori bound of an integer, so the output bits are those of the throughput kernel and of the oracle
// (tests/test_gpu_exactness.py::test_latency_configuration_*).
//
// Key stream: one ring of NSLOT whole GGSW rows, filled by TMA in the order the teams consume them (round r, slot d of the
// level, team): every ring entry is read by ONE team, whose last warp to release it issues the copy that reuses it.
#pragma once
#include "kernels_fft.cuh"

namespace tfhe {
namespace fft {

template <class K>
struct LatencyLayout {
    using C = typename K::F;
    static_assert(K::HALVES == 1 && !K::SINGLE_BUF, "latency kernel: whole-row key slots, two exchange buffers per sub-team");
    static constexpr int FULL_ROUNDS = K::L / K::CTS, LAST_TEAMS = K::L % K::CTS, ROUNDS = FULL_ROUNDS + (LAST_TEAMS ? 1 : 0);
    static constexpr int ACC = 0;                                        // u32 acc[P][N]
    static constexpr int STASH = ACC + K::P * K::N * 4;                  // P x STASH_BYTES (owner's digits of levels >= 1)
    static constexpr int BUFS = STASH + K::P * K::STASH_BYTES;           // cplx [CTS][P][2][MPAD]
    static constexpr int SUBBUF_BYTES = 2 * C::MPAD * 16;
    static constexpr int AT = BUFS + K::CTS * K::P * SUBBUF_BYTES;       // u16 at[n+1]
    static constexpr size_t ring_offset(size_t n) { return ((size_t)AT + (n + 1) * 2 + 127) & ~(size_t)127; }
    static constexpr size_t smem_bytes(size_t n) { return ring_offset(n) + (size_t)K::NSLOT * K::SLOT_BYTES + 2 * K::NSLOT * 8 + 4 * K::NSLOT + 16; }
    // position q of a step's key stream -> storage row (level-major: lev * P + d) of the row consumed there
    __device__ static uint32_t row_of_position(uint32_t q) {
        constexpr uint32_t FULL = (uint32_t)(FULL_ROUNDS * K::P * K::CTS);
        uint32_t r, d, team;
        if (q < FULL) {
            r = q / (uint32_t)(K::P * K::CTS);
            const uint32_t rem = q % (uint32_t)(K::P * K::CTS);
            d = rem / (uint32_t)K::CTS;
            team = rem % (uint32_t)K::CTS;
        } else {
            constexpr uint32_t NT = LAST_TEAMS ? LAST_TEAMS : 1;
            r = (uint32_t)FULL_ROUNDS;
            d = (q - FULL) / NT;
            team = (q - FULL) % NT;
        }
        return (r * (uint32_t)K::CTS + team) * (uint32_t)K::P + d;
    }
    // position of (round r, slot d of the level, team) inside a step
    __device__ static uint32_t position(uint32_t r, uint32_t d, uint32_t team) {
        const uint32_t nt = r < (uint32_t)FULL_ROUNDS ? (uint32_t)K::CTS : (uint32_t)LAST_TEAMS;
        return r * (uint32_t)(K::P * K::CTS) + d * nt + team;
    }
};

template <class K>
__global__ void __launch_bounds__(K::THREADS, 1) pbs_fft_latency_kernel(const __grid_constant__ FftArgs a) {
    using C = typename K::F;
    using LL = LatencyLayout<K>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t team = tid / K::TEAM_THREADS, tt = tid % K::TEAM_THREADS, sub = tt / K::T, t = tt % K::T;
    const uint32_t ct = blockIdx.x;
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem + LL::ACC);
    typename K::stash_t *stash = reinterpret_cast<typename K::stash_t *>(smem + LL::STASH + sub * K::STASH_BYTES);   // the OWNER's digits of polynomial `sub`
    auto subbuf = [&](uint32_t tm_, uint32_t sb_) { return reinterpret_cast<cplx *>(smem + LL::BUFS + (tm_ * K::P + sb_) * LL::SUBBUF_BYTES); };
    cplx *buf0 = subbuf(team, sub), *buf1 = buf0 + C::MPAD;
    uint16_t *at = reinterpret_cast<uint16_t *>(smem + LL::AT);
    uint8_t *ring = smem + LL::ring_offset(a.n);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES), *empty = full + K::NSLOT;
    uint32_t *claimed = reinterpret_cast<uint32_t *>(empty + K::NSLOT);
    constexpr bool WARP_SUB = K::T == 32;
    const uint32_t team_bar = 1 + team * (WARP_SUB ? 1 : K::P + 1), sub_bar = team_bar + 1 + sub;
    static_assert(K::CTS * (WARP_SUB ? 1 : K::P + 1) <= 15, "named barrier ids");
    auto sub_sync = [&]() {
        if constexpr (WARP_SUB) __syncwarp();
        else team_bar_id(sub_bar, K::T);
    };
    const uint32_t jbB = jbase_B<C>(t);
    const cplx twB_base = pass_tw_base<C::QB>(a.tw.twB + (t >> C::QB) * C::NB_TW, 1);
    const cplx twC_base = pass_tw_base<C::LOGE>(a.tw.twC + t, C::T);
    const uint32_t total_pos = a.n * (uint32_t)K::ROWS;

    if (tid == 0) {
        for (int s = 0; s < K::NSLOT; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, K::P * K::WARPS_PER_SUB);   // every ring entry is read by the warps of ONE team
            claimed[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // utils.rs:23-33 mod switch; acc = trivial GLWE of the encoded test vector times X^{-b~} (bootstrapping.rs:79-86)
    const uint32_t *lwe = a.lwe_in + (size_t)ct * (a.n + 1);
    for (uint32_t i = tid; i <= a.n; i += K::THREADS) at[i] = (uint16_t)mod_switch(__ldg(lwe + i), K::LOGN);
    __syncthreads();
    {
        const uint32_t b = at[a.n];
        uint32_t li = a.lut_idx ? __ldg(a.lut_idx + ct) : 0u;
        if (li >= a.n_luts) {
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
    }
    __syncthreads();

    const uint8_t *ksrc = reinterpret_cast<const uint8_t *>(a.bsk_fft);
    auto issue_pos = [&](uint32_t g) {   // global stream position g -> TMA copy of its row into ring entry g % NSLOT
        const uint32_t s = g % K::NSLOT;
        const size_t row = (size_t)(g / (uint32_t)K::ROWS) * K::ROWS + LL::row_of_position(g % (uint32_t)K::ROWS);
        mbar_expect_tx(full + s, K::SLOT_BYTES);
        bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + row * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
        bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + row * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
    };
    auto release_pos = [&](uint32_t g) {   // lane 0 of a warp that is done with position g
        const uint32_t s = g % K::NSLOT, u = g / K::NSLOT, nx = g + (uint32_t)K::NSLOT;
        mbar_arrive(empty + s);
        if (nx < total_pos && mbar_test(empty + s, u & 1u)) {
            if (atomicCAS(claimed + s, u, u + 1u) == u) issue_pos(nx);
        }
    };
    if (tid == 0)
        for (uint32_t g = 0; g < (uint32_t)K::NSLOT && g < total_pos; g++) issue_pos(g);

    FftRegs<K> R;
    double maxfrac = 0.0;
    uint32_t accv[2 * K::E];   // owner: the 2E words of acc[sub] this thread decomposes and updates
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc[sub * K::N + j];
        accv[2 * e + 1] = acc[sub * K::N + j + K::M];
    }
    // teams that have work: team h runs levels h, h + CTS, ...
    constexpr uint32_t NTEAMS = K::CTS < K::L ? K::CTS : K::L;
    const bool works = team < NTEAMS;

#pragma unroll 1
    for (uint32_t i = 0; i < a.n; i++) {
        const uint32_t rot = at[i], base = i * (uint32_t)K::ROWS;
        if (rot == 0) {
            // diff == 0 => external product == 0 exactly: every team consumes its ring entries of this step without using them
            if (works) {
#pragma unroll 1
                for (uint32_t r = 0; r < (uint32_t)LL::ROUNDS; r++) {
                    if (r * K::CTS + team >= (uint32_t)K::L) break;
#pragma unroll 1
                    for (uint32_t d = 0; d < (uint32_t)K::P; d++) {
                        const uint32_t g = base + LL::position(r, d, team);
                        mbar_wait(full + (g % K::NSLOT), (g / K::NSLOT) & 1u, a.err_flag);
                        __syncwarp();
                        if (lane == 0) release_pos(g);
                    }
                }
            }
            continue;
        }
        zero_acc<K>(R);
        if (team == 0)   // decompose polynomial `sub` of rot(acc) - acc: level 0 stays in R.x (pass A done), the other levels go to the stash
            phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(acc + pp * K::N, j, rot, K::LOGN) - accv[k]; });
        __syncthreads();   // B1: the digits of this step are in the stash
        if (works) {
#pragma unroll 1
            for (uint32_t r = 0; r < (uint32_t)LL::ROUNDS; r++) {
                const uint32_t lev = r * K::CTS + team;
                if (lev >= (uint32_t)K::L) break;
                if (lev != 0) phase_F1a<K, 2>(R, t, sub, lev, stash, a.tw.twA, [&](uint32_t, uint32_t) { return 0u; });
                if (r != 0) team_bar_id(team_bar, K::TEAM_THREADS);   // the row published in the previous round has been read
                store_A<C>(R.x, buf0, t);
                sub_sync();
                phase_F2v<K>(R, jbB, twB_base, buf0, buf1);
                sub_sync();
                phase_F3v<K>(R, t, twC_base, buf1);
                phase_xstore<K>(R, t, buf0);
                {   // own row: slot 0 of the level
                    const uint32_t g = base + LL::position(r, 0u, team), s = g % K::NSLOT;
                    mbar_wait(full + s, (g / K::NSLOT) & 1u, a.err_flag);
                    phase_mac<K, true>(R, t, sub, reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES), buf0, 0u);
                    __syncwarp();
                    if (lane == 0) release_pos(g);
                }
                team_bar_id(team_bar, K::TEAM_THREADS);   // all P transformed rows of this level are published
#pragma unroll 1
                for (uint32_t d = 1; d < (uint32_t)K::P; d++) {
                    const uint32_t g = base + LL::position(r, d, team), s = g % K::NSLOT;
                    mbar_wait(full + s, (g / K::NSLOT) & 1u, a.err_flag);
                    phase_mac<K, false>(R, t, sub, reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES), subbuf(team, (sub + d) % (uint32_t)K::P), 0u);
                    __syncwarp();
                    if (lane == 0) release_pos(g);
                }
            }
            team_bar_id(team_bar, K::TEAM_THREADS);       // the last published rows have been read: the buffers are free
            if (team != 0) {                              // hand the partial sums (column `sub`, both limbs) to the owner
                store_C<C>(R.acc[0], buf0, t);
                store_C<C>(R.acc[1], buf1, t);
            }
        }
        __syncthreads();   // B2: partial sums are in the helpers' buffers
        if (team == 0) {
#pragma unroll 1
            for (uint32_t h = 1; h < NTEAMS; h++) {
                const cplx *p0 = subbuf(h, sub), *p1 = p0 + C::MPAD;
                cplx v[K::E];
                load_C<C>(v, p0, t);
#pragma unroll
                for (int e = 0; e < K::E; e++) { R.acc[0][e].re = add_d(R.acc[0][e].re, v[e].re); R.acc[0][e].im = add_d(R.acc[0][e].im, v[e].im); }
                load_C<C>(v, p1, t);
#pragma unroll
                for (int e = 0; e < K::E; e++) { R.acc[1][e].re = add_d(R.acc[1][e].re, v[e].re); R.acc[1][e].im = add_d(R.acc[1][e].im, v[e].im); }
            }
            phase_J1v<K>(R, t, twC_base, buf0, buf1);
            sub_sync();
            phase_J2av<K>(R, jbB, twB_base, buf0, buf1);
            sub_sync();
            phase_J2b<K>(R, jbB, buf0, buf1);
            sub_sync();
            phase_J3r<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, accv, maxfrac);
            team_bar_id(team_bar, K::TEAM_THREADS);       // acc is up to date for every sub-team's rotated reads of the next step
        }
    }
    if (team == 0) {
        uint32_t *out = a.glwe_out + ((size_t)ct * K::P + sub) * K::N;
        for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[sub * K::N + idx];
    }
    (void)maxfrac;
}

}  // namespace fft
}  // namespace tfhe
