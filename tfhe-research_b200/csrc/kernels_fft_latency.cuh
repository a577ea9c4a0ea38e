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
// Key stream: every team has ONE shared-memory slot for a whole GGSW row, with its own full/empty mbarriers and its own
// sequence of rows (round r of its levels, slot d of the level); the last warp of the team to release a row issues the TMA
// copy of the team's next row -- possibly the first row of the next step, which then lands while the owner inverts.  (A slot
// is never shared between teams: mbarrier parity waits are only unambiguous for consumers at most one phase apart.)
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
    static_assert(K::NSLOT == K::CTS, "one key slot per team");
    // team h runs levels h, h + CTS, ...: rows it consumes per step, and the storage row (level-major: lev * P + d) of the
    // seq-th row of its sequence
    __device__ static uint32_t levels_of(uint32_t team) { return (uint32_t)FULL_ROUNDS + (team < (uint32_t)LAST_TEAMS ? 1u : 0u); }
    __device__ static size_t row_of(uint32_t team, uint32_t seq) {
        const uint32_t rps = levels_of(team) * (uint32_t)K::P, step = seq / rps, q = seq % rps;
        const uint32_t r = q / (uint32_t)K::P, d = q % (uint32_t)K::P;
        return (size_t)step * K::ROWS + (r * (uint32_t)K::CTS + team) * (uint32_t)K::P + d;
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
    // teams that have work: team h runs levels h, h + CTS, ...
    constexpr uint32_t NTEAMS = K::CTS < K::L ? K::CTS : K::L;
    const bool works = team < NTEAMS;
    const uint32_t my_levels = LL::levels_of(team), my_rows = a.n * my_levels * (uint32_t)K::P;   // rows this team consumes in all

    if (tid == 0) {
        for (int s = 0; s < K::NSLOT; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, K::P * K::WARPS_PER_SUB);   // slot s is read by the warps of team s only
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
    auto issue_row = [&](uint32_t tm_, uint32_t seq) {   // TMA copy of the seq-th row of team tm_'s sequence into the team's slot
        const size_t row = LL::row_of(tm_, seq);
        mbar_expect_tx(full + tm_, K::SLOT_BYTES);
        bulk_g2s(ring + tm_ * K::SLOT_BYTES, ksrc + row * K::SLOT_BYTES, K::LIMB_BYTES, full + tm_);
        bulk_g2s(ring + tm_ * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + row * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + tm_);
    };
    uint32_t seq = 0;   // rows of my team's sequence consumed so far (the same in every warp of the team)
    auto wait_row = [&]() { mbar_wait(full + team, seq & 1u, a.err_flag); };
    auto release_row = [&]() {   // lane 0 of a warp that is done with the team's current row; advances seq in every lane's copy via the caller
        mbar_arrive(empty + team);
        if (seq + 1u < my_rows && mbar_test(empty + team, seq & 1u)) {
            if (atomicCAS(claimed + team, seq, seq + 1u) == seq) issue_row(team, seq + 1u);
        }
    };
    const cplx *slot = reinterpret_cast<const cplx *>(ring + team * K::SLOT_BYTES);
    if (tid == 0)
        for (uint32_t h = 0; h < NTEAMS; h++) issue_row(h, 0u);

    // measurement only (a.prof != nullptr): cycles of thread 0 of CTA 0 per phase: [0] decompose, [1] wait B1, [2] levels, [3] wait B2,
    // [4] add partial sums, [5] inverse + update, [6] whole loop; of thread 0 of team 1: [8] wait B1, [9] levels, [10] wait B2
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tt == 0 && team < 2;
    unsigned long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [8] forward transform, [9] wait own row, [10] mac own, [11] team barrier, [12] wait peer row, [13] mac peer
    long long tprev = prof ? clock64() : 0;
    auto tick = [&](int k) {
        if (prof) {
            const long long now = clock64();
            pc[k] += (unsigned long long)(now - tprev);
            tprev = now;
        }
    };
    const long long tstart = tprev;
    FftRegs<K> R;
    double maxfrac = 0.0;
    uint32_t accv[2 * K::E];   // owner: the 2E words of acc[sub] this thread decomposes and updates
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc[sub * K::N + j];
        accv[2 * e + 1] = acc[sub * K::N + j + K::M];
    }
#pragma unroll 1
    for (uint32_t i = 0; i < a.n; i++) {
        const uint32_t rot = at[i];
        // the GGSW two steps ahead -> L2 (a lone ciphertext streams the key from HBM exactly once: without this every ring
        // refill pays the DRAM round trip on the critical path); one bulk prefetch per row, issued by one thread per team
        if (tt == 0 && works && i + 2u < a.n) {
#pragma unroll 1
            for (uint32_t q = team; q < (uint32_t)K::ROWS; q += NTEAMS) {
                const uint8_t *src = ksrc + ((size_t)(i + 2u) * K::ROWS + q) * K::SLOT_BYTES;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)K::SLOT_BYTES) : "memory");
            }
        }
        if (rot == 0) {
            // diff == 0 => external product == 0 exactly: every team consumes its ring entries of this step without using them
            if (works) {
#pragma unroll 1
                for (uint32_t q = 0; q < my_levels * (uint32_t)K::P; q++) {
                    wait_row();
                    __syncwarp();
                    if (lane == 0) release_row();
                    seq++;
                }
            }
            continue;
        }
        zero_acc<K>(R);
        if (team == 0)   // decompose polynomial `sub` of rot(acc) - acc: level 0 stays in R.x (pass A done), the other levels go to the stash
            phase_F1a<K, 1>(R, t, sub, 0u, stash, a.tw.twA, [&](uint32_t pp, uint32_t j, int k) { return rot_coeff(acc + pp * K::N, j, rot, K::LOGN) - accv[k]; });
        tick(0);
        __syncthreads();   // B1: the digits of this step are in the stash
        tick(1);
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
                tick(8);
                wait_row();   // own row: slot 0 of the level
                tick(9);
                phase_mac<K, true>(R, t, sub, slot, buf0, 0u);
                __syncwarp();
                if (lane == 0) release_row();
                seq++;
                tick(10);
                team_bar_id(team_bar, K::TEAM_THREADS);   // all P transformed rows of this level are published
                tick(11);
#pragma unroll 1
                for (uint32_t d = 1; d < (uint32_t)K::P; d++) {
                    wait_row();
                    tick(12);
                    phase_mac<K, false>(R, t, sub, slot, subbuf(team, (sub + d) % (uint32_t)K::P), 0u);
                    __syncwarp();
                    if (lane == 0) release_row();
                    seq++;
                    tick(13);
                }
            }
            team_bar_id(team_bar, K::TEAM_THREADS);       // the last published rows have been read: the buffers are free
            if (team != 0) {                              // hand the partial sums (column `sub`, both limbs) to the owner
                store_C<C>(R.acc[0], buf0, t);
                store_C<C>(R.acc[1], buf1, t);
            }
        }
        tick(2);
        __syncthreads();   // B2: partial sums are in the helpers' buffers
        tick(3);
        if (team == 0) {
#pragma unroll
            for (uint32_t h = 1; h < NTEAMS; h++) {   // both limbs of a helper are requested before the first is added
                const cplx *p0 = subbuf(h, sub), *p1 = p0 + C::MPAD;
                cplx v0[K::E], v1[K::E];
                load_C<C>(v0, p0, t);
                load_C<C>(v1, p1, t);
#pragma unroll
                for (int e = 0; e < K::E; e++) { R.acc[0][e].re = add_d(R.acc[0][e].re, v0[e].re); R.acc[0][e].im = add_d(R.acc[0][e].im, v0[e].im); }
#pragma unroll
                for (int e = 0; e < K::E; e++) { R.acc[1][e].re = add_d(R.acc[1][e].re, v1[e].re); R.acc[1][e].im = add_d(R.acc[1][e].im, v1[e].im); }
            }
            tick(4);
            phase_J1v<K>(R, t, twC_base, buf0, buf1);
            sub_sync();
            phase_J2av<K>(R, jbB, twB_base, buf0, buf1);
            sub_sync();
            phase_J2b<K>(R, jbB, buf0, buf1);
            sub_sync();
            phase_J3r<K>(R, t, a.tw.twA, buf0, buf1, acc + sub * K::N, accv, maxfrac);
            team_bar_id(team_bar, K::TEAM_THREADS);       // acc is up to date for every sub-team's rotated reads of the next step
            tick(5);
        }
    }
    if (prof) {
        pc[6] = (unsigned long long)(clock64() - tstart);
        for (int k = 0; k < 16; k++) a.prof[team * 16 + k] = pc[k];
    }
    if (team == 0) {
        uint32_t *out = a.glwe_out + ((size_t)ct * K::P + sub) * K::N;
        for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[sub * K::N + idx];
    }
    (void)maxfrac;
}

}  // namespace fft
}  // namespace tfhe
