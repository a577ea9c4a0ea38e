// kernels_fft_cluster.cuh -- blind rotation of ONE ciphertext by a thread-block CLUSTER of L CTAs, one gadget level each
// (FFT path; batches of at most sm_count / L ciphertexts: the "per-PBS latency" metric).
//
// A lone ciphertext is a chain of n dependent CMUX steps.  Inside one SM its step cannot go faster than the SM's FP64 pipe
// (measured: with all teams of a CTA on one ciphertext the forward transforms of a step already saturate it,
// profiles/r02_latency_phases.txt), while 147 other SMs idle.  Here the L levels of a step run on L SMs:
//   * every CTA of the cluster holds a REPLICA of the GLWE accumulator and decomposes it itself (decomposer.rs:42-80; the
//     replicas stay bit-identical because they only ever add the same integers), keeping the digits of ITS level;
//   * it transforms its P digit rows, multiplies them with its level's P GGSW rows (streamed by its own TMA ring),
//     inverse-transforms ITS partial sum and rounds it: the partial external product of one level is itself an exact integer
//     vector (a sum of integer convolutions), so rounding per level is as exact as rounding the total;
//   * the u32 partial results are exchanged through distributed shared memory: every CTA PUSHES its partial result into the
//     shared memory of all CTAs of the cluster (st.shared::cluster, fire and forget), one cluster barrier (release / acquire)
//     makes them visible, and every CTA adds the L partials from its LOCAL shared memory to its replica -- wrapping u32
//     additions, order-independent (ggsw.rs:132-178: out[c] = sum over rows, + acc).  (Pulling the peers' partials with
//     ld.shared::cluster after the barrier put the remote round trips on the critical path: 1.5k cycles per step.)
// Per step: one cluster barrier; partial buffers are double-buffered so the barrier of step i+1 also frees the buffer of step i.
#pragma once
#include "kernels_fft.cuh"

namespace tfhe {
namespace fft {

template <class K>
struct ClusterLayout {
    using C = typename K::F;
    static_assert(K::CTS == 1 && K::HALVES == 1 && !K::SINGLE_BUF, "cluster kernel: one team per CTA, whole-row key slots, two exchange buffers");
    static_assert(K::L <= 8, "one CTA per level: portable cluster size");
    static constexpr int ACC = 0;                                        // u32 acc[P][N]          (replica)
    static constexpr int PART = ACC + K::P * K::N * 4;                   // u32 part[2][L][P][N]   (partial results of ALL levels, two executed steps)
    static constexpr int BUFS = PART + 2 * K::L * K::P * K::N * 4;       // cplx [P][2][MPAD]
    static constexpr int SUBBUF_BYTES = 2 * C::MPAD * 16;
    static constexpr int AT = BUFS + K::P * SUBBUF_BYTES;                // u16 at[n+1]
    static constexpr size_t ring_offset(size_t n) { return ((size_t)AT + (n + 1) * 2 + 127) & ~(size_t)127; }
    static constexpr size_t smem_bytes(size_t n) { return ring_offset(n) + (size_t)K::NSLOT * K::SLOT_BYTES + 2 * K::NSLOT * 8 + 4 * K::NSLOT + 16; }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <class K>
__global__ void __launch_bounds__(K::TEAM_THREADS, 1) pbs_fft_cluster_kernel(const __grid_constant__ FftArgs a) {
    using C = typename K::F;
    using LL = ClusterLayout<K>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, sub = tid / K::T, t = tid % K::T;
    const uint32_t lev = cluster_ctarank();            // this CTA's gadget level
    const uint32_t ct = blockIdx.x / (uint32_t)K::L;
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem + LL::ACC);
    uint32_t *part = reinterpret_cast<uint32_t *>(smem + LL::PART);
    auto subbuf = [&](uint32_t sb_) { return reinterpret_cast<cplx *>(smem + LL::BUFS + sb_ * LL::SUBBUF_BYTES); };
    cplx *buf0 = subbuf(sub), *buf1 = buf0 + C::MPAD;
    uint16_t *at = reinterpret_cast<uint16_t *>(smem + LL::AT);
    uint8_t *ring = smem + LL::ring_offset(a.n);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES), *empty = full + K::NSLOT;
    uint32_t *claimed = reinterpret_cast<uint32_t *>(empty + K::NSLOT);
    constexpr bool WARP_SUB = K::T == 32;
    constexpr uint32_t TEAM_BAR = 1;
    const uint32_t sub_bar = 2 + sub;
    auto sub_sync = [&]() {
        if constexpr (WARP_SUB) __syncwarp();
        else team_bar_id(sub_bar, K::T);
    };
    auto team_sync = [&]() { team_bar_id(TEAM_BAR, K::TEAM_THREADS); };
    const uint32_t jbB = jbase_B<C>(t);
    const cplx twB_base = pass_tw_base<C::QB>(a.tw.twB + (t >> C::QB) * C::NB_TW, 1);
    const cplx twC_base = pass_tw_base<C::LOGE>(a.tw.twC + t, C::T);
    const uint32_t total_rows = a.n * (uint32_t)K::P;   // rows this CTA consumes: P per step (slots d = 0..P-1 of its level)

    if (tid == 0) {
        for (int s = 0; s < K::NSLOT; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, K::P * K::WARPS_PER_SUB);
            claimed[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // utils.rs:23-33 mod switch; acc = trivial GLWE of the encoded test vector times X^{-b~} (bootstrapping.rs:79-86) -- in every CTA
    const uint32_t *lwe = a.lwe_in + (size_t)ct * (a.n + 1);
    for (uint32_t i = tid; i <= a.n; i += K::TEAM_THREADS) at[i] = (uint16_t)mod_switch(__ldg(lwe + i), K::LOGN);
    __syncthreads();
    {
        const uint32_t b = at[a.n];
        uint32_t li = a.lut_idx ? __ldg(a.lut_idx + ct) : 0u;
        if (li >= a.n_luts) {
            atomicOr(a.err_flag, 4u);
            li = 0u;
        }
        const uint32_t *lut = a.luts + (size_t)li * K::N;
        for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += K::TEAM_THREADS) {
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

    // key stream of this level: row q of the CTA's sequence = (step q / P, slot d = q % P); all warps consume every ring entry in order,
    // the last warp to release an entry issues the copy that reuses it
    const uint8_t *ksrc = reinterpret_cast<const uint8_t *>(a.bsk_fft);
    auto issue_row = [&](uint32_t q) {
        const uint32_t s = q % K::NSLOT;
        const size_t row = (size_t)(q / (uint32_t)K::P) * K::ROWS + (size_t)lev * K::P + q % (uint32_t)K::P;
        mbar_expect_tx(full + s, K::SLOT_BYTES);
        bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + row * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
        bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + row * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
    };
    auto release_row = [&](uint32_t q) {   // lane 0 of a warp that is done with row q
        const uint32_t s = q % K::NSLOT, u = q / K::NSLOT, nx = q + (uint32_t)K::NSLOT;
        mbar_arrive(empty + s);
        if (nx < total_rows && mbar_test(empty + s, u & 1u)) {
            if (atomicCAS(claimed + s, u, u + 1u) == u) issue_row(nx);
        }
    };
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)K::NSLOT && q < total_rows; q++) issue_row(q);

    // measurement only (a.prof != nullptr): cycles of thread 0 of the first CTA per phase: [0] digits + pass A, [1] rest of the forward
    // transform, [2] wait own row, [3] mac own, [4] team barrier, [5] wait + mac peer rows, [6] inverse + rounding, [7] cluster barrier,
    // [8] accumulate through distributed shared memory, [9] whole loop
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    unsigned long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
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
    uint32_t accv[2 * K::E];   // the 2E words of acc[sub] this thread decomposes and updates
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc[sub * K::N + j];
        accv[2 * e + 1] = acc[sub * K::N + j + K::M];
    }
    uint32_t q = 0;      // position in this CTA's key stream
    uint32_t nexec = 0;  // executed (not skipped) steps: selects the partial-result buffer
    cluster_sync_all();   // every CTA of the cluster is resident and initialised before anyone reads a peer's shared memory

#pragma unroll 1
    for (uint32_t i = 0; i < a.n; i++) {
        const uint32_t rot = at[i];
        if (rot == 0) {   // diff == 0 => external product == 0 exactly, in every CTA of the cluster alike: consume this step's rows
#pragma unroll 1
            for (uint32_t d = 0; d < (uint32_t)K::P; d++, q++) {
                mbar_wait(full + (q % K::NSLOT), (q / K::NSLOT) & 1u, a.err_flag);
                __syncwarp();
                if (lane == 0) release_row(q);
            }
            continue;
        }
        zero_acc<K>(R);
        // digits of level `lev` of rot(acc) - acc, polynomial `sub` (decomposer.rs:27-80: all L digits are computed, one is kept)
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            int32_t dl[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t j = (((uint32_t)e << C::LOGT) | t) + (uint32_t)h * K::M;
                int32_t d[K::L];
                decompose_signed<K::LOGB, K::L>(rot_coeff(acc + sub * K::N, j, rot, K::LOGN) - accv[2 * e + h], d);
                dl[h] = d[0];
#pragma unroll
                for (int l = 1; l < K::L; l++) dl[h] = lev == (uint32_t)l ? d[l] : dl[h];
            }
            R.x[e] = cplx{i2d(dl[0]), i2d(dl[1])};
        }
        fwd_pass<C::LOGE, C::LOGE>(R.x, a.tw.twA);
        tick(0);
        store_A<C>(R.x, buf0, t);
        sub_sync();
        phase_F2v<K>(R, jbB, twB_base, buf0, buf1);
        sub_sync();
        phase_F3v<K>(R, t, twC_base, buf1);
        phase_xstore<K>(R, t, buf0);
        tick(1);
        {   // own row: slot 0 of the level
            const uint32_t s = q % K::NSLOT;
            mbar_wait(full + s, (q / K::NSLOT) & 1u, a.err_flag);
            tick(2);
            phase_mac<K, true>(R, t, sub, reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES), buf0, 0u);
            __syncwarp();
            if (lane == 0) release_row(q);
            q++;
        }
        tick(3);
        team_sync();   // all P transformed rows of this level are published
        tick(4);
#pragma unroll 1
        for (uint32_t d = 1; d < (uint32_t)K::P; d++, q++) {
            const uint32_t s = q % K::NSLOT;
            mbar_wait(full + s, (q / K::NSLOT) & 1u, a.err_flag);
            phase_mac<K, false>(R, t, sub, reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES), subbuf((sub + d) % (uint32_t)K::P), 0u);
            __syncwarp();
            if (lane == 0) release_row(q);
        }
        team_sync();   // the published rows have been read: the buffers are free
        tick(5);
        // inverse transforms of this level's partial sum (column `sub`, both limbs), rounded to the exact integers
        phase_J1v<K>(R, t, twC_base, buf0, buf1);
        sub_sync();
        phase_J2av<K>(R, jbB, twB_base, buf0, buf1);
        sub_sync();
        phase_J2b<K>(R, jbB, buf0, buf1);
        sub_sync();
        // partial results of executed step `nexec` live in part[nexec & 1][level][sub][j] of EVERY CTA; a buffer is rewritten two
        // executed steps later, i.e. one cluster barrier after its last reader
        uint32_t *stepbuf = part + (size_t)(nexec & 1u) * K::L * K::P * K::N;
        uint32_t *mine = stepbuf + ((size_t)lev * K::P + sub) * K::N;
        nexec++;
        load_A<C>(R.acc[0], buf0, t);
        load_A<C>(R.acc[1], buf1, t);
        inv_pass<C::LOGE, C::LOGE>(R.acc[0], a.tw.twA);
        inv_pass<C::LOGE, C::LOGE>(R.acc[1], a.tw.twA);
        uint32_t pv[2 * K::E];
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            pv[2 * e] = round_u32<false>(R.acc[0][e].re, maxfrac) + (round_u32<false>(R.acc[1][e].re, maxfrac) << 16);
            pv[2 * e + 1] = round_u32<false>(R.acc[0][e].im, maxfrac) + (round_u32<false>(R.acc[1][e].im, maxfrac) << 16);
        }
        const uint32_t mine_addr = smem_u32(mine);
#pragma unroll
        for (uint32_t r = 1; r < (uint32_t)K::L; r++) {   // push to the peers (same offset in their shared memory)
            uint32_t peer = lev + r;
            peer = peer >= (uint32_t)K::L ? peer - (uint32_t)K::L : peer;
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine_addr), "r"(peer));
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) {
                const uint32_t j = (((uint32_t)(k >> 1) << C::LOGT) | t) + (uint32_t)(k & 1) * K::M;
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + j * 4u), "r"(pv[k]) : "memory");
            }
        }
        tick(6);
        cluster_sync_all();   // the partial results of this step are visible in every CTA (and every peer is done reading those of two steps ago)
        tick(7);
        // acc += sum over the levels (own partial from registers, the others from local shared memory)
        // the peers' words: all loads of a peer are issued before the first sum (a run-time `l != lev` test per word made the
        // compiler wait for every load in turn: 1.0 k of the step's 8 k cycles)
#pragma unroll
        for (uint32_t r = 1; r < (uint32_t)K::L; r++) {
            uint32_t l = lev + r;
            l = l >= (uint32_t)K::L ? l - (uint32_t)K::L : l;
            const uint32_t *src = stepbuf + ((size_t)l * K::P + sub) * K::N + t;
            uint32_t v[2 * K::E];
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) v[k] = src[((uint32_t)(k >> 1) << C::LOGT) + (uint32_t)(k & 1) * K::M];
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) pv[k] += v[k];
        }
#pragma unroll
        for (int k = 0; k < 2 * K::E; k++) {
            const uint32_t j = (((uint32_t)(k >> 1) << C::LOGT) | t) + (uint32_t)(k & 1) * K::M;
            accv[k] += pv[k];
            acc[sub * K::N + j] = accv[k];
        }
        sub_sync();   // acc[sub] (read with a rotation by this sub-team only) is up to date before the next step's digits
        tick(8);
    }
    if (prof) {
        pc[9] = (unsigned long long)(clock64() - tstart);
        for (int k = 0; k < 10; k++) a.prof[k] = pc[k];
    }
    cluster_sync_all();   // no CTA leaves while a peer may still read its partial results
    if (lev == 0) {
        uint32_t *out = a.glwe_out + ((size_t)ct * K::P + sub) * K::N;
        for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[sub * K::N + idx];
    }
    (void)maxfrac;
}


// ---------------------------------------------------------------------------------------------------------------------------
// The same cluster scheme with every CTA's work split by KEY LIMB over twice the warps: 2P sub-teams (column c, limb).  The P
// limb-0 sub-teams decompose and transform the digit rows and publish them; all 2P sub-teams then multiply-accumulate and
// inverse-transform ONE limb each (single-limb inverse through two buffers: two sub-team barriers) -- half the dependent FP64 work
// per warp and two warps per scheduler instead of one, which is what a lone ciphertext lacks (a single team runs at a third of
// its FP64 issue rate, profiles/r02_latency_phases.txt).  The two rounded limbs of a column are combined locally
// (lo + (hi << 16)) before the push, so the cluster exchange is unchanged.
template <class K>
struct ClusterSplitLayout {
    using C = typename K::F;
    static_assert(K::CTS == 1 && K::HALVES == 1 && !K::SINGLE_BUF && K::L <= 8, "cluster kernel configuration");
    static constexpr int SUBS = 2 * K::P, THREADS = SUBS * K::T;
    static constexpr int ACC = 0;                                        // u32 acc[P][N]          (replica)
    static constexpr int PART = ACC + K::P * K::N * 4;                   // u32 part[2][L][P][N]
    static constexpr int BUFS = PART + 2 * K::L * K::P * K::N * 4;       // cplx [2P][2][MPAD]
    static constexpr int SUBBUF_BYTES = 2 * C::MPAD * 16;
    static constexpr int AT = BUFS + SUBS * SUBBUF_BYTES;                // u16 at[n+1]
    static constexpr size_t ring_offset(size_t n) { return ((size_t)AT + (n + 1) * 2 + 127) & ~(size_t)127; }
    // ring slots, their full / empty barriers, the two barriers of the partial-result exchange, the refill claims
    static constexpr size_t smem_bytes(size_t n) { return ring_offset(n) + (size_t)K::NSLOT * K::SLOT_BYTES + 2 * K::NSLOT * 8 + 2 * 8 + 4 * K::NSLOT + 16; }
};
#ifndef TFHE_CLUSTER_ASYNC_PUSH
#define TFHE_CLUSTER_ASYNC_PUSH 1
#endif
// st.async: a remote shared-memory store that reports its bytes to an mbarrier of the DESTINATION CTA.  With it the exchange of the
// partial results needs no cluster-wide rendezvous: every CTA waits on its own barrier until the (L - 1) peers' words have landed.
__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, uint32_t v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return done != 0;
}

template <class K>
__global__ void __launch_bounds__(ClusterSplitLayout<K>::THREADS, 1) pbs_fft_cluster_split_kernel(const __grid_constant__ FftArgs a) {
    using C = typename K::F;
    using LL = ClusterSplitLayout<K>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, sub2 = tid / K::T, t = tid % K::T;
    const uint32_t col = sub2 % (uint32_t)K::P, limb = sub2 / (uint32_t)K::P;   // sub-teams 0..P-1: limb 0 (+ digits and forward transforms)
    const uint32_t lev = cluster_ctarank();
    const uint32_t ct = blockIdx.x / (uint32_t)K::L;
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem + LL::ACC);
    uint32_t *part = reinterpret_cast<uint32_t *>(smem + LL::PART);
    auto subbuf = [&](uint32_t sb_) { return reinterpret_cast<cplx *>(smem + LL::BUFS + sb_ * LL::SUBBUF_BYTES); };
    cplx *buf0 = subbuf(sub2), *buf1 = buf0 + C::MPAD;
    uint16_t *at = reinterpret_cast<uint16_t *>(smem + LL::AT);
    uint8_t *ring = smem + LL::ring_offset(a.n);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + K::NSLOT * K::SLOT_BYTES), *empty = full + K::NSLOT;
    uint64_t *pbar = empty + K::NSLOT;                 // [2] partial results of an executed step have landed (TFHE_CLUSTER_ASYNC_PUSH)
    uint32_t *claimed = reinterpret_cast<uint32_t *>(pbar + 2);
    constexpr bool WARP_SUB = K::T == 32;
    const uint32_t sub_bar = 1 + sub2;
    static_assert(WARP_SUB || LL::SUBS <= 15, "named barrier ids");
    auto sub_sync = [&]() {
        if constexpr (WARP_SUB) __syncwarp();
        else team_bar_id(sub_bar, K::T);
    };
    const uint32_t jbB = jbase_B<C>(t);
    const cplx twB_base = pass_tw_base<C::QB>(a.tw.twB + (t >> C::QB) * C::NB_TW, 1);
    const cplx twC_base = pass_tw_base<C::LOGE>(a.tw.twC + t, C::T);
    const uint32_t total_rows = a.n * (uint32_t)K::P;

    if (tid == 0) {
        for (int s = 0; s < K::NSLOT; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, LL::SUBS * K::WARPS_PER_SUB);   // every warp of the CTA consumes every ring entry
            claimed[s] = 0;
        }
        mbar_init(pbar + 0, 1);
        mbar_init(pbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t *lwe = a.lwe_in + (size_t)ct * (a.n + 1);
    for (uint32_t i = tid; i <= a.n; i += LL::THREADS) at[i] = (uint16_t)mod_switch(__ldg(lwe + i), K::LOGN);
    __syncthreads();
    {
        const uint32_t b = at[a.n];
        uint32_t li = a.lut_idx ? __ldg(a.lut_idx + ct) : 0u;
        if (li >= a.n_luts) {
            atomicOr(a.err_flag, 4u);
            li = 0u;
        }
        const uint32_t *lut = a.luts + (size_t)li * K::N;
        for (uint32_t idx = tid; idx < (uint32_t)(K::P * K::N); idx += LL::THREADS) {
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
    auto issue_row = [&](uint32_t q) {
        const uint32_t s = q % K::NSLOT;
        const size_t row = (size_t)(q / (uint32_t)K::P) * K::ROWS + (size_t)lev * K::P + q % (uint32_t)K::P;
        mbar_expect_tx(full + s, K::SLOT_BYTES);
        bulk_g2s(ring + s * K::SLOT_BYTES, ksrc + row * K::SLOT_BYTES, K::LIMB_BYTES, full + s);
        bulk_g2s(ring + s * K::SLOT_BYTES + K::LIMB_BYTES, ksrc + row * K::SLOT_BYTES + K::LIMB_BYTES, K::LIMB_BYTES, full + s);
    };
    auto release_row = [&](uint32_t q) {
        const uint32_t s = q % K::NSLOT, u = q / K::NSLOT, nx = q + (uint32_t)K::NSLOT;
        mbar_arrive(empty + s);
        if (nx < total_rows && mbar_test(empty + s, u & 1u)) {
            if (atomicCAS(claimed + s, u, u + 1u) == u) issue_row(nx);
        }
    };
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)K::NSLOT && q < total_rows; q++) issue_row(q);

    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    unsigned long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = prof ? clock64() : 0;
    auto tick = [&](int k) {
        if (prof) {
            const long long now = clock64();
            pc[k] += (unsigned long long)(now - tprev);
            tprev = now;
        }
    };
    const long long tstart = tprev;
    cplx x[K::E], ac[K::E];    // transformed digit row (limb-0 sub-teams) / this sub-team's accumulator: column `col`, limb `limb`
    double maxfrac = 0.0;
    uint32_t accv[2 * K::E];   // limb-0 sub-teams: the 2E words of acc[col] this thread decomposes and updates
#pragma unroll
    for (int e = 0; e < K::E; e++) {
        const uint32_t j = ((uint32_t)e << C::LOGT) | t;
        accv[2 * e] = acc[col * K::N + j];
        accv[2 * e + 1] = acc[col * K::N + j + K::M];
    }
    uint32_t q = 0, nexec = 0;
    cluster_sync_all();

#pragma unroll 1
    for (uint32_t i = 0; i < a.n; i++) {
        const uint32_t rot = at[i];
        if (rot == 0) {
#pragma unroll 1
            for (uint32_t d = 0; d < (uint32_t)K::P; d++, q++) {
                mbar_wait(full + (q % K::NSLOT), (q / K::NSLOT) & 1u, a.err_flag);
                __syncwarp();
                if (lane == 0) release_row(q);
            }
            continue;
        }
#pragma unroll
        for (int e = 0; e < K::E; e++) ac[e] = cplx{0.0, 0.0};
        if (limb == 0) {
            // digits of level `lev` of rot(acc) - acc, polynomial `col`, and their forward transform; published in buf0
#pragma unroll
            for (int e = 0; e < K::E; e++) {
                int32_t dl[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t j = (((uint32_t)e << C::LOGT) | t) + (uint32_t)h * K::M;
                    int32_t d[K::L];
                    decompose_signed<K::LOGB, K::L>(rot_coeff(acc + col * K::N, j, rot, K::LOGN) - accv[2 * e + h], d);
                    dl[h] = d[0];
#pragma unroll
                    for (int l = 1; l < K::L; l++) dl[h] = lev == (uint32_t)l ? d[l] : dl[h];
                }
                x[e] = cplx{i2d(dl[0]), i2d(dl[1])};
            }
            fwd_pass<C::LOGE, C::LOGE>(x, a.tw.twA);
            tick(0);
            store_A<C>(x, buf0, t);
            sub_sync();
            {
                cplx tw[C::NB_TW];
                derive_pass_tw<C::QB>(tw, twB_base);
                load_B<C>(x, buf0, jbB);
                fwd_pass<C::LOGE, C::QB>(x, tw);
                store_B<C>(x, buf1, jbB);
            }
            sub_sync();
            {
                cplx tw[C::NC_TW];
                derive_pass_tw<C::LOGE>(tw, twC_base);
                load_C<C>(x, buf1, t);
                fwd_pass<C::LOGE, C::LOGE>(x, tw);
            }
#pragma unroll
            for (int e = 0; e < K::E; e++) buf0[e * K::T + t] = x[e];   // slot order, as phase_xstore
            tick(1);
        }
        __syncthreads();   // all P transformed rows of this level are published
        tick(2);
        // multiply-accumulate of ONE limb: slot d of the level holds at column position `col` the row of polynomial (col + d) mod P
#pragma unroll 1
        for (uint32_t d = 0; d < (uint32_t)K::P; d++, q++) {
            const uint32_t s = q % K::NSLOT;
            mbar_wait(full + s, (q / K::NSLOT) & 1u, a.err_flag);
            const cplx *g = reinterpret_cast<const cplx *>(ring + s * K::SLOT_BYTES) + (limb * (uint32_t)K::P + col) * (uint32_t)K::M + t;
            const cplx *xb = subbuf((col + d) % (uint32_t)K::P);   // the publishing (limb-0) sub-team's buf0
#pragma unroll
            for (int e = 0; e < K::E; e++) {
                const cplx xv = xb[e * K::T + t], gv = g[e * K::T];
                ac[e].re = fma_d(-xv.im, gv.im, fma_d(xv.re, gv.re, ac[e].re));
                ac[e].im = fma_d(xv.im, gv.re, fma_d(xv.re, gv.im, ac[e].im));
            }
            __syncwarp();
            if (lane == 0) release_row(q);
        }
        tick(3);
        __syncthreads();   // the published rows have been read: the buffers are free
        tick(4);
        // single-limb inverse through the sub-team's two buffers
        {
            cplx tw[C::NC_TW];
            derive_pass_tw<C::LOGE>(tw, twC_base);
            inv_pass<C::LOGE, C::LOGE>(ac, tw);
            store_C<C>(ac, buf0, t);
        }
        sub_sync();
        {
            cplx tw[C::NB_TW];
            derive_pass_tw<C::QB>(tw, twB_base);
            load_B<C>(ac, buf0, jbB);
            inv_pass<C::LOGE, C::QB>(ac, tw);
            store_B<C>(ac, buf1, jbB);
        }
        sub_sync();
        load_A<C>(ac, buf1, t);
        inv_pass<C::LOGE, C::LOGE>(ac, a.tw.twA);
        uint32_t pv[2 * K::E];
#pragma unroll
        for (int e = 0; e < K::E; e++) {
            pv[2 * e] = round_u32<false>(ac[e].re, maxfrac);
            pv[2 * e + 1] = round_u32<false>(ac[e].im, maxfrac);
        }
        // the high limb hands its words (already shifted) to the low-limb sub-team of the same column through its own buf0
        uint32_t *hand = reinterpret_cast<uint32_t *>(subbuf(col + (uint32_t)K::P));
        if (limb == 1) {
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) hand[k * K::T + t] = pv[k] << 16;
        }
        tick(5);
        __syncthreads();
        uint32_t *stepbuf = part + (size_t)(nexec & 1u) * K::L * K::P * K::N;
        nexec++;
        if (limb == 0) {
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) pv[k] += hand[k * K::T + t];
            uint32_t *mine = stepbuf + ((size_t)lev * K::P + col) * K::N;
            const uint32_t mine_addr = smem_u32(mine);
#if TFHE_CLUSTER_ASYNC_PUSH
            const uint32_t bar_addr = smem_u32(pbar + ((nexec - 1u) & 1u));
#endif
#pragma unroll
            for (uint32_t r = 1; r < (uint32_t)K::L; r++) {   // push to the peers (same offset in their shared memory)
                uint32_t peer = lev + r;
                peer = peer >= (uint32_t)K::L ? peer - (uint32_t)K::L : peer;
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine_addr), "r"(peer));
#if TFHE_CLUSTER_ASYNC_PUSH
                uint32_t remote_bar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(bar_addr), "r"(peer));
#endif
#pragma unroll
                for (int k = 0; k < 2 * K::E; k++) {
                    const uint32_t j = (((uint32_t)(k >> 1) << C::LOGT) | t) + (uint32_t)(k & 1) * K::M;
#if TFHE_CLUSTER_ASYNC_PUSH
                    st_async_u32(remote + j * 4u, pv[k], remote_bar);
#else
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + j * 4u), "r"(pv[k]) : "memory");
#endif
                }
            }
        }
        tick(6);
#if TFHE_CLUSTER_ASYNC_PUSH
        {
            // this CTA's barrier of the step: armed with the bytes the L - 1 peers send (every word of P polynomials from each), waited
            // for by the threads that accumulate.  The two barriers alternate with the two partial-result buffers; a peer cannot be two
            // executed steps ahead (its next push needs this CTA's words of the step between), so neither a buffer nor a barrier phase
            // is reused early, and no cluster-wide barrier is needed inside the loop.
            uint64_t *bar = pbar + ((nexec - 1u) & 1u);
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)(K::L - 1) * (uint32_t)K::P * (uint32_t)K::N * 4u);
            if (limb == 0) {
                const uint32_t par = ((nexec - 1u) >> 1) & 1u;
                unsigned long long t0 = 0;
                for (uint32_t spin = 1; !mbar_try_cluster(bar, par); spin++) {
                    if ((spin & 255u) == 0u) {
                        if (*(volatile uint32_t *)a.err_flag & 2u) break;
                        const unsigned long long now = global_timer_ns();
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 10000000000ull) { atomicOr(a.err_flag, 2u); break; }
                    }
                }
            }
        }
#else
        cluster_sync_all();
#endif
        tick(7);
        if (limb == 0) {
            // the peers' words: all loads of a peer are issued before the first sum (see pbs_fft_cluster_kernel)
#pragma unroll
            for (uint32_t r = 1; r < (uint32_t)K::L; r++) {
                uint32_t l = lev + r;
                l = l >= (uint32_t)K::L ? l - (uint32_t)K::L : l;
                const uint32_t *src = stepbuf + ((size_t)l * K::P + col) * K::N + t;
                uint32_t v[2 * K::E];
#pragma unroll
                for (int k = 0; k < 2 * K::E; k++) v[k] = src[((uint32_t)(k >> 1) << C::LOGT) + (uint32_t)(k & 1) * K::M];
#pragma unroll
                for (int k = 0; k < 2 * K::E; k++) pv[k] += v[k];
            }
#pragma unroll
            for (int k = 0; k < 2 * K::E; k++) {
                const uint32_t j = (((uint32_t)(k >> 1) << C::LOGT) | t) + (uint32_t)(k & 1) * K::M;
                accv[k] += pv[k];
                acc[col * K::N + j] = accv[k];
            }
            sub_sync();   // acc[col] (read with a rotation by this sub-team only) is up to date before the next step's digits
        }
        tick(8);
    }
    if (prof) {
        pc[9] = (unsigned long long)(clock64() - tstart);
        for (int k = 0; k < 10; k++) a.prof[k] = pc[k];
    }
    cluster_sync_all();
    if (lev == 0 && limb == 0) {
        uint32_t *out = a.glwe_out + ((size_t)ct * K::P + col) * K::N;
        for (uint32_t idx = t; idx < (uint32_t)K::N; idx += K::T) out[idx] = acc[col * K::N + idx];
    }
    (void)maxfrac;
}

}  // namespace fft
}  // namespace tfhe
