// tests/emu/emu.cpp -- lock-step CPU emulation of ONE CTA of the blind-rotation kernel.
// TEST INFRASTRUCTURE: it steps the product's own __host__ __device__ phase functions
// (tfhe-research_b200/csrc/pbs_team.cuh) for every thread of a CTA, phase by phase, exactly where
// the CUDA kernel has its barriers.  It lets the CPU-only test suite check indexing, twiddle tables,
// bounds (TFHE_EMU_CHECKS) and bit-exactness against the oracle without a GPU.  Not shipped.
#define TFHE_EMU_CHECKS 1
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../tfhe-research_b200/csrc/host_tables.hpp"
#include "../../tfhe-research_b200/csrc/pbs_team.cuh"

using namespace tfhe;

namespace {

template <class K>
struct Emu {
    using C = typename K::Ntt;
    HostTw tw;
    PrimeTab prime[2];
    std::vector<TeamRegs<K>> regs;          // [2 primes][T]
    std::vector<uint32_t> acc, res, buf;    // shared memory images
    std::vector<uint8_t> dig;
    Emu() : regs(2 * K::T), acc(K::P * K::N), res(2 * K::P * K::N), buf(2 * 2 * C::NPAD), dig(K::ROWS * K::N * K::DIG_BYTES) {
        build_tw_tables(C::LOGN, C::LOGE, tw);
        fill_prime_tab(0, C::LOGN, C::LOGE, prime[0]);
        fill_prime_tab(1, C::LOGN, C::LOGE, prime[1]);
        for (int pr = 0; pr < 2; pr++)
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) team_init<K>(regs[pr * K::T + t], tables(pr), t);
    }
    TwTables tables(int pr) const {
        TwTables t;
        t.fwdB = reinterpret_cast<const uint2 *>(tw.fwdB[pr].data());
        t.fwdC = reinterpret_cast<const uint2 *>(tw.fwdC[pr].data());
        t.invB = reinterpret_cast<const uint2 *>(tw.invB[pr].data());
        t.invC = reinterpret_cast<const uint2 *>(tw.invC[pr].data());
        return t;
    }
    uint32_t *bufp(int pr, int which) { return buf.data() + (pr * 2 + which) * C::NPAD; }

    template <int PR>
    void transform_poly(const uint32_t *g, uint32_t *out) {
        TwTables t = tables(PR);
        const PrimeTab &pt = prime[PR];
        for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_T1<K>(regs[PR * K::T + th], th, jbase_B<C>(th), pt, t, g, bufp(PR, 0));
        for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_F2<K>(regs[PR * K::T + th], jbase_B<C>(th), pt, bufp(PR, 0), bufp(PR, 1));
        for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_T3<K>(regs[PR * K::T + th], th, pt, t, bufp(PR, 1), out);
    }
    // raw GGSW [ROWS][P][N] -> [2][ROWS][P][N] (NTT domain, slot order)
    void transform_ggsw(const uint32_t *raw, uint32_t *ntt) {
        for (int r = 0; r < K::ROWS; r++)
            for (int c = 0; c < K::P; c++) {
                const uint32_t *g = raw + ((size_t)r * K::P + c) * K::N;
                transform_poly<0>(g, ntt + ((size_t)(0 * K::ROWS + r) * K::P + c) * K::N);
                transform_poly<1>(g, ntt + ((size_t)(1 * K::ROWS + r) * K::P + c) * K::N);
            }
    }
    template <int PR>
    void team_phase(const uint32_t *ggsw_ntt) {
        TwTables t = tables(PR);
        TeamRegs<K> *R = &regs[PR * K::T];
        const PrimeTab &pt = prime[PR];
        for (uint32_t th = 0; th < (uint32_t)K::T; th++) team_zero_acc<K>(R[th]);
        for (int r = 0; r < K::ROWS; r++) {
            const uint32_t *g_row = ggsw_ntt + (size_t)(PR * K::ROWS + r) * K::P * K::N;
            for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_F1<K>(R[th], th, jbase_B<C>(th), pt, t, dig.data(), r, bufp(PR, 0));
            for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_F2<K>(R[th], jbase_B<C>(th), pt, bufp(PR, 0), bufp(PR, 1));
            for (uint32_t th = 0; th < (uint32_t)K::T; th++) {
                phase_F3a<K>(R[th], th, pt, t, bufp(PR, 1));
                phase_F3b<K, true>(R[th], th, g_row);
            }
        }
        for (int c = 0; c < K::P; c++) {
            for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_I1<K>(R[th], th, jbase_B<C>(th), c, pt, t, bufp(PR, 0));
            for (uint32_t th = 0; th < (uint32_t)K::T; th++) phase_I2<K>(R[th], jbase_B<C>(th), pt, bufp(PR, 0), bufp(PR, 1));
            for (uint32_t th = 0; th < (uint32_t)K::T; th++)
                phase_I3<K>(R[th], th, pt, bufp(PR, 1), res.data() + (size_t)(PR * K::P + c) * K::N);
        }
    }
    // acc <- ExtProd(ggsw, diff) + acc, diff given by functor (reads a snapshot of acc)
    template <class DiffFn>
    void step(const uint32_t *ggsw_ntt, DiffFn diff) {
        for (uint32_t tid = 0; tid < (uint32_t)K::THREADS; tid++) phase_digits<K>(tid, dig.data(), diff);
        team_phase<0>(ggsw_ntt);
        team_phase<1>(ggsw_ntt);
        for (uint32_t tid = 0; tid < (uint32_t)K::THREADS; tid++) phase_crt<K>(tid, res.data(), acc.data());
    }
};

using K_P0 = PbsCfg<9, 3, 2, 6, 4>;
using K_P1 = PbsCfg<10, 4, 1, 3, 8>;
using K_P2 = PbsCfg<11, 4, 1, 3, 8>;

template <class K>
int run_transform(const uint32_t *raw, uint32_t *ntt) {
    Emu<K> e;
    e.transform_ggsw(raw, ntt);
    return 0;
}
// mode 0: blind-rotate step  acc <- cmux(ggsw, acc, acc * X^a)   (bootstrapping.rs:94-104)
// mode 1: external product   out <- ExtProd(ggsw, glwe)           (ggsw.rs:132-161)
template <class K>
int run_step(int mode, const uint32_t *ggsw_ntt, uint32_t *glwe, uint32_t a) {
    Emu<K> e;
    if (mode == 0) {
        memcpy(e.acc.data(), glwe, sizeof(uint32_t) * K::P * K::N);
        const uint32_t *acc = e.acc.data();
        e.step(ggsw_ntt, [=](uint32_t p, uint32_t j) { return rot_coeff(acc + p * K::N, j, a, K::LOGN) - acc[p * K::N + j]; });
    } else {
        std::vector<uint32_t> in(glwe, glwe + K::P * K::N);
        const uint32_t *src = in.data();
        e.step(ggsw_ntt, [=](uint32_t p, uint32_t j) { return src[p * K::N + j]; });
    }
    memcpy(glwe, e.acc.data(), sizeof(uint32_t) * K::P * K::N);
    return 0;
}

}  // namespace

extern "C" {
int emu_transform_ggsw(int cfg, const uint32_t *raw, uint32_t *ntt) {
    switch (cfg) {
    case 0: return run_transform<K_P0>(raw, ntt);
    case 1: return run_transform<K_P1>(raw, ntt);
    case 2: return run_transform<K_P2>(raw, ntt);
    }
    return -1;
}
int emu_step(int cfg, int mode, const uint32_t *ggsw_ntt, uint32_t *glwe, uint32_t a) {
    switch (cfg) {
    case 0: return run_step<K_P0>(mode, ggsw_ntt, glwe, a);
    case 1: return run_step<K_P1>(mode, ggsw_ntt, glwe, a);
    case 2: return run_step<K_P2>(mode, ggsw_ntt, glwe, a);
    }
    return -1;
}
uint32_t emu_mod_switch(uint32_t v, int logn) { return mod_switch(v, logn); }
void emu_decompose(uint32_t v, int log_base, int levels, int32_t *out) {
    if (log_base == 4 && levels == 6) decompose_signed<4, 6>(v, out);
    else if (log_base == 4 && levels == 5) decompose_signed<4, 5>(v, out);
    else if (log_base == 8 && levels == 3) decompose_signed<8, 3>(v, out);
    else if (log_base == 2 && levels == 8) decompose_signed<2, 8>(v, out);
    else if (log_base == 8 && levels == 4) decompose_signed<8, 4>(v, out);
    else if (log_base == 4 && levels == 8) decompose_signed<4, 8>(v, out);
}
uint32_t emu_crt(uint32_t r0, uint32_t r1) { return crt_to_u32(r0, r1); }
}
