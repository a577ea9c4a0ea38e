// tests/emu/emu_fft.cpp -- lock-step CPU emulation of ONE team of the FP64-FFT blind-rotation kernel.
// TEST INFRASTRUCTURE: steps the product's own __host__ __device__ phase functions (fft_team.cuh) for every
// thread of a team, phase by phase, exactly where the CUDA kernel has its barriers, so the CPU-only suite can pin
// the fold, twiddle tables, limb split, layouts and rounding against the oracle bit for bit.  Not shipped.
// Build with -ffp-contract=off: the phase functions pin every FP64 operation, the emulation reproduces them.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../tfhe-research_b200/csrc/host_tables_fft.hpp"

using namespace tfhe::fft;
using tfhe::rot_coeff;

namespace {

template <class K>
struct EmuF {
    using C = typename K::F;
    HostFftTw tw;
    std::vector<FftRegs<K>> regs;            // [P sub-teams][T]
    std::vector<uint32_t> acc;
    std::vector<typename K::stash_t> stash;  // [P][(L-1)*2E*T]
    std::vector<cplx> buf;                   // [P][2][MPAD]
    double maxfrac = 0.0;
    static constexpr size_t STASH = (K::L > 1 ? K::L - 1 : 1) * 2 * K::E * K::T;
    EmuF() : regs(K::P * K::T), acc(K::P * K::N), stash(K::P * STASH), buf(K::P * 2 * C::MPAD) {
        build_fft_tables(C::LOGM, C::LOGE, tw);
    }
    FftRegs<K> &R(int s, uint32_t t) { return regs[s * K::T + t]; }
    cplx *b0(int s) { return buf.data() + (size_t)s * 2 * C::MPAD; }
    cplx *b1(int s) { return b0(s) + C::MPAD; }
    const cplx *twB(uint32_t t) const { return tw.B.data() + (t >> C::QB) * C::NB_TW; }
    const cplx *twC() const { return tw.C.data(); }

    // raw GGSW [ROWS][P][N] -> [ROWS (level-major)][2 limbs][P][M] complex, slot order, scaled by 1/M
    void transform_ggsw(const uint32_t *raw, cplx *out) {
        for (int r = 0; r < K::ROWS; r++)
            for (int limb = 0; limb < 2; limb++)
                for (int c = 0; c < K::P; c++) {
                    const uint32_t *g = raw + ((size_t)r * K::P + c) * K::N;
                    const size_t row = key_row_index<K>(r / K::L, r % K::L);
                    cplx *o = out + ((row * K::HALVES * 2 + limb) * K::P + c) * K::MH;   // half 0; phase_T3 adds the half stride
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_T1<K>(R(0, t), t, limb, g, tw.A.data(), b0(0));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2<K>(R(0, t), jbase_B<C>(t), twB(t), b0(0), b1(0));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_T3<K>(R(0, t), t, twC(), b1(0), o);
                }
    }
    // BMMP step (notes/BMMP Bootstrapping.md): key = 3 GGSWs [ROWS][3][2][P][M], exponents ex[3]; acc += ExtProd(bundle, acc)
    void step_bmmp(const cplx *key, const uint32_t *ex) {
        static_assert(K::HALVES == 1, "whole-row key slots only");
        std::vector<uint32_t> snap(acc);
        const uint32_t *a = snap.data();
        run(key, [=](uint32_t p, uint32_t j) { return a[p * K::N + j]; }, 3, ex);
    }
    template <class DiffFn>
    void step(const cplx *key, DiffFn diff) { run(key, diff, 1, nullptr); }
    template <class DiffFn>
    void run(const cplx *key, DiffFn diff, int keys, const uint32_t *ex) {
        for (auto &r : regs) zero_acc<K>(r);
        for (int lev = 0; lev < K::L; lev++) {
            for (int s = 0; s < K::P; s++) {   // sub-teams run concurrently on the GPU; any order between barriers is legal
                for (uint32_t t = 0; t < (uint32_t)K::T; t++)
                    phase_F1<K>(R(s, t), t, s, lev, stash.data() + s * STASH, tw.A.data(), b0(s), diff);
                if constexpr (K::SINGLE_BUF) {   // the kernel's single-buffer flow: every loop below is one barrier interval
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2a<K>(R(s, t), jbase_B<C>(t), twB(t), b0(s));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2b<K>(R(s, t), jbase_B<C>(t), b0(s));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F3<K>(R(s, t), t, twC(), b0(s));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_xstore<K>(R(s, t), t, b0(s));
                } else {
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2<K>(R(s, t), jbase_B<C>(t), twB(t), b0(s), b1(s));
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) {
                        phase_F3<K>(R(s, t), t, twC(), b1(s));
                        phase_xstore<K>(R(s, t), t, b0(s));
                    }
                }
            }
            for (int p = 0; p < K::P; p++)
                for (int which = 0; which < keys; which++)
                    for (int h = 0; h < K::HALVES; h++) {
                        const cplx *slot = key + (((size_t)key_row_index<K>(p, lev) * keys + which) * K::HALVES + h) * 2 * K::P * K::MH;
                        for (int s = 0; s < K::P; s++)
                            for (uint32_t t = 0; t < (uint32_t)K::T; t++) {
                                if constexpr (K::HALVES == 1) {
                                    if (keys == 3) {
                                        if (p == s) phase_mac_bmmp<K, true>(R(s, t), t, s, slot, nullptr, tw.Z.data(), ex[which], bmmp_base<K>(tw.Z.data(), t, ex[which]));
                                        else phase_mac_bmmp<K, false>(R(s, t), t, s, slot, b0(p), tw.Z.data(), ex[which], bmmp_base<K>(tw.Z.data(), t, ex[which]));
                                        continue;
                                    }
                                }
                                if (p == s) phase_mac<K, true>(R(s, t), t, s, slot, nullptr, h);
                                else phase_mac<K, false>(R(s, t), t, s, slot, b0(p), h);
                            }
                    }
        }
        for (int s = 0; s < K::P; s++) {
            if constexpr (K::SINGLE_BUF) {
                std::vector<uint32_t> lo((size_t)K::T * 2 * K::E);
                uint32_t *ac = acc.data() + (size_t)s * K::N;
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K1<K, 0>(R(s, t), t, twC(), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K2a<K, 0>(R(s, t), jbase_B<C>(t), twB(t), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K2b<K, 0>(R(s, t), jbase_B<C>(t), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K3_lo<K>(R(s, t), t, tw.A.data(), b0(s), lo.data() + (size_t)t * 2 * K::E, maxfrac);
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K1<K, 1>(R(s, t), t, twC(), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K2a<K, 1>(R(s, t), jbase_B<C>(t), twB(t), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K2b<K, 1>(R(s, t), jbase_B<C>(t), b0(s));
                for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_K3_hi<K>(R(s, t), t, tw.A.data(), b0(s), lo.data() + (size_t)t * 2 * K::E, ac, maxfrac);
                continue;
            }
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_J1<K>(R(s, t), t, twC(), b0(s), b1(s));
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_J2a<K>(R(s, t), jbase_B<C>(t), twB(t), b0(s), b1(s));
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_J2b<K>(R(s, t), jbase_B<C>(t), b0(s), b1(s));
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_J3<K>(R(s, t), t, tw.A.data(), b0(s), b1(s), acc.data() + (size_t)s * K::N, maxfrac);
        }
    }
};

using F_P0 = FftPbsCfg<9, 3, 2, 6, 4, 4>;
using F_P1 = FftPbsCfg<10, 3, 1, 3, 8, 3>;
using F_P2 = FftPbsCfg<11, 4, 1, 3, 8, 2, true, true, 2>;   // single exchange buffer, half-row key slots

template <class K>
int run_transform(const uint32_t *raw, double *out) {
    EmuF<K> e;
    e.transform_ggsw(raw, reinterpret_cast<cplx *>(out));
    return 0;
}
template <class K>
int run_step(int mode, const double *key, uint32_t *glwe, uint32_t a, double *maxfrac) {
    EmuF<K> e;
    const cplx *k = reinterpret_cast<const cplx *>(key);
    if (mode == 0) {  // blind-rotate step: acc <- cmux(ggsw, acc, acc * X^a)   (bootstrapping.rs:94-104)
        memcpy(e.acc.data(), glwe, sizeof(uint32_t) * K::P * K::N);
        std::vector<uint32_t> snap(e.acc);
        const uint32_t *acc = snap.data();
        e.step(k, [=](uint32_t p, uint32_t j) { return rot_coeff(acc + p * K::N, j, a, K::LOGN) - acc[p * K::N + j]; });
    } else {          // external product: out <- ExtProd(ggsw, glwe)           (ggsw.rs:132-161)
        std::vector<uint32_t> in(glwe, glwe + K::P * K::N);
        const uint32_t *src = in.data();
        e.step(k, [=](uint32_t p, uint32_t j) { return src[p * K::N + j]; });
    }
    memcpy(glwe, e.acc.data(), sizeof(uint32_t) * K::P * K::N);
    *maxfrac = e.maxfrac;
    return 0;
}

// three raw GGSWs -> [ROWS][3][2][P][M]; then one BMMP step on glwe with exponents (a + a', a, a')
template <class K>
int run_bmmp(const uint32_t *raw3, uint32_t *glwe, uint32_t a0, uint32_t a1, double *maxfrac) {
    EmuF<K> e;
    const size_t per = (size_t)K::ROWS * 2 * K::P * K::M, gg = (size_t)K::ROWS * K::P * K::N;
    std::vector<cplx> one(per), key(3 * per);
    for (int which = 0; which < 3; which++) {
        e.transform_ggsw(raw3 + which * gg, one.data());
        for (int row = 0; row < K::ROWS; row++)
            memcpy(&key[((size_t)row * 3 + which) * 2 * K::P * K::M], &one[(size_t)row * 2 * K::P * K::M], sizeof(cplx) * 2 * K::P * K::M);
    }
    memcpy(e.acc.data(), glwe, sizeof(uint32_t) * K::P * K::N);
    const uint32_t ex[3] = {(a0 + a1) & (2u * K::N - 1u), a0, a1};
    e.step_bmmp(key.data(), ex);
    memcpy(glwe, e.acc.data(), sizeof(uint32_t) * K::P * K::N);
    *maxfrac = e.maxfrac;
    return 0;
}

}  // namespace

extern "C" {
int emu_fft_step_bmmp(int cfg, const uint32_t *raw3, uint32_t *glwe, uint32_t a0, uint32_t a1, double *maxfrac) {
    switch (cfg) {
    case 0: return run_bmmp<F_P0>(raw3, glwe, a0, a1, maxfrac);
    case 1: return run_bmmp<F_P1>(raw3, glwe, a0, a1, maxfrac);
    }
    return -1;   // the BMMP variant is instantiated for whole-row key slots only (P0, P1)
}
int emu_fft_transform_ggsw(int cfg, const uint32_t *raw, double *out) {
    switch (cfg) {
    case 0: return run_transform<F_P0>(raw, out);
    case 1: return run_transform<F_P1>(raw, out);
    case 2: return run_transform<F_P2>(raw, out);
    }
    return -1;
}
int emu_fft_step(int cfg, int mode, const double *key, uint32_t *glwe, uint32_t a, double *maxfrac) {
    switch (cfg) {
    case 0: return run_step<F_P0>(mode, key, glwe, a, maxfrac);
    case 1: return run_step<F_P1>(mode, key, glwe, a, maxfrac);
    case 2: return run_step<F_P2>(mode, key, glwe, a, maxfrac);
    }
    return -1;
}
}
