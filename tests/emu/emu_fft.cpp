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
    std::vector<FftRegs<K>> regs;
    std::vector<uint32_t> acc;
    std::vector<int16_t> stash;
    std::vector<cplx> buf;
    double maxfrac = 0.0;
    EmuF() : regs(K::T), acc(K::P * K::N), stash((K::L > 1 ? K::L - 1 : 1) * 2 * K::E * K::T), buf(2 * C::MPAD) {
        build_fft_tables(C::LOGM, C::LOGE, tw);
    }
    cplx *b0() { return buf.data(); }
    cplx *b1() { return buf.data() + C::MPAD; }
    const cplx *twB(uint32_t t) const { return tw.B.data() + (t >> C::QB) * C::NB_TW; }
    const cplx *twC(uint32_t t) const { return tw.C.data() + t * C::NC_TW; }

    // raw GGSW [ROWS][P][N] -> [ROWS][2 limbs][P][M] complex, slot order, scaled by 1/M
    void transform_ggsw(const uint32_t *raw, cplx *out) {
        for (int r = 0; r < K::ROWS; r++)
            for (int limb = 0; limb < 2; limb++)
                for (int c = 0; c < K::P; c++) {
                    const uint32_t *g = raw + ((size_t)r * K::P + c) * K::N;
                    cplx *o = out + (((size_t)r * 2 + limb) * K::P + c) * K::M;
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_T1<K>(regs[t], t, limb, g, tw.A.data(), b0());
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2<K>(regs[t], jbase_B<C>(t), twB(t), b0(), b1());
                    for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_T3<K>(regs[t], t, twC(t), b1(), o);
                }
    }
    template <class DiffFn>
    void step(const cplx *key, DiffFn diff) {
        for (uint32_t t = 0; t < (uint32_t)K::T; t++) zero_acc<K>(regs[t]);
        for (int r = 0; r < K::ROWS; r++) {
            const uint32_t p = r / K::L, lev = r % K::L;
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F1<K>(regs[t], t, p, lev, stash.data(), tw.A.data(), b0(), diff);
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_F2<K>(regs[t], jbase_B<C>(t), twB(t), b0(), b1());
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) {
                phase_F3<K>(regs[t], t, twC(t), b1());
                phase_mac<K, 0>(regs[t], t, key + ((size_t)r * 2 + 0) * K::P * K::M);
                phase_mac<K, 1>(regs[t], t, key + ((size_t)r * 2 + 1) * K::P * K::M);
            }
        }
        std::vector<uint32_t> lo((size_t)K::T * 2 * K::E);
        for (int sel = 0; sel < 2 * K::P; sel++) {
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_I1<K>(regs[t], t, sel, twC(t), b0());
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) phase_I2<K>(regs[t], jbase_B<C>(t), twB(t), b0(), b1());
            for (uint32_t t = 0; t < (uint32_t)K::T; t++) {
                phase_I3<K>(regs[t], t, tw.A.data(), b1());
                if ((sel & 1) == 0) phase_round_lo<K>(regs[t], lo.data() + (size_t)t * 2 * K::E, maxfrac);
                else phase_round_hi<K>(regs[t], t, lo.data() + (size_t)t * 2 * K::E, acc.data() + (size_t)(sel >> 1) * K::N, maxfrac);
            }
        }
    }
};

using F_P0 = FftPbsCfg<9, 3, 2, 6, 4, 4>;
using F_P1 = FftPbsCfg<10, 3, 1, 3, 8, 4>;
using F_P2 = FftPbsCfg<11, 4, 1, 3, 8, 4>;

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

}  // namespace

extern "C" {
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
