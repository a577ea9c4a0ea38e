// host_tables_fft.hpp -- host-side twiddle tables of the FP64 FFT path (fft_team.cuh).  Pure C++, shared by the
// product library and tests/emu.  Twiddles are evaluated in long double (64-bit mantissa) and rounded once.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "fft_team.cuh"

namespace tfhe {
namespace fft {

// forward twiddle of stage st (0-based), block b in [0, 2^st):  zeta^((M >> (st+1)) * (1 + 4*brv_st(b))),
// zeta = exp(2 pi i / 4M)   (root tree of X^M - i, see fft_team.cuh)
inline cplx fft_twiddle(int logm, int st, uint32_t b) {
    const uint64_t M = 1ull << logm;
    const uint64_t ex = ((M >> (st + 1)) * (1ull + 4ull * brv_c(b, st))) % (4 * M);
    const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)ex / (long double)(4 * M);
    return cplx{(double)cosl(ang), (double)sinl(ang)};
}

struct HostFftTw {
    std::vector<cplx> A, B, C;  // A: [E-1]; B: [2^LOGE][NB_TW]; C: [E-1][T]
    std::vector<cplx> Z;        // [4M = 2N] zeta^m (BMMP monomial factors)
};
inline void build_fft_tables(int logm, int loge, HostFftTw &out) {
    const int qb = logm - 2 * loge;
    const int T = 1 << (logm - loge);
    out.A.clear(); out.B.clear(); out.C.clear(); out.Z.clear();
    for (uint64_t m = 0; m < (4ull << logm); m++) {
        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)(4ull << logm);
        out.Z.push_back(m == 0 ? cplx{1.0, 0.0} : cplx{(double)cosl(ang), (double)sinl(ang)});
    }
    for (int u = 0; u < loge; u++)
        for (uint32_t m = 0; m < (1u << u); m++) out.A.push_back(fft_twiddle(logm, u, m));
    for (uint32_t hA = 0; hA < (1u << loge); hA++)
        for (int u = 0; u < qb; u++)
            for (uint32_t m = 0; m < (1u << u); m++) out.B.push_back(fft_twiddle(logm, loge + u, (hA << u) | m));
    for (int u = 0; u < loge; u++)   // layout [NC_TW][T]
        for (uint32_t m = 0; m < (1u << u); m++)
            for (uint32_t t = 0; t < (uint32_t)T; t++) out.C.push_back(fft_twiddle(logm, loge + qb + u, (t << u) | m));
}

// per-lane entries of the tensor-memory-exchange passes (fft_tmem.cuh TmemTw), M = 256: the twiddle of the LAST stage of each
// pass for the block whose register-borne bits are zero
inline void build_fft_tmem_table(std::vector<cplx> &out) {
    out.clear();
    for (uint32_t i = 0; i < 4; i++) out.push_back(fft_twiddle(8, 4, i << 1));     // w(4, (0 j6 j5 0)),         lane & 3  = (j6 j5)
    for (uint32_t i = 0; i < 16; i++) out.push_back(fft_twiddle(8, 6, i << 1));    // w(6, (0 j6 j5 j4 j3 0)),   lane & 15 = (j6 j5 j4 j3)
    for (uint32_t i = 0; i < 32; i++) out.push_back(fft_twiddle(8, 7, i));         // w(7, (0 0 j5 j4 j3 j2 j1)), lane      = (j5 j4 j3 j2 j1)
}

// per-thread entries of the tensor-memory tail of the M = 512 transform (fft_tmem.cuh TailTw)
inline void build_fft_tail_table(std::vector<cplx> &out) {
    out.clear();
    for (uint32_t i = 0; i < 32; i++) {   // i = (j8 j7 j6 j4 j3): w(7, (j8 j7 j6 0 j4 j3 0))
        const uint32_t b = ((i >> 2) << 4) | ((i & 3u) << 1);
        out.push_back(fft_twiddle(9, 7, b));
    }
    for (uint32_t i = 0; i < 64; i++) {   // i = (j8 j7 j4 j3 j2 j1): w(8, (j8 j7 0 0 j4 j3 j2 j1))
        const uint32_t b = ((i >> 4) << 6) | (i & 15u);
        out.push_back(fft_twiddle(9, 8, b));
    }
}

// per-thread entries of the M = 1024 transform with a tensor-memory tail (fft_tmem.cuh Tail16Tw)
inline void build_fft_tail16_table(std::vector<cplx> &out) {
    out.clear();
    for (uint32_t hA = 0; hA < 16; hA++) out.push_back(fft_twiddle(10, 7, hA << 3));   // pass B'': stages 4..7, block hA
    for (uint32_t i = 0; i < 64; i++) {   // i = (j9 j8 j7 j6 j3 j2): w(9, (j9 j8 j7 j6 0 0 j3 j2 0))
        const uint32_t b = ((i >> 2) << 5) | ((i & 3u) << 1);
        out.push_back(fft_twiddle(10, 9, b));
    }
}

}  // namespace fft
}  // namespace tfhe
