// fft_tmem.cuh -- the register passes of the folded FFT exchanged through TENSOR MEMORY instead of shared memory.
//   M = 256  (N = 512, the reference's default set; sub-team = one warp): every exchange, and the rows the sub-teams publish to each other
//   M = 512  (N = 1024; sub-team = two warps): the last three stages ("tail9"), after pass A, one shared-memory exchange and pass B
//   M = 1024 (N = 2048; 16 points per thread): the last two stages ("tail16"), after pass A, one shared-memory exchange and a 4-stage pass B''
// plus per-thread derived twiddles parked in spare tensor-memory columns (they never change, and do not fit in registers).
// The first part of this file describes the primitive and the M = 256 transform; the tails are at the end.
//
// Why: the blind rotation is bound by the shared-memory data pipe (128 B/clk/SM), and a third of its bytes are the exchanges
// between the register passes of the transforms.  Tensor memory has its own data path: a warp that stores its registers with
// one tcgen05.st shape and loads them back with another gets them TRANSPOSED between lanes, and that traffic runs beside the
// shared-memory pipe, not on it (tools/tmem_xchg.cu, profiles/r02_tmem_xchg.txt: an exchange of 8 complex doubles per thread
// costs 777 cycles per 12 warps through shared memory, 276 through tensor memory, 794 for BOTH at once).
//
// The primitive (measured lane for lane by tools/tmem_xchg.cu): tcgen05.st.32x32b -- thread = TMEM lane L, its words = columns --
// followed by tcgen05.ld.16x256b at lane offsets 0 and 16 -- thread t receives lanes t/4 + {0, 8, 16, 24}, columns 2 (t%4) + {0, 1} of
// every group of 8 columns.  With the 8 complex values of a thread laid out as column 8 (2 part + r2) + 2 (r1 r0) + word
// (part = re / im, r = register index (r2 r1 r0)) this is the index-bit permutation
//       writer lane (L4 L3 L2 L1 L0), register (r2 r1 r0)   ->   reader lane (L2 L1 L0 r1 r0), register (L4 L3 r2):
// two register bits and two lane bits swap places.  The reverse (st.16x256b, ld.32x32b) undoes it.
//
// The transform (same butterfly network and twiddles as fft_team.cuh: stage s pairs index bit LOGM-1-s with w(s, j >> (LOGM - s))):
//   layout A   lane (j4 j3 j2 j1 j0)  regs (j7 j6 j5)   stages 0 1 2  (pass A of fft_team.cuh, twiddles from the constant bank)
//   swap  ->   lane (j2 j1 j0 j6 j5)  regs (j4 j3 j7)   stages 3 4
//   swap  ->   lane (j0 j6 j5 j4 j3)  regs (j2 j1 j7)   stages 5 6
//   swap  ->   lane (j5 j4 j3 j2 j1)  regs (j0 j6 j7)   stage  7      = layout F: the spectral layout of this path
// (a register renaming between the passes puts the two bits just processed into the (r1 r0) position: free).  The inverse runs
// the mirror image.  Twiddles: one table entry per thread and pass (the block with all register-borne bits zero); the entries for
// the register-borne bits follow by squaring (w(s, b) = w(s+1, 2b)^2), by i (the bit processed in the previous stage of the pass)
// and by fixed roots of unity (bit j7 / j6, which ride along in the registers) -- the error terms are those of derive_pass_tw
// (at most one squaring and two constant multiplications per entry).
// Device-only code (tcgen05): not part of the CPU emulation; pinned by the GPU parity tests on both arithmetic paths.
#pragma once
#include "fft_team.cuh"

// measured choices (tools/gpu_call8.sh, P0 batch 4096): the swap stores as 4 x tcgen05.st.x8 (fewer register moves than one x32:
// 102.95 -> 99.49 ms), 16x256b loads / stores as .x4 (104.26 -> 102.18 ms), a peer's row requested before the wait
// for the key slot; publishing / fetching a row as 4 x .x8 instead of one .x32 was slower (100.3 / 100.0 vs 99.5 ms)
// (the same two choices hold for the tails: P1 61.3 vs 62.9 (x32 store) / 61.8 (x1 shapes) ms, P2 94.9 vs 99.1 / 96.2 ms)
#ifndef TFHE_TMEM_ST8
#define TFHE_TMEM_ST8 1
#endif
#ifndef TFHE_TMEM_LOADFIRST
#define TFHE_TMEM_LOADFIRST 1
#endif
#ifndef TFHE_TMEM_X4
#define TFHE_TMEM_X4 1
#endif

namespace tfhe {
namespace fft {

// spectral position of index j in layout F: lane (j5 j4 j3 j2 j1), register (j0 j6 j7) -> r * 32 + lane  (slot order of this path)
__host__ __device__ constexpr uint32_t tmem_slot_of_index(uint32_t j) {
    const uint32_t lane = (j >> 1) & 31u, r = ((j & 1u) << 2) | (((j >> 6) & 1u) << 1) | ((j >> 7) & 1u);
    return r * 32u + lane;
}

// the same for layout F9 (M = 512): thread warp j8, lane (j7 j4 j3 j2 j1), register (j0 j6 j5) -> r * 64 + thread
__host__ __device__ constexpr uint32_t tail9_slot_of_index(uint32_t j) {
    const uint32_t th = (((j >> 8) & 1u) << 5) | (((j >> 7) & 1u) << 4) | ((j >> 1) & 15u), r = ((j & 1u) << 2) | (((j >> 6) & 1u) << 1) | ((j >> 5) & 1u);
    return r * 64u + th;
}

// the same for layout F10 (M = 1024, 16 points per thread): thread warp j9, lane (j8 j7 j6 j3 j2), register (j5 j1 j0 j4)
__host__ __device__ constexpr uint32_t tail10_thread_of_index(uint32_t j) { return (((j >> 9) & 1u) << 5) | (((j >> 6) & 7u) << 2) | ((j >> 2) & 3u); }
__host__ __device__ constexpr uint32_t tail10_reg_of_index(uint32_t j) { return (((j >> 5) & 1u) << 3) | (((j >> 1) & 1u) << 2) | ((j & 1u) << 1) | ((j >> 4) & 1u); }

__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// tcgen05.ld is asynchronous: its destination registers may only be read after tcgen05.wait::ld.  The words are passed through the
// wait as in/out operands so that the compiler cannot schedule a use of them above it.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&w)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]),
                   "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(w[18]), "+r"(w[19]), "+r"(w[20]), "+r"(w[21]), "+r"(w[22]),
                   "+r"(w[23]), "+r"(w[24]), "+r"(w[25]), "+r"(w[26]), "+r"(w[27]), "+r"(w[28]), "+r"(w[29]), "+r"(w[30]), "+r"(w[31])
                 :
                 : "memory");
}

// stores of the swap: all 32 words of the thread, column 8 (2 part + r2) + 2 (r1 r0) + word
__device__ __forceinline__ void tmem_swap2_store(const cplx (&x)[8], uint32_t taddr) {
    uint32_t v[32];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int c = 8 * (r >> 2) + 2 * (r & 3);
        v[c] = (uint32_t)__double2loint(x[r].re);
        v[c + 1] = (uint32_t)__double2hiint(x[r].re);
        v[16 + c] = (uint32_t)__double2loint(x[r].im);
        v[16 + c + 1] = (uint32_t)__double2hiint(x[r].im);
    }
#if TFHE_TMEM_ST8
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr + 8u * q), "r"(v[8 * q]), "r"(v[8 * q + 1]), "r"(v[8 * q + 2]),
                     "r"(v[8 * q + 3]), "r"(v[8 * q + 4]), "r"(v[8 * q + 5]), "r"(v[8 * q + 6]), "r"(v[8 * q + 7])
                     : "memory");
#else
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
#endif
}
// loads of the swap: new register (L4 L3 r2) <- lane half h = L4, row + 8 = L3, column group parity = r2.
// One 16x256b.x4 per lane half: repetition g reads column group g (8 columns), 4 words each (rows t/4 and t/4 + 8).
__device__ __forceinline__ void tmem_swap2_load(cplx (&x)[8], uint32_t taddr) {
    uint32_t w[32];
#pragma unroll
    for (int h = 0; h < 2; h++) {
#if TFHE_TMEM_X4
        const int o = 16 * h;
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(w[o]), "=r"(w[o + 1]), "=r"(w[o + 2]), "=r"(w[o + 3]), "=r"(w[o + 4]), "=r"(w[o + 5]), "=r"(w[o + 6]), "=r"(w[o + 7]), "=r"(w[o + 8]),
                       "=r"(w[o + 9]), "=r"(w[o + 10]), "=r"(w[o + 11]), "=r"(w[o + 12]), "=r"(w[o + 13]), "=r"(w[o + 14]), "=r"(w[o + 15])
                     : "r"(taddr + ((uint32_t)(16 * h) << 16))
                     : "memory");
#else
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int o = 4 * (4 * h + g);
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w[o]), "=r"(w[o + 1]), "=r"(w[o + 2]), "=r"(w[o + 3])
                         : "r"(taddr + ((uint32_t)(16 * h) << 16) + (uint32_t)(8 * g))
                         : "memory");
        }
#endif
    }
    tmem_wait_ld(w);
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int o = 4 * (4 * h + g), r0 = (h << 2) | (g & 1), r1 = r0 | 2;
            if (g < 2) {
                x[r0].re = __hiloint2double((int)w[o + 1], (int)w[o]);
                x[r1].re = __hiloint2double((int)w[o + 3], (int)w[o + 2]);
            } else {
                x[r0].im = __hiloint2double((int)w[o + 1], (int)w[o]);
                x[r1].im = __hiloint2double((int)w[o + 3], (int)w[o + 2]);
            }
        }
}
// the reverse direction: registers (c2 c1 c0) = (L4 L3 r2) go back to lane (L4 L3 ...), register (r2 r1 r0)
__device__ __forceinline__ void tmem_unswap2_store(const cplx (&x)[8], uint32_t taddr) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t v[16];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int r0 = (h << 2) | (g & 1), r1 = r0 | 2;
            const double d0 = g < 2 ? x[r0].re : x[r0].im, d1 = g < 2 ? x[r1].re : x[r1].im;
            v[4 * g] = (uint32_t)__double2loint(d0);
            v[4 * g + 1] = (uint32_t)__double2hiint(d0);
            v[4 * g + 2] = (uint32_t)__double2loint(d1);
            v[4 * g + 3] = (uint32_t)__double2hiint(d1);
        }
#if TFHE_TMEM_X4
        asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr + ((uint32_t)(16 * h) << 16)),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
                     "r"(v[13]), "r"(v[14]), "r"(v[15])
                     : "memory");
#else
#pragma unroll
        for (int g = 0; g < 4; g++)
            asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr + ((uint32_t)(16 * h) << 16) + (uint32_t)(8 * g)), "r"(v[4 * g]),
                         "r"(v[4 * g + 1]), "r"(v[4 * g + 2]), "r"(v[4 * g + 3])
                         : "memory");
#endif
    }
}
__device__ __forceinline__ void tmem_unswap2_load(cplx (&x)[8], uint32_t taddr) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
          "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    tmem_wait_ld(v);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int c = 8 * (r >> 2) + 2 * (r & 3);
        x[r].re = __hiloint2double((int)v[c + 1], (int)v[c]);
        x[r].im = __hiloint2double((int)v[16 + c + 1], (int)v[16 + c]);
    }
}
// whole swaps (store, wait, load, wait)
__device__ __forceinline__ void tmem_swap2(cplx (&x)[8], uint32_t taddr) {
    tmem_swap2_store(x, taddr);
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
}
__device__ __forceinline__ void tmem_unswap2(cplx (&x)[8], uint32_t taddr) {
    tmem_unswap2_store(x, taddr);
    tmem_wait_st();
    tmem_unswap2_load(x, taddr);
}

// fixed roots of unity exp(2 pi i / 2^k)
template <int K> struct RootOfUnity;
template <> struct RootOfUnity<4> { static constexpr double c = 0.92387953251128675613, s = 0.38268343236508977173; };   // 2 pi / 16
template <> struct RootOfUnity<5> { static constexpr double c = 0.98078528040323044913, s = 0.19509032201612826785; };   // 2 pi / 32
template <> struct RootOfUnity<6> { static constexpr double c = 0.99518472667219688624, s = 0.09801714032956060199; };   // 2 pi / 64
template <> struct RootOfUnity<7> { static constexpr double c = 0.99879545620517239271, s = 0.04906767432741801426; };   // 2 pi / 128
template <> struct RootOfUnity<8> { static constexpr double c = 0.99969881869620422012, s = 0.02454122852291228803; };   // 2 pi / 256

// twiddles of a two-stage pass on registers (b_hi b_lo j7): stage s on b_hi with w(s, (j7 | lane bits)), stage s+1 on b_lo with
// w(s+1, (j7 | lane bits | b_hi)).  wb = w(s+1, block with j7 = 0, b_hi = 0) from the table; KA / KB: the root of unity that bit j7
// contributes at stage s / s+1 (exp(2 pi i / 2^(s+1)), exp(2 pi i / 2^(s+2)))
template <int KA, int KB>
struct TwoStageTw {
    cplx wa[2];      // [j7]
    cplx wb[2][2];   // [j7][b_hi]
    __device__ __forceinline__ explicit TwoStageTw(const cplx base) {
        wb[0][0] = base;
        wb[0][1] = cmul_i(base);                                                   // the block's low bit is the bit of the previous stage: times zeta^M = i
        wb[1][0] = cmul_c(base, RootOfUnity<KB>::c, RootOfUnity<KB>::s);
        wb[1][1] = cmul_i(wb[1][0]);
        wa[0] = csq(base);                                                          // w(s, b) = w(s+1, 2b)^2
        wa[1] = cmul_c(wa[0], RootOfUnity<KA>::c, RootOfUnity<KA>::s);
    }
    // the same from stored values (d[0] = wb[1][0], d[1] = wa[0], d[2] = wa[1]: what the constructor above computes)
    __device__ __forceinline__ TwoStageTw(const cplx base, const cplx (&d)[4]) {
        wb[0][0] = base;
        wb[0][1] = cmul_i(base);
        wb[1][0] = d[0];
        wb[1][1] = cmul_i(d[0]);
        wa[0] = d[1];
        wa[1] = d[2];
    }
};
// forward: registers r = (b_hi b_lo j7)
template <int KA, int KB>
__device__ __forceinline__ void fwd_two_stages(cplx (&x)[8], const TwoStageTw<KA, KB> &tw) {
#pragma unroll
    for (int r = 0; r < 4; r++) ct_bfly(x[r], x[r + 4], tw.wa[r & 1]);
#pragma unroll
    for (int r = 0; r < 8; r++)
        if ((r & 2) == 0) ct_bfly(x[r], x[r + 2], tw.wb[r & 1][(r >> 2) & 1]);
}
template <int KA, int KB>
__device__ __forceinline__ void inv_two_stages(cplx (&x)[8], const TwoStageTw<KA, KB> &tw) {
#pragma unroll
    for (int r = 0; r < 8; r++)
        if ((r & 2) == 0) gs_bfly(x[r], x[r + 2], tw.wb[r & 1][(r >> 2) & 1]);
#pragma unroll
    for (int r = 0; r < 4; r++) gs_bfly(x[r], x[r + 4], tw.wa[r & 1]);
}
// last stage: registers r = (j0 j6 j7), twiddle w(7, (j7 j6 | lane bits)) = base * root256^j7 * root128^j6
struct LastStageTw {
    cplx w[2][2];   // [j7][j6]
    __device__ __forceinline__ explicit LastStageTw(const cplx base) {
        w[0][0] = base;
        w[1][0] = cmul_c(base, RootOfUnity<8>::c, RootOfUnity<8>::s);
        w[0][1] = cmul_c(base, RootOfUnity<7>::c, RootOfUnity<7>::s);
        w[1][1] = cmul_c(w[1][0], RootOfUnity<7>::c, RootOfUnity<7>::s);
    }
    __device__ __forceinline__ LastStageTw(const cplx base, const cplx (&d)[4]) {   // d = w[1][0], w[0][1], w[1][1] as computed above (, base)
        w[0][0] = base;
        w[1][0] = d[0];
        w[0][1] = d[1];
        w[1][1] = d[2];
    }
};
// register renaming between the passes: (b_hi b_lo j7) -> (j7 b_hi b_lo), so that the two bits just processed leave with the swap
__device__ __forceinline__ void rename_out(cplx (&x)[8]) {
    cplx y[8];
#pragma unroll
    for (int r = 0; r < 8; r++) y[((r & 1) << 2) | (r >> 1)] = x[r];
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = y[r];
}
__device__ __forceinline__ void rename_in(cplx (&x)[8]) {   // the inverse renaming: (j7 b_hi b_lo) -> (b_hi b_lo j7)
    cplx y[8];
#pragma unroll
    for (int r = 0; r < 8; r++) y[((r & 3) << 1) | (r >> 2)] = x[r];
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = y[r];
}

// per-thread table entries (host_tables_fft.hpp build_fft_tmem_table): [0..3] w(4, (0 j6 j5 0)) by lane & 3,
// [4..19] w(6, (0 j6 j5 j4 j3 0)) by lane & 15, [20..51] w(7, (0 0 j5 j4 j3 j2 j1)) by lane
struct TmemTw {
    cplx p2, p3, p4;
    uint32_t cols;   // tensor-memory address of this warp's 40 twiddle columns (TFHE_TMEM_TWSTORE)
};
__device__ __forceinline__ TmemTw load_tmem_tw(const cplx *table, uint32_t lane) {
    TmemTw t;
    t.p2 = table[lane & 3u];
    t.p3 = table[4u + (lane & 15u)];
    t.p4 = table[20u + lane];
    t.cols = 0;
    return t;
}
// A thread's derived twiddles never change.  They do not fit in registers, but they fit in tensor memory: computed once per kernel,
// stored in 48 of the warp's columns ([0, 16) pass 2: wb[1][0], wa[0], wa[1], the table entry itself; [16, 32) pass 3; [32, 48) last
// stage: w[1][0], w[0][1], w[1][1], entry) and fetched with one tcgen05.ld.x16 per pass instead of 13 / 13 / 12 FP64 operations (the
// same values: same bits).  TFHE_TMEM_TWBASE: the table entries are taken from there too instead of 12 registers.
__device__ __forceinline__ void tmem_store_cplx(uint32_t taddr, const cplx v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"((uint32_t)__double2loint(v.re)), "r"((uint32_t)__double2hiint(v.re)),
                 "r"((uint32_t)__double2loint(v.im)), "r"((uint32_t)__double2hiint(v.im))
                 : "memory");
}
__device__ __forceinline__ void tmem_tw_setup(TmemTw &tw, uint32_t cols) {
    tw.cols = cols;
    const TwoStageTw<4, 5> a(tw.p2);
    const TwoStageTw<6, 7> b(tw.p3);
    const LastStageTw l(tw.p4);
    tmem_store_cplx(cols + 0u, a.wb[1][0]);
    tmem_store_cplx(cols + 4u, a.wa[0]);
    tmem_store_cplx(cols + 8u, a.wa[1]);
    tmem_store_cplx(cols + 12u, tw.p2);
    tmem_store_cplx(cols + 16u, b.wb[1][0]);
    tmem_store_cplx(cols + 20u, b.wa[0]);
    tmem_store_cplx(cols + 24u, b.wa[1]);
    tmem_store_cplx(cols + 28u, tw.p3);
    tmem_store_cplx(cols + 32u, l.w[1][0]);
    tmem_store_cplx(cols + 36u, l.w[0][1]);
    tmem_store_cplx(cols + 40u, l.w[1][1]);
    tmem_store_cplx(cols + 44u, tw.p4);
    tmem_wait_st();
}
struct TwRaw {
    uint32_t u[16];
};
__device__ __forceinline__ void tmem_tw_request(TwRaw &r, uint32_t taddr) {   // asynchronous: completed by the wait of the next swap
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r.u[0]), "=r"(r.u[1]), "=r"(r.u[2]), "=r"(r.u[3]), "=r"(r.u[4]), "=r"(r.u[5]), "=r"(r.u[6]), "=r"(r.u[7]), "=r"(r.u[8]), "=r"(r.u[9]),
                   "=r"(r.u[10]), "=r"(r.u[11]), "=r"(r.u[12]), "=r"(r.u[13]), "=r"(r.u[14]), "=r"(r.u[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_tw_claim(const TwRaw &r0, cplx (&d)[4]) {   // after a tcgen05.wait::ld that follows the request
    TwRaw r = r0;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.u[0]), "+r"(r.u[1]), "+r"(r.u[2]), "+r"(r.u[3]), "+r"(r.u[4]), "+r"(r.u[5]), "+r"(r.u[6]), "+r"(r.u[7]), "+r"(r.u[8]), "+r"(r.u[9]),
                   "+r"(r.u[10]), "+r"(r.u[11]), "+r"(r.u[12]), "+r"(r.u[13]), "+r"(r.u[14]), "+r"(r.u[15])
                 :
                 : "memory");
#pragma unroll
    for (int k = 0; k < 4; k++) {
        d[k].re = __hiloint2double((int)r.u[4 * k + 1], (int)r.u[4 * k]);
        d[k].im = __hiloint2double((int)r.u[4 * k + 3], (int)r.u[4 * k + 2]);
    }
}
#ifndef TFHE_TMEM_TWSTORE
#define TFHE_TMEM_TWSTORE 1
#endif
#ifndef TFHE_TMEM_TWBASE
#define TFHE_TMEM_TWBASE 1
#endif
constexpr uint32_t TMEM_TW_COLS = 48;
// forward transform after pass A (x in layout A: registers (j7 j6 j5), stages 0..2 done) -> layout F
__device__ __forceinline__ void tmem_fwd_rest(cplx (&x)[8], uint32_t taddr, const TmemTw &tw) {
#if TFHE_TMEM_TWSTORE
    TwRaw raw;
    cplx d[4];
    tmem_swap2_store(x, taddr);           // j7 stays, (j6 j5) leave: regs (j4 j3 j7)
    tmem_tw_request(raw, tw.cols);          // x is dead here: the request costs no registers
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
    tmem_tw_claim(raw, d);
    fwd_two_stages<4, 5>(x, TwoStageTw<4, 5>(TFHE_TMEM_TWBASE ? d[3] : tw.p2, d));   // stages 3, 4
    rename_out(x);                        // (j7 j4 j3)
    tmem_swap2_store(x, taddr);           // regs (j2 j1 j7)
    tmem_tw_request(raw, tw.cols + 16u);          // x is dead here: the request costs no registers
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
    tmem_tw_claim(raw, d);
    fwd_two_stages<6, 7>(x, TwoStageTw<6, 7>(TFHE_TMEM_TWBASE ? d[3] : tw.p3, d));   // stages 5, 6
    rename_out(x);                        // (j7 j2 j1)
    tmem_swap2_store(x, taddr);           // regs (j0 j6 j7)
    tmem_tw_request(raw, tw.cols + 32u);          // x is dead here: the request costs no registers
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
    tmem_tw_claim(raw, d);
    const LastStageTw l(TFHE_TMEM_TWBASE ? d[3] : tw.p4, d);
#else
    tmem_swap2(x, taddr);                 // j7 stays, (j6 j5) leave: regs (j4 j3 j7)
    fwd_two_stages<4, 5>(x, TwoStageTw<4, 5>(tw.p2));   // stages 3, 4
    rename_out(x);                        // (j7 j4 j3)
    tmem_swap2(x, taddr);                 // regs (j2 j1 j7)
    fwd_two_stages<6, 7>(x, TwoStageTw<6, 7>(tw.p3));   // stages 5, 6
    rename_out(x);                        // (j7 j2 j1)
    tmem_swap2(x, taddr);                 // regs (j0 j6 j7)
    const LastStageTw l(tw.p4);
#endif
#pragma unroll
    for (int r = 0; r < 4; r++) ct_bfly(x[r], x[r + 4], l.w[r & 1][(r >> 1) & 1]);   // stage 7
}
// inverse of two accumulators (the low- and high-limb products of a column) from layout F back to layout A (stages 7..3 undone;
// the caller finishes with the inverse of pass A); the two use disjoint column ranges so that their transfers overlap
__device__ __forceinline__ void tmem_unswap2_pair(cplx (&a)[8], cplx (&b)[8], uint32_t taddr, uint32_t taddr_b) {
    tmem_unswap2_store(a, taddr);
    tmem_unswap2_store(b, taddr_b);
    tmem_wait_st();
    tmem_unswap2_load(a, taddr);
    tmem_unswap2_load(b, taddr_b);
}
// taddr_b: 32 more columns for the second accumulator -- the caller passes the publish buffer that is free at this point
__device__ __forceinline__ void tmem_inv_rest(cplx (&a)[8], cplx (&b)[8], uint32_t taddr, uint32_t taddr_b, const TmemTw &tw) {
#if TFHE_TMEM_TWSTORE
    TwRaw raw;
    cplx d[4];
    tmem_tw_request(raw, tw.cols + 32u);
    tmem_tw_claim(raw, d);
    {
        const LastStageTw l(TFHE_TMEM_TWBASE ? d[3] : tw.p4, d);
#else
    {
        const LastStageTw l(tw.p4);
#endif
#pragma unroll
        for (int r = 0; r < 4; r++) {
            gs_bfly(a[r], a[r + 4], l.w[r & 1][(r >> 1) & 1]);
            gs_bfly(b[r], b[r + 4], l.w[r & 1][(r >> 1) & 1]);
        }
    }
#if TFHE_TMEM_TWSTORE
    tmem_unswap2_store(a, taddr);         // regs (j7 j2 j1)
    tmem_unswap2_store(b, taddr_b);
    tmem_tw_request(raw, tw.cols + 16u);
    tmem_wait_st();
    tmem_unswap2_load(a, taddr);
    tmem_unswap2_load(b, taddr_b);
    tmem_tw_claim(raw, d);
    const TwoStageTw<6, 7> t3(TFHE_TMEM_TWBASE ? d[3] : tw.p3, d);
#else
    tmem_unswap2_pair(a, b, taddr, taddr_b);       // regs (j7 j2 j1)
    const TwoStageTw<6, 7> t3(tw.p3);
#endif
    rename_in(a); rename_in(b);           // (j2 j1 j7)
    inv_two_stages<6, 7>(a, t3);
    inv_two_stages<6, 7>(b, t3);
#if TFHE_TMEM_TWSTORE
    tmem_unswap2_store(a, taddr);         // regs (j7 j4 j3)
    tmem_unswap2_store(b, taddr_b);
    tmem_tw_request(raw, tw.cols);
    tmem_wait_st();
    tmem_unswap2_load(a, taddr);
    tmem_unswap2_load(b, taddr_b);
    tmem_tw_claim(raw, d);
    const TwoStageTw<4, 5> t2(TFHE_TMEM_TWBASE ? d[3] : tw.p2, d);
#else
    tmem_unswap2_pair(a, b, taddr, taddr_b);       // regs (j7 j4 j3)
    const TwoStageTw<4, 5> t2(tw.p2);
#endif
    rename_in(a); rename_in(b);           // (j4 j3 j7)
    inv_two_stages<4, 5>(a, t2);
    inv_two_stages<4, 5>(b, t2);
    tmem_unswap2_pair(a, b, taddr, taddr_b);       // regs (j7 j6 j5): layout A
}

// ---- rows published through tensor memory.  With team = lane quarter (warp % 4) the P sub-teams (warps) of a ciphertext share
// their 32 lanes, and the multiply-accumulate is pointwise: thread t of every sub-team needs exactly the points thread t of the
// publishing sub-team holds.  A transformed digit row is therefore published with one store of the thread's 8 points into the
// publisher's columns (column 4 e + {re lo, re hi, im lo, im hi}) and fetched by the peers with one load of the same lane.
__device__ __forceinline__ void tmem_store_row(const cplx (&x)[8], uint32_t taddr) {
    uint32_t v[32];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        v[4 * e] = (uint32_t)__double2loint(x[e].re);
        v[4 * e + 1] = (uint32_t)__double2hiint(x[e].re);
        v[4 * e + 2] = (uint32_t)__double2loint(x[e].im);
        v[4 * e + 3] = (uint32_t)__double2hiint(x[e].im);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
  // the caller waits (tmem_wait_st) before it tells its peers
}
__device__ __forceinline__ void tmem_load_row(cplx (&x)[8], uint32_t taddr) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
          "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    tmem_wait_ld(v);
#pragma unroll
    for (int e = 0; e < 8; e++) {
        x[e].re = __hiloint2double((int)v[4 * e + 1], (int)v[4 * e]);
        x[e].im = __hiloint2double((int)v[4 * e + 3], (int)v[4 * e + 2]);
    }
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tensor-memory columns of a sub-team (warp): [0, 32) exchanges of the transforms, [32, 64) and [64, 96) the published row, double
// buffered by level parity (the paired inverse exchanges its second accumulator through the buffer that is free at that point)
constexpr uint32_t TMEM_SUB_COLS = 96, TMEM_PUB_COL = 32;

// ---- derived pass twiddles kept in tensor memory (FftPbsCfg::TWT: kernels whose exchanges stay in shared memory).  A block = up to 8
// complex values of a thread in 32 columns of its lane: written once per kernel, fetched with one tcgen05.ld.x32 per pass.
struct TwRaw32 {
    uint32_t u[32];
};
template <int N>
__device__ __forceinline__ void tmem_tw_block_store(uint32_t taddr, const cplx *tw) {
    static_assert(N <= 8, "one block holds 8 complex values");
#pragma unroll
    for (int k = 0; k < N; k++) tmem_store_cplx(taddr + 4u * k, tw[k]);
}
__device__ __forceinline__ void tmem_tw_block_request(TwRaw32 &r, uint32_t taddr) {   // asynchronous
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r.u[0]), "=r"(r.u[1]), "=r"(r.u[2]), "=r"(r.u[3]), "=r"(r.u[4]), "=r"(r.u[5]), "=r"(r.u[6]), "=r"(r.u[7]), "=r"(r.u[8]), "=r"(r.u[9]), "=r"(r.u[10]),
          "=r"(r.u[11]), "=r"(r.u[12]), "=r"(r.u[13]), "=r"(r.u[14]), "=r"(r.u[15]), "=r"(r.u[16]), "=r"(r.u[17]), "=r"(r.u[18]), "=r"(r.u[19]), "=r"(r.u[20]),
          "=r"(r.u[21]), "=r"(r.u[22]), "=r"(r.u[23]), "=r"(r.u[24]), "=r"(r.u[25]), "=r"(r.u[26]), "=r"(r.u[27]), "=r"(r.u[28]), "=r"(r.u[29]), "=r"(r.u[30]),
          "=r"(r.u[31])
        : "r"(taddr)
        : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_tw_block_claim(TwRaw32 &r, cplx *tw) {
    tmem_wait_ld(r.u);
#pragma unroll
    for (int k = 0; k < N; k++) {
        tw[k].re = __hiloint2double((int)r.u[4 * k + 1], (int)r.u[4 * k]);
        tw[k].im = __hiloint2double((int)r.u[4 * k + 3], (int)r.u[4 * k + 2]);
    }
}
constexpr uint32_t TMEM_TWT_COLS = 64;   // per warp: pass B at +0, pass C at +32

// ---- M = 512, two warps per sub-team (FftPbsCfg::TAIL).  Pass A and pass B run as in fft_team.cuh with ONE shared-memory exchange
// between them (layout A -> layout B'': fft_team.cuh store_Asw / load_Bsw); the last three stages stay inside each warp:
//   layout B''  warp j8, lane (j2 j1 j0 j6 j7)  regs (j5 j4 j3)   stages 3 4 5
//   swap  ->    warp j8, lane (j0 j6 j7 j4 j3)  regs (j2 j1 j5)   stages 6 7      (TwoStageTw<4, 5>: j5 rides along, 3 bits above j2)
//   swap  ->    warp j8, lane (j7 j4 j3 j2 j1)  regs (j0 j6 j5)   stage  8        = layout F9: the spectral layout of this path
// A bit of the block index at position k (from the LSB) contributes the factor exp(2 pi i / 2^(k + 2)) to a twiddle, whatever the stage
// (host_tables_fft.hpp fft_twiddle): j5 at stage 8 sits at k = 4 (root 2^6), j6 at k = 5 (root 2^7).
template <int K0, int K1>   // roots contributed by register bit 0 / register bit 1
struct LastStageTwT {
    cplx w[2][2];   // [r0][r1]
    __device__ __forceinline__ explicit LastStageTwT(const cplx base) {
        w[0][0] = base;
        w[1][0] = cmul_c(base, RootOfUnity<K0>::c, RootOfUnity<K0>::s);
        w[0][1] = cmul_c(base, RootOfUnity<K1>::c, RootOfUnity<K1>::s);
        w[1][1] = cmul_c(w[1][0], RootOfUnity<K1>::c, RootOfUnity<K1>::s);
    }
    __device__ __forceinline__ explicit LastStageTwT(const cplx (&d)[4]) {   // stored: w[1][0], w[0][1], w[1][1], base
        w[0][0] = d[3];
        w[1][0] = d[0];
        w[0][1] = d[1];
        w[1][1] = d[2];
    }
};
struct TailTw {
    cplx p67, p8;    // w(7, (j8 j7 j6 0 j4 j3 0)), w(8, (j8 j7 0 0 j4 j3 j2 j1)) of this thread
    uint32_t cols;   // tensor-memory address of the stored copies (TFHE_TMEM_TAIL_TWSTORE): [0, 16) stages 6 7, [16, 32) stage 8
};
#ifndef TFHE_TMEM_TAIL_TWSTORE
#define TFHE_TMEM_TAIL_TWSTORE 1
#endif
__device__ __forceinline__ void tmem_tail_tw_setup(TailTw &tw, uint32_t cols) {
    tw.cols = cols;
    const TwoStageTw<4, 5> a(tw.p67);
    const LastStageTwT<6, 7> l(tw.p8);
    tmem_store_cplx(cols + 0u, a.wb[1][0]);
    tmem_store_cplx(cols + 4u, a.wa[0]);
    tmem_store_cplx(cols + 8u, a.wa[1]);
    tmem_store_cplx(cols + 12u, tw.p67);
    tmem_store_cplx(cols + 16u, l.w[1][0]);
    tmem_store_cplx(cols + 20u, l.w[0][1]);
    tmem_store_cplx(cols + 24u, l.w[1][1]);
    tmem_store_cplx(cols + 28u, tw.p8);
    tmem_wait_st();
}
// table (host_tables_fft.hpp build_fft_tail_table): [0..31] by (j8 j7 j6 j4 j3), [32..95] by (j8 j7 j4 j3 j2 j1); t = warp << 5 | lane
__device__ __forceinline__ TailTw load_tail_tw(const cplx *table, uint32_t t) {
    const uint32_t w = t >> 5, l = t & 31u;
    TailTw tw;
    // after the first swap: lane (j0 j6 j7 j4 j3)
    tw.p67 = table[(w << 4) | (((l >> 2) & 1u) << 3) | (((l >> 3) & 1u) << 2) | (l & 3u)];
    // after the second swap: lane (j7 j4 j3 j2 j1)
    tw.p8 = table[32u + ((w << 5) | l)];
    tw.cols = 0;
    return tw;
}
__device__ __forceinline__ void tmem_fwd_tail9(cplx (&x)[8], uint32_t taddr, const TailTw &tw) {
#if TFHE_TMEM_TAIL_TWSTORE
    TwRaw raw;
    cplx d[4];
    tmem_swap2_store(x, taddr);                             // regs (j2 j1 j5)
    tmem_tw_request(raw, tw.cols);
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
    tmem_tw_claim(raw, d);
    fwd_two_stages<4, 5>(x, TwoStageTw<4, 5>(d[3], d));     // stages 6, 7
    rename_out(x);                                          // (j5 j2 j1)
    tmem_swap2_store(x, taddr);                             // regs (j0 j6 j5)
    tmem_tw_request(raw, tw.cols + 16u);
    tmem_wait_st();
    tmem_swap2_load(x, taddr);
    tmem_tw_claim(raw, d);
    const LastStageTwT<6, 7> l(d);
#else
    tmem_swap2(x, taddr);                                   // regs (j2 j1 j5)
    fwd_two_stages<4, 5>(x, TwoStageTw<4, 5>(tw.p67));      // stages 6, 7
    rename_out(x);                                          // (j5 j2 j1)
    tmem_swap2(x, taddr);                                   // regs (j0 j6 j5)
    const LastStageTwT<6, 7> l(tw.p8);
#endif
#pragma unroll
    for (int r = 0; r < 4; r++) ct_bfly(x[r], x[r + 4], l.w[r & 1][(r >> 1) & 1]);   // stage 8
}
__device__ __forceinline__ void tmem_inv_tail9(cplx (&a)[8], cplx (&b)[8], uint32_t taddr, const TailTw &tw) {
#if TFHE_TMEM_TAIL_TWSTORE
    TwRaw raw;
    cplx d[4];
    tmem_tw_request(raw, tw.cols + 16u);
    tmem_tw_claim(raw, d);
    {
        const LastStageTwT<6, 7> l(d);
#else
    {
        const LastStageTwT<6, 7> l(tw.p8);
#endif
#pragma unroll
        for (int r = 0; r < 4; r++) {
            gs_bfly(a[r], a[r + 4], l.w[r & 1][(r >> 1) & 1]);
            gs_bfly(b[r], b[r + 4], l.w[r & 1][(r >> 1) & 1]);
        }
    }
#if TFHE_TMEM_TAIL_TWSTORE
    tmem_unswap2_store(a, taddr);                           // regs (j5 j2 j1)
    tmem_unswap2_store(b, taddr + 32u);
    tmem_tw_request(raw, tw.cols);
    tmem_wait_st();
    tmem_unswap2_load(a, taddr);
    tmem_unswap2_load(b, taddr + 32u);
    tmem_tw_claim(raw, d);
    const TwoStageTw<4, 5> t2(d[3], d);
#else
    tmem_unswap2_pair(a, b, taddr, taddr + 32u);            // regs (j5 j2 j1)
    const TwoStageTw<4, 5> t2(tw.p67);
#endif
    rename_in(a); rename_in(b);                             // (j2 j1 j5)
    inv_two_stages<4, 5>(a, t2);
    inv_two_stages<4, 5>(b, t2);
    tmem_unswap2_pair(a, b, taddr, taddr + 32u);            // regs (j5 j4 j3): layout B''
}
constexpr uint32_t TMEM_TAIL_COLS = 64;   // per warp: 32 per accumulator of the paired inverse
constexpr uint32_t TMEM_TAIL_TW_COLS = 32; // per warp, behind the swap columns of all warps of a lane quarter

// ---- M = 1024, 16 points per thread, two warps per sub-team (FftPbsCfg::TAIL16).  Pass A (stages 0..3) -> ONE shared-memory exchange
// (fft_team.cuh store_Asw16 / load_Bsw16) -> pass B'' (stages 4..7 on registers (j5 j4 j3 j2)) -> one swap inside the warp -> stages 8, 9:
//   layout B''16  warp j9, lane (j1 j0 j8 j7 j6)  regs (j5; j4 j3 j2)
//   swap  ->      warp j9, lane (j8 j7 j6 j3 j2)  regs (j5; j1 j0 j4)   stages 8 9     = layout F10
// The 16 registers are two halves (j5 = 0 / 1) swapped side by side in 2 x 32 columns; each half is the 8-register pass of the M = 256 /
// 512 kernels with j4 riding along (TwoStageTw<4, 5>), and j5 contributes the root 2^6 to the stage-9 twiddle of its half.
struct Tail16Tw {
    cplx pB;         // w(7, hA << 3): base of pass B''
    cplx p9;         // w(9, (j9 j8 j7 j6 0 0 j3 j2 0)): base of the last two stages, half j5 = 0
    uint32_t cols;   // tensor-memory address of the parked stage-8/9 twiddles: 2 halves x (wb[1][0], wa[0], wa[1], base) = 32 columns
};
// table (host_tables_fft.hpp build_fft_tail16_table): [0..15] by hA = (j9 j8 j7 j6), [16..79] by (j9 j8 j7 j6 j3 j2); t = warp << 5 | lane
__device__ __forceinline__ Tail16Tw load_tail16_tw(const cplx *table, uint32_t t, uint32_t hA) {
    Tail16Tw tw;
    tw.pB = table[hA];
    tw.p9 = table[16u + t];   // after the swap: warp j9, lane (j8 j7 j6 j3 j2)
    tw.cols = 0;
    return tw;
}
__device__ __forceinline__ void tmem_tail16_tw_setup(Tail16Tw &tw, uint32_t cols) {
    tw.cols = cols;
    const cplx b1 = cmul_c(tw.p9, RootOfUnity<6>::c, RootOfUnity<6>::s);
    const TwoStageTw<4, 5> h0(tw.p9), h1(b1);
    tmem_store_cplx(cols + 0u, h0.wb[1][0]);
    tmem_store_cplx(cols + 4u, h0.wa[0]);
    tmem_store_cplx(cols + 8u, h0.wa[1]);
    tmem_store_cplx(cols + 12u, tw.p9);
    tmem_store_cplx(cols + 16u, h1.wb[1][0]);
    tmem_store_cplx(cols + 20u, h1.wa[0]);
    tmem_store_cplx(cols + 24u, h1.wa[1]);
    tmem_store_cplx(cols + 28u, b1);
    tmem_wait_st();
}
__device__ __forceinline__ void tmem_fwd_tail16(cplx (&x)[16], uint32_t taddr, const Tail16Tw &tw) {
    cplx(&lo)[8] = *reinterpret_cast<cplx(*)[8]>(&x[0]);
    cplx(&hi)[8] = *reinterpret_cast<cplx(*)[8]>(&x[8]);
    TwRaw32 raw;
    cplx d[8];
    tmem_swap2_store(lo, taddr);
    tmem_swap2_store(hi, taddr + 32u);
    tmem_tw_block_request(raw, tw.cols);      // x is dead here
    tmem_wait_st();
    tmem_swap2_load(lo, taddr);
    tmem_swap2_load(hi, taddr + 32u);
    tmem_tw_block_claim<8>(raw, d);
    const cplx d0[4] = {d[0], d[1], d[2], d[3]}, d1[4] = {d[4], d[5], d[6], d[7]};
    fwd_two_stages<4, 5>(lo, TwoStageTw<4, 5>(d0[3], d0));   // stages 8, 9, half j5 = 0
    fwd_two_stages<4, 5>(hi, TwoStageTw<4, 5>(d1[3], d1));   // half j5 = 1
}
__device__ __forceinline__ void tmem_inv_tail16(cplx (&x)[16], uint32_t taddr, const Tail16Tw &tw) {
    cplx(&lo)[8] = *reinterpret_cast<cplx(*)[8]>(&x[0]);
    cplx(&hi)[8] = *reinterpret_cast<cplx(*)[8]>(&x[8]);
    TwRaw32 raw;
    cplx d[8];
    tmem_tw_block_request(raw, tw.cols);
    tmem_tw_block_claim<8>(raw, d);
    const cplx d0[4] = {d[0], d[1], d[2], d[3]}, d1[4] = {d[4], d[5], d[6], d[7]};
    inv_two_stages<4, 5>(lo, TwoStageTw<4, 5>(d0[3], d0));
    inv_two_stages<4, 5>(hi, TwoStageTw<4, 5>(d1[3], d1));
    tmem_unswap2_store(lo, taddr);
    tmem_unswap2_store(hi, taddr + 32u);
    tmem_wait_st();
    tmem_unswap2_load(lo, taddr);
    tmem_unswap2_load(hi, taddr + 32u);   // regs (j5; j4 j3 j2): layout B''16
}
constexpr uint32_t TMEM_TAIL16_COLS = 64, TMEM_TAIL16_TW_COLS = 32;   // per warp

}  // namespace fft
}  // namespace tfhe
