// kernels_ks_tcgen05.cuh -- the key-switching product on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM, TMA).
//
// Same arithmetic as ks_mma_kernel (kernels.cuh K4-MMA), exactly: key_switching.rs:88 is the matrix product
//     out[b][c] = -sum_r D[b][r] KSK[r][c]  (mod 2^32),  D = small signed digits (int8), KSK = sum_pl 2^(8 pl) byte_pl
// so  sum_r D KSK = sum_pl 2^(8 pl) (sum_r D[b][r] byte_pl[r][c])  and every inner sum is an s8 x u8 dot product with an exact
// 32-bit integer result (|sum| <= KD * 2^beta * 255 < 2^31, checked at key upload).  A = digits [B][KD] (K-major, as
// ks_digits_kernel writes them), B = the byte-transposed key [4 (n+1)'][KD] (K-major, built once at upload): a plain TN GEMM
// with s32 accumulation, M = batch, N = 4 x key columns, K = KD.
//
// Blackwell-native form: one CTA = a 128 x 256 tile of the product; a TMA producer thread streams 128-byte-wide K slabs of
// both operands into a 4-stage shared-memory ring (cp.async.bulk.tensor.2d, 128-byte swizzle), one thread issues
// tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32 per instruction) with the accumulator in tensor memory
// (256 columns), tcgen05.commit hands ring slots back to the producer and the finished accumulator to four epilogue warps,
// which read their 32 TMEM lanes with tcgen05.ld, recombine the four byte planes of every output word with shifts mod 2^32,
// negate, add the body to the last column (key_switching.rs:92-100) and store.  All waits are bounded (mbar_wait).
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace tfhe {

constexpr int KT_BM = 128, KT_BN = 256, KT_BK = 128, KT_STAGES = 4, KT_THREADS = 256;
constexpr int KT_A_BYTES = KT_BM * KT_BK, KT_B_BYTES = KT_BN * KT_BK, KT_STAGE_BYTES = KT_A_BYTES + KT_B_BYTES;
constexpr int KT_SMEM = KT_STAGES * KT_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers, tmem pointer */;

// tcgen05 instruction descriptor, kind::i8: D = S32 (c_format 2 at bit 4), A = signed 8 bit (1 at bit 7), B = unsigned 8 bit (0 at
// bit 10), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
constexpr uint32_t KT_IDESC = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(KT_BN >> 3) << 17) | ((uint32_t)(KT_BM >> 4) << 24);
// shared-memory matrix descriptor of a K-major, 128-byte-swizzled tile (rows of 128 bytes, 8-row groups 1024 bytes apart):
// start address >> 4, leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 >> 4, version 1 (Blackwell),
// layout type 2 = SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t kt_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void kt_tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void kt_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a),
        "l"(desc_b), "r"(KT_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void kt_commit(uint64_t *bar) {   // arrives on the mbarrier when all previously issued MMAs have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(KT_THREADS, 1)
ks_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const uint32_t *__restrict__ body,
                  uint32_t *__restrict__ out, uint32_t KD, uint32_t n, uint32_t batch, uint32_t *err_flag) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle-128B tiles: 1024-byte aligned
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + KT_STAGES * KT_STAGE_BYTES), *empty = full + KT_STAGES, *acc_full = empty + KT_STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n0 = blockIdx.x * KT_BN, m0 = blockIdx.y * KT_BM, nk = KD / KT_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KT_STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) {   // one warp allocates the accumulator: 256 TMEM columns x 128 lanes x 32 bit
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(KT_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer
        for (uint32_t kb = 0; kb < nk; kb++) {
            const uint32_t s = kb % KT_STAGES, use = kb / KT_STAGES;
            if (use > 0) mbar_wait(empty + s, (use - 1u) & 1u, err_flag);
            mbar_expect_tx(full + s, KT_STAGE_BYTES);
            kt_tma_load_2d(smem + s * KT_STAGE_BYTES, &map_a, (int)(kb * KT_BK), (int)m0, full + s);
            kt_tma_load_2d(smem + s * KT_STAGE_BYTES + KT_A_BYTES, &map_b, (int)(kb * KT_BK), (int)n0, full + s);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (one thread)
        for (uint32_t kb = 0; kb < nk; kb++) {
            const uint32_t s = kb % KT_STAGES, use = kb / KT_STAGES;
            mbar_wait(full + s, use & 1u, err_flag);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_u32(smem + s * KT_STAGE_BYTES), b_addr = a_addr + KT_A_BYTES;
#pragma unroll
            for (uint32_t k = 0; k < KT_BK / 32; k++)   // K = 32 bytes per instruction: advance inside the 128-byte swizzled row
                kt_mma_i8(tmem_acc, kt_smem_desc(a_addr + k * 32u), kt_smem_desc(b_addr + k * 32u), (kb | k) != 0u);
            kt_commit(empty + s);                       // the slot is free once these MMAs have read it
        }
        kt_commit(acc_full);                            // the accumulator is complete
    } else if (warp >= 4) {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = rows of the tile
        mbar_wait(acc_full, 0u, err_flag);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t q = warp & 3u, b = m0 + q * 32u + lane, ncols = n + 1;
        const uint32_t bodyv = b < batch ? body[b] : 0u;
#pragma unroll 1
        for (uint32_t cc = 0; cc < KT_BN; cc += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_acc + ((q * 32u) << 16) + cc;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int w = 0; w < 4; w++) {   // byte planes 0..3 of key column c sit in four consecutive accumulator columns
                const uint32_t c = (n0 + cc) / 4u + (uint32_t)w;
                const uint32_t sum = v[4 * w] + (v[4 * w + 1] << 8) + (v[4 * w + 2] << 16) + (v[4 * w + 3] << 24);
                if (b < batch && c < ncols) {
                    uint32_t r = 0u - sum;          // key_switching.rs:96 negate
                    if (c == n) r += bodyv;         // key_switching.rs:98-100
                    out[(size_t)b * ncols + c] = r;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(KT_BN) : "memory");
    }
}

}  // namespace tfhe
