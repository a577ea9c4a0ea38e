/*
 * tfhe_b200.h -- C ABI of the B200-native TFHE programmable-bootstrapping path.
 *
 * Drop-in boundary for the hot path of Janmajayamall/tfhe-research (a Rust crate with no FFI of its
 * own): every entry point below replaces one crate-internal function and keeps its
 * argument meaning, data layout (flat little-endian u32, row-major, exactly the ndarray layouts of
 * the reference) and bit-exact results.  Citations are `file:line` into the reference's src/.
 * INTEGRATION.md shows the Rust `extern "C"` shim a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative tfhe_status; nothing unwinds across the ABI
 *    (the reference panics via assert!/unwrap -- those sites map to TFHE_E_ASSERT);
 *  - a tfhe_ctx is bound to ONE CUDA device and is single-caller, like the reference's synchronous
 *    single-threaded calls; multi-GPU = either one process (one ctx) per GPU with the batch sharded by the
 *    caller, or ONE process driving all GPUs of the box through a tfhe_mgpu (section "multi-GPU" below);
 *  - pointer arguments of the batched device entry points may be host OR device pointers (queried
 *    with cudaPointerGetAttributes); host buffers are copied with cudaMemcpyAsync into device staging
 *    buffers owned by the ctx (pass pinned host memory for full-speed, truly asynchronous copies);
 *  - there is NO CPU fallback: device entry points fail with TFHE_E_CUDA when no GPU is present.
 *
 * Layouts:  N = 2^log_poly_degree, k = glwe_dimension, n = lwe_dimension,
 *  l = pbs_levels, l_ks = ks_levels.
 *   LWE  ciphertext  u32[n+1]                 (a_0..a_{n-1}, b)  body LAST        lwe.rs:110-115
 *   GLWE ciphertext  u32[k+1][N]              rows 0..k-1 masks, row k body       glwe.rs:186-188
 *   GGSW ciphertext  u32[(k+1)*l][k+1][N]     row = poly*l + level                ggsw.rs:37-41
 *   BSK              u32[n][(k+1)*l][k+1][N]                                      bootstrapping.rs:18-21
 *   KSK              u32[k*N*l_ks][n+1]       row = s_index*l_ks + level          key_switching.rs:13-15
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tfhe_status {
    TFHE_OK = 0,
    TFHE_E_PARAM = -1,   /* invalid / unsupported parameter set or argument */
    TFHE_E_CUDA = -2,    /* CUDA runtime error, or no CUDA device (there is no CPU fallback) */
    TFHE_E_OOM = -3,
    TFHE_E_ASSERT = -4,  /* a reference assert! would have fired (glwe.rs:144, lwe.rs:84, test_vector.rs:41, bootstrapping.rs:127) */
    TFHE_E_NCCL = -5
} tfhe_status;

/* lib.rs:23-34 TfheParams (glwe_poly_degree holds log2 N exactly like the reference's field) */
typedef struct tfhe_params {
    uint32_t glwe_dimension;
    uint32_t glwe_poly_degree; /* log2 N */
    uint32_t lwe_dimension;
    uint32_t padding_bits;
    uint32_t log_p;
    uint32_t log_q;            /* must be 32 */
    uint32_t ks_log_base, ks_levels;
    uint32_t pbs_log_base, pbs_levels;
    double lwe_std_dev, glwe_std_dev;
} tfhe_params;

/* lib.rs:76-124 `impl Default for TfheParams`; test_cfg != 0 gives the cfg(test) variant (n = 4). */
int tfhe_params_default(int test_cfg, tfhe_params *out);
/* Named presets of SURVEY 8(d): "P0" (= default), "P0t" (= cfg(test)), "P1", "P2". */
int tfhe_params_preset(const char *name, tfhe_params *out);
/* TFHE_E_PARAM unless: log_q == 32, log_base | 32, log_base*levels <= 32 (decomposer.rs, H2),
 * 2^log_p | N, and (N, k, l, log_base) is one of the instantiated kernel configurations. */
int tfhe_params_validate(const tfhe_params *p);

/* gate opcodes for tfhe_gate_batch / tfhe_test_vector_boolean.  AND, OR follow boolean.rs:9-53; XOR is
 * the same construction with f = l ^ r; NAND/NOR/XNOR are trivial(1) - gate. */
typedef enum tfhe_gate { TFHE_AND = 0, TFHE_OR = 1, TFHE_XOR = 2, TFHE_NAND = 3, TFHE_NOR = 4, TFHE_XNOR = 5 } tfhe_gate;

/* ------------------------------------------------------------------ host-side (C++) mirror */
/* test_vector.rs:38-67 / :23-35 / :5-20 */
int tfhe_test_vector_from_lut(const tfhe_params *p, const uint32_t *lut, size_t lut_len, uint32_t *tv_out /* N */);
int tfhe_test_vector_identity(const tfhe_params *p, uint32_t *tv_out /* N */);
int tfhe_test_vector_boolean(const tfhe_params *p, int gate /* AND, OR or XOR */, uint32_t *tv_out /* N */);
/* lwe.rs:83-88 / :102-107 (decode is the reference's floor shift, H5) / :138-160 / :162-173 */
int tfhe_lwe_encode(const tfhe_params *p, uint32_t message, uint32_t *plaintext_out);
int tfhe_lwe_decode(const tfhe_params *p, uint32_t plaintext, uint32_t *message_out);
int tfhe_lwe_encrypt(const tfhe_params *p, const uint32_t *lwe_sk, size_t sk_len, uint32_t plaintext, uint64_t seed,
                     uint64_t index, uint32_t *ct_out /* sk_len+1 */);
int tfhe_lwe_decrypt(const uint32_t *lwe_sk, size_t sk_len, const uint32_t *ct, uint32_t *plaintext_out);
/* bootstrapping.rs:23-56 bootstrapping_key_gen (+ LweSecretKey::random lwe.rs:54-58, GlweSecretKey::random
 * glwe.rs:177-181).  Seeded (thread_rng is not reproducible); multi-threaded on the host. */
int tfhe_keygen(const tfhe_params *p, uint64_t seed, uint32_t *lwe_sk /* n */, uint32_t *glwe_sk /* k*N */,
                uint32_t *bsk, uint32_t *ksk);

/* ------------------------------------------------------------------ device context and keys */
typedef struct tfhe_ctx tfhe_ctx;
typedef struct tfhe_bk tfhe_bk; /* device-resident BootstrappingKey (bootstrapping.rs:18-21) */

int tfhe_ctx_create(const tfhe_params *p, int device, tfhe_ctx **out);
void tfhe_ctx_destroy(tfhe_ctx *ctx);
const char *tfhe_last_error(const tfhe_ctx *ctx);
/* Run all subsequent work of this ctx on an existing CUDA stream (cudaStream_t as void*); NULL = own stream. */
int tfhe_ctx_set_stream(tfhe_ctx *ctx, void *cuda_stream);
/* The stream this ctx launches on (cudaStream_t as void*). */
void *tfhe_ctx_get_stream(const tfhe_ctx *ctx);
/* Kernels launched by this ctx since creation (the bench's gpu_launches claim). */
uint64_t tfhe_ctx_launch_count(const tfhe_ctx *ctx);

/* Arithmetic path of the external product (both reproduce ggsw.rs:132-161 bit for bit):
 *   TFHE_PATH_NTT  2-prime RNS number-theoretic transform + CRT (integer IMAD pipe), every parameter set;
 *   TFHE_PATH_FFT  double-precision folded FFT with a limb-split key, provably exact, FP64 pipe;
 *                  instantiated for the parameter sets listed by tfhe_ctx_set_pbs_path's TFHE_E_PARAM.
 * The path is fixed per bootstrapping key: select it BEFORE tfhe_bk_upload (default: env TFHE_B200_PBS_PATH=ntt|fft,
 * else the faster instantiated path for the parameter set). */
#define TFHE_PATH_NTT 0
#define TFHE_PATH_FFT 1
int tfhe_ctx_set_pbs_path(tfhe_ctx *ctx, int path);
int tfhe_ctx_get_pbs_path(const tfhe_ctx *ctx);
/* Arithmetic of the key-switching product key_switching.rs:88 (out = -D . KSK mod 2^32), batches above 8 ciphertexts
 * (both give the reference's bits):
 *   TFHE_KS_IMAD  wrapping 32-bit multiply-adds on the integer pipe, every parameter set;
 *   TFHE_KS_MMA   integer tensor cores, exact through the four byte planes of the key (s8 x u8 -> s32 sums, recombined
 *                 mod 2^32); needs k*N*ks_levels % 64 == 0 and k*N*ks_levels * 2^ks_log_base * 255 < 2^31, else IMAD runs.
 *   TFHE_KS_TCGEN05  the same byte-plane product on the 5th-generation tensor cores: tcgen05.mma kind::i8 (s8 x u8 -> s32) with
 *                 the accumulator in tensor memory and both operands streamed by TMA (kernels_ks_tcgen05.cuh); additionally
 *                 needs k*N*ks_levels % 128 == 0, else MMA runs.
 * Default: env TFHE_B200_KS=imad|mma|tcgen05, else the fastest that applies.  May be switched at any time. */
#define TFHE_KS_IMAD 0
#define TFHE_KS_MMA 1
#define TFHE_KS_TCGEN05 2
int tfhe_ctx_set_ks_path(tfhe_ctx *ctx, int path);
/* FFT path only.  Exactness rests on the a-priori error bound (2^-9 before rounding).  With checking
 * switched on (env TFHE_B200_FFT_CHECK=1 or tfhe_ctx_set_fft_check) the blind rotation runs a kernel variant that also
 * records the largest distance to the nearest integer of every value it rounds (about 5 % slower);
 * tfhe_fft_rounding_margin returns that maximum since the last call and resets it (0 when nothing was recorded). */
int tfhe_ctx_set_fft_check(tfhe_ctx *ctx, int on);
/* FFT path, blind rotations of at most one ciphertext per SM (small batches, single-PBS latency); env
 * TFHE_B200_LATENCY_CFG=0|1|2|3|4.  A lone ciphertext on an SM is a chain of n dependent CMUX steps, each L levels + one inverse:
 *   4 (default)  like 3, with every CTA's multiply-accumulate and inverse transform split by key limb over twice the warps, and the
 *                partial results exchanged with st.async onto the destination CTA's mbarrier instead of a cluster barrier;
 *   3            batches that fit co-resident clusters (at most about sm_count / L ciphertexts): one thread-block CLUSTER of L CTAs per ciphertext, one gadget
 *                level per CTA (own accumulator replica, own key ring, own inverse transform; the exact u32 partial results
 *                are exchanged through distributed shared memory, one cluster barrier per step; kernels_fft_cluster.cuh;
 *                P0 and P1 shapes); larger batches fall through to mode 2, then 0;
 *   2            one ciphertext per CTA, ALL teams of the CTA on it: the L levels of a step (forward transforms and
 *                multiply-accumulates) are spread over the teams, partial sums are combined by the owner team
 *                (kernels_fft_latency.cuh; instantiated for the P0 and P1 shapes, otherwise mode 1);
 *   1            one ciphertext per CTA, one team, the idle shared memory holds a deep key ring (5-7 rows in flight);
 *   0            always the throughput configuration (2-4 ciphertexts per CTA, two-slot ring).
 * Same bits in every mode. */
int tfhe_ctx_set_latency_config(tfhe_ctx *ctx, int mode);
/* FFT path, throughput kernels of N = 512 (the reference's default set) and N = 1024: where the register passes of the transforms
 * exchange their data.  1 (default; env TFHE_B200_FFT_TMEM=0|1): tensor memory -- tcgen05.st / tcgen05.ld of different shapes transpose a
 * warp's registers on a data path of their own (fft_tmem.cuh); N = 512 moves every exchange and the published rows there, N = 1024 the
 * last three stages of every transform (one shared-memory exchange instead of two).  Keys uploaded while this is on carry a second
 * copy of the transformed BSK in that kernel's spectral order.  0: shared memory.  Switching it on only affects keys uploaded
 * afterwards.  Same bits. */
int tfhe_ctx_set_fft_exchange(tfhe_ctx *ctx, int tensor_memory);
int tfhe_fft_rounding_margin(tfhe_ctx *ctx, double *out);

/* Copies BSK+KSK (host or device pointers) to the ctx's device and transforms the BSK into the domain of the ctx's
 * arithmetic path (2-prime NTT: kernel K0; or limb-split FFT).  The caller keeps ownership of the inputs. */
int tfhe_bk_upload(tfhe_ctx *ctx, const uint32_t *bsk, const uint32_t *ksk, tfhe_bk **out);
/* BMMP variant (notes/BMMP Bootstrapping.md:13-25; the reference holds prose only, so parity with the
 * Rust crate is UNPINNED): blind rotation unrolled by two with key triples
 *   bk[3i] = GGSW(s_2i s_2i+1), bk[3i+1] = GGSW(s_2i (1 - s_2i+1)), bk[3i+2] = GGSW(s_2i+1 (1 - s_2i)),
 *   acc += ExtProd((X^(a+a') - 1) bk[3i] + (X^a - 1) bk[3i+1] + (X^a' - 1) bk[3i+2], acc).
 * tfhe_keygen_bmmp derives the same secret keys and KSK as tfhe_keygen from the seed; bsk3 = u32[3n/2][(k+1)l][k+1][N].
 * A key uploaded with tfhe_bk_upload_bmmp is used through the ordinary entry points (tfhe_bootstrap_batch,
 * tfhe_blind_rotate, tfhe_gate(s)_batch, ...); needs an even n and an FFT-path instantiation. */
int tfhe_keygen_bmmp(const tfhe_params *p, uint64_t seed, uint32_t *lwe_sk /* n */, uint32_t *glwe_sk /* k*N */,
                     uint32_t *bsk3, uint32_t *ksk);
int tfhe_bk_upload_bmmp(tfhe_ctx *ctx, const uint32_t *bsk3, const uint32_t *ksk, tfhe_bk **out);
/* Keys may be freed before or after their context: tfhe_ctx_destroy orphans the ctx's live keys (they can then only be
 * freed; every other use returns TFHE_E_PARAM). */
void tfhe_bk_free(tfhe_bk *bk);
/* Arithmetic path the key was transformed for (TFHE_PATH_*; BMMP keys: TFHE_PATH_FFT). */
int tfhe_bk_get_path(const tfhe_bk *bk);
/* Inspection (parity tests of the one-off key transform): size in bytes of the transformed BSK held on the device,
 * and a copy of it to host memory.  NTT path: u32[n][2][(k+1)l][k+1][N]; FFT path: f64 pairs[n][l][k+1 (slot d)][2][k+1 (c)][N/2],
 * diagonal-major: slot d of a level holds at column position c the polynomial (GGSW row of polynomial (c + d) mod (k+1),
 * column c) -- the order in which the blind rotation consumes it. */
size_t tfhe_bk_transformed_bytes(const tfhe_bk *bk);
int tfhe_bk_read_transformed(const tfhe_bk *bk, void *out, size_t bytes);

/* ------------------------------------------------------------------ the hot path */
/* bootstrapping.rs:58-120 `bootstrap`, batched.  luts = T unencoded test vectors [T][N] (values < 2^log_p,
 * else TFHE_E_ASSERT like glwe.rs:144); lut_idx[b] selects the test vector of ciphertext b (NULL: all 0);
 * an entry >= n_luts is never dereferenced: the call returns TFHE_E_PARAM (checked on the device, so lut_idx may be
 * a device pointer). */
int tfhe_bootstrap_batch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in /* [B][n+1] */,
                         const uint32_t *luts /* [T][N] */, size_t n_luts, const uint32_t *lut_idx /* [B] or NULL */,
                         size_t batch, uint32_t *lwe_out /* [B][n+1] */);
/* boolean.rs:9-53 `and`/`or` (+ XOR/NAND/NOR/XNOR), batched: ct_in = 2*ct1 + ct0, then bootstrap. */
int tfhe_gate_batch(tfhe_ctx *ctx, const tfhe_bk *bk, int gate, const uint32_t *ct0, const uint32_t *ct1, size_t batch,
                    uint32_t *out);
/* Same with one opcode per ciphertext (mixed-gate circuit level). gates[b] in tfhe_gate. */
int tfhe_gates_batch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint8_t *gates /* [B], host */, const uint32_t *ct0,
                     const uint32_t *ct1, size_t batch, uint32_t *out);

/* ------------------------------------------------------------------ multi-GPU: one process, all GPUs of the box */
/* Every ciphertext's bootstrap is independent (bootstrapping.rs:58-65), so a batch is split into contiguous, balanced index
 * ranges, one per GPU, with the keys REPLICATED on every GPU.  A tfhe_mgpu owns one tfhe_ctx per device and, for
 * device-resident batches, an NCCL communicator over them (libnccl.so.2 is opened at run time; without it the
 * device-pointer form returns TFHE_E_NCCL, the host-pointer form needs no NCCL):
 *   host pointers    each GPU copies its own shard in and out (cudaMemcpyAsync on its stream), all GPUs concurrently;
 *   device pointers  (memory of one of the tfhe_mgpu's devices, the "root") grouped ncclSend/ncclRecv scatter over NVLink,
 *                    local bootstraps, grouped ncclSend/ncclRecv gather into the root's output buffer.
 * There is no collective inside the data path.  Results are bit-identical to the single-GPU entry points. */
typedef struct tfhe_mgpu tfhe_mgpu;
typedef struct tfhe_mgpu_bk tfhe_mgpu_bk;
/* devices == NULL: devices 0..n_gpus-1. */
int tfhe_mgpu_create(const tfhe_params *p, int n_gpus, const int *devices, tfhe_mgpu **out);
void tfhe_mgpu_destroy(tfhe_mgpu *m);
int tfhe_mgpu_n_gpus(const tfhe_mgpu *m);
tfhe_ctx *tfhe_mgpu_ctx(tfhe_mgpu *m, int i);   /* the i-th device's context (path selection, timing, launch counts) */
const char *tfhe_mgpu_last_error(const tfhe_mgpu *m);
/* Replicates the key: uploaded and transformed on every device concurrently (bsk/ksk: host pointers). */
int tfhe_mgpu_bk_upload(tfhe_mgpu *m, const uint32_t *bsk, const uint32_t *ksk, tfhe_mgpu_bk **out);
int tfhe_mgpu_bk_upload_bmmp(tfhe_mgpu *m, const uint32_t *bsk3, const uint32_t *ksk, tfhe_mgpu_bk **out);
void tfhe_mgpu_bk_free(tfhe_mgpu_bk *bk);
/* tfhe_bootstrap_batch / tfhe_gates_batch over all GPUs.  lwe_in, lwe_out (ct0, ct1, out): all host or all on ONE
 * device of the tfhe_mgpu; luts, lut_idx, gates: host pointers. */
int tfhe_mgpu_bootstrap_batch(tfhe_mgpu *m, const tfhe_mgpu_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts,
                              const uint32_t *lut_idx, size_t batch, uint32_t *lwe_out);
int tfhe_mgpu_gates_batch(tfhe_mgpu *m, const tfhe_mgpu_bk *bk, const uint8_t *gates, const uint32_t *ct0, const uint32_t *ct1,
                          size_t batch, uint32_t *out);
/* Wall-clock split (ms) of the last tfhe_mgpu_* batch call: out[0] scatter, out[1] compute (slowest GPU), out[2] gather,
 * out[3] whole call. */
int tfhe_mgpu_last_timing(const tfhe_mgpu *m, double out[4]);

/* ------------------------------------------------------------------ compositions of the same kernels */
/* Key-switch-FIRST ordering (notes/TFHE.md:365-400): input and output live under the GLWE-derived LWE key
 * (dimension kN, lwe.rs:62-73 `LweSecretKey::from(&glwe_sk)`): key_switch_lwe -> mod switch + blind rotation ->
 * sample_extract.  Same kernels as tfhe_bootstrap_batch, composed in the other order. */
int tfhe_bootstrap_batch_ks_first(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in /* [B][kN+1] */,
                                  const uint32_t *luts /* [T][N] */, size_t n_luts, const uint32_t *lut_idx /* [B] or NULL */,
                                  size_t batch, uint32_t *lwe_out /* [B][kN+1] */);
/* k-input boolean gate (notes/Boolean Gates.md:9-11): ct_in = sum_i 2^i * cts[i] (Horner steps of lwe.rs:9-23 as in
 * boolean.rs:18), then one PBS with the LUT of f.  truth_table bit j = f(j) for j = sum_i 2^i * bit_i, j < 2^k.
 * Needs 2 <= k <= log_p (TFHE_E_PARAM otherwise).  f(0) = 1 is served as trivial(1) - gate(1 - f), because the
 * reference's LUT construction is only sound for f(0) = 0. */
int tfhe_gate_k_batch(tfhe_ctx *ctx, const tfhe_bk *bk, uint32_t k_inputs, uint32_t truth_table,
                      const uint32_t *const *cts /* k_inputs pointers to [B][n+1] */, size_t batch, uint32_t *out);

/* ------------------------------------------------------------------ flat wire / on-disk format */
/* The reference has no serialisation.  One little-endian file = header {magic "TFHEB200", u32 version = 1,
 * u32 kind, tfhe_params, u64 count: see tfhe_file_header, 80 bytes} + `count` u32 words in
 * exactly the layouts listed at the top of this header.  Lets an outside party with cargo exchange golden vectors. */
typedef enum tfhe_file_kind { TFHE_FILE_LWE_BATCH = 1, TFHE_FILE_GLWE_BATCH = 2, TFHE_FILE_BSK = 3, TFHE_FILE_KSK = 4,
                              TFHE_FILE_LWE_SK = 5, TFHE_FILE_GLWE_SK = 6, TFHE_FILE_TEST_VECTOR = 7 } tfhe_file_kind;
typedef struct tfhe_file_header {
    char magic[8];       /* "TFHEB200" */
    uint32_t version;    /* 1 */
    uint32_t kind;       /* tfhe_file_kind */
    tfhe_params params;  /* 56 bytes */
    uint64_t count;      /* number of u32 words that follow */
} tfhe_file_header;      /* 80 bytes */
int tfhe_file_write(const char *path, int kind, const tfhe_params *p, const uint32_t *words, uint64_t count);
/* Reads the header (always) and, when words != NULL, up to `capacity` words (TFHE_E_PARAM if the file holds more). */
int tfhe_file_read(const char *path, tfhe_file_header *hdr_out, uint32_t *words, uint64_t capacity);

/* ------------------------------------------------------------------ sub-operations (parity tests) */
/* utils.rs:23-33 switch_modulus(values, 32, log2(N)+1) */
int tfhe_switch_modulus(tfhe_ctx *ctx, const uint32_t *values, size_t len, uint32_t *out);
/* decomposer.rs:42-80 with the PBS (which=0) or KS (which=1) decomposer: out[len][levels], wrapped u32 digits */
int tfhe_decompose(tfhe_ctx *ctx, int which, const uint32_t *values, size_t len, uint32_t *out);
/* glwe.rs:20-34: out[b] = glwe[b] * X^{index[b]} (index as in Monomial, may be negative) */
int tfhe_glwe_mul_monomial(tfhe_ctx *ctx, const uint32_t *glwe /* [B][k+1][N] */, const int64_t *index /* [B], host */,
                           size_t batch, uint32_t *out);
/* utils.rs:155-160 poly_mul (the Toeplitz product) for a batch of pairs: out[b] = a[b] (*) g[b] in
 * Z_{2^32}[X]/(X^N+1), both operands any words mod 2^32 as in the reference (a is read as u32 when a coefficient is
 * outside [-1024, 1024]).  Small signed a (the digit range of the path) is one exact 2-prime transform product; otherwise the
 * four byte limbs of a are multiplied separately and recombined with shifts mod 2^32 (the same bits). */
int tfhe_negacyclic_mul(tfhe_ctx *ctx, const int32_t *a /* [B][N], host or device */, const uint32_t *g /* [B][N] */,
                        size_t batch, uint32_t *out /* [B][N] */);
/* ggsw.rs:132-161: out[b] = external_product(BSK[ggsw_index[b]], glwe[b]) */
int tfhe_external_product(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *ggsw_index /* [B], host */,
                          const uint32_t *glwe, size_t batch, uint32_t *out);
/* ggsw.rs:164-178: out[b] = cmux(BSK[ggsw_index[b]], ct0[b], ct1[b]); unlike the reference ct1 is NOT clobbered (H8) */
int tfhe_cmux(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *ggsw_index /* [B], host */, const uint32_t *ct0,
              const uint32_t *ct1, size_t batch, uint32_t *out);
/* bootstrapping.rs:67-105: mod switch, X^{-b} v(X), n CMUXes; returns the accumulator GLWEs [B][k+1][N] */
int tfhe_blind_rotate(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts,
                      const uint32_t *lut_idx, size_t batch, uint32_t *glwe_out);
/* bootstrapping.rs:122-156 with sample_index 0: [B][k+1][N] -> [B][kN+1] */
int tfhe_sample_extract(tfhe_ctx *ctx, const uint32_t *glwe, size_t batch, uint32_t *lwe_out);
/* key_switching.rs:63-103: [B][kN+1] -> [B][n+1] */
int tfhe_key_switch(tfhe_ctx *ctx, const tfhe_bk *bk, const uint32_t *lwe_in, size_t batch, uint32_t *lwe_out);
/* lwe.rs:9-23 as used by boolean.rs:18: out = 2*ct1 + ct0 */
int tfhe_gate_linear(tfhe_ctx *ctx, const uint32_t *ct0, const uint32_t *ct1, size_t batch, uint32_t *out);

/* ------------------------------------------------------------------ measurement helpers */
/* Measures the integer-pipe peaks of this device with dependent-free unrolled loops:
 * out[0] = 32-bit IMAD lane-ops/s, out[1] = IMAD.HI (mul.hi.u32), out[2] = IMAD.WIDE (mad.wide.u32),
 * out[3..5] = lazy Shoup NTT butterflies/s (IMAD.HI + 2 IMAD + 2 IADD3 each) issued from registers:
 * [3] one shared twiddle, [4] per-butterfly twiddle registers (as in the NTT passes), [5] like [4], q immediate;
 * out[6..7] = FMA-bound butterfly stream (16 values, 8 twiddle pairs in registers, no corrections) with 128-thread
 * CTAs at 3 and at 8 CTAs per SM -- the blind-rotation kernel's own residency and a saturating one. */
int tfhe_measure_int_peak(tfhe_ctx *ctx, double out[8]);
/* FP64 pipe of this device (the FFT path's arithmetic): out[0] = DFMA/s with two register operands + a constant (the
 * pipe's issue rate), out[1] = DFMA/s with three distinct register operands (as in the multiply-accumulate),
 * out[2] = forward FFT butterflies/s (6 DFMA each) issued from registers, out[3] reserved. */
int tfhe_measure_fp64_peak(tfhe_ctx *ctx, double out[4]);
/* Device time (ms, CUDA events on the ctx stream) of the last tfhe_bootstrap_batch / tfhe_gate(s)_batch:
 * out[0] blind rotation kernel, out[1] key-switch kernels, out[2] whole call incl. copies. */
int tfhe_last_timing(const tfhe_ctx *ctx, double out[3]);

#ifdef __cplusplus
}
#endif
#endif
