/*
 * tfhe_oracle.h -- CPU restatement of Janmajayamall/tfhe-research (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity ORACLE for the B200 path.  It is a line-by-line C restatement of the
 * reference's Rust crate (all citations are into /root/reference/src/).  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libtfhe_b200.so) never links or calls anything in oracle/.
 *
 * Parity pinning: the reference holds NO golden ciphertext vectors (every randomised test uses
 * thread_rng) and cargo/rustc are absent, so the oracle is pinned by
 *   (1) the reference's two deterministic tests (utils.rs:265-272 poly_mul_works,
 *       decomposer.rs:103-115 decomposition) re-run here,
 *   (2) the hand-derived KATs of SURVEY.md 9-C,
 *   (3) an independent numpy restatement (oracle/np_oracle.py) diffed against this file,
 *   (4) the reference's functional tests (encrypt/decrypt, key switch, bootstrap, AND gate).
 * Bit-level parity of `bootstrap` against the Rust binary itself is therefore "pinned by
 * construction" only.
 *
 * All arithmetic is u32 wrapping (release-build semantics of the reference).
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:23-34 TfheParams (+ decomposer.rs:2-6 DecomposerParams flattened) */
typedef struct orc_params {
    uint32_t glwe_dimension;   /* k */
    uint32_t glwe_poly_degree; /* log2 N  (the reference stores the LOG here, lib.rs:25,41) */
    uint32_t lwe_dimension;    /* n */
    uint32_t padding_bits;
    uint32_t log_p;
    uint32_t log_q;            /* always 32 */
    uint32_t ks_log_base, ks_levels;
    uint32_t pbs_log_base, pbs_levels;
    double lwe_std_dev, glwe_std_dev;
} orc_params;

/* lib.rs:76-124 impl Default (test_cfg!=0 -> cfg(test): n=4) */
void orc_params_default(int test_cfg, orc_params *out);

/* 1 = materialise the N x N Toeplitz matrix like utils.rs:113-160 (faithful, slow; used for the
 * CPU baseline), 0 = same sums without the temporary (identical results; used by tests). */
void orc_set_faithful_toeplitz(int on);

/* ---- utils.rs ---- */
uint32_t orc_integer_division(uint32_t a, uint32_t divisor);                       /* utils.rs:13-18 */
void orc_switch_modulus(const uint32_t *v, size_t len, uint32_t log_from, uint32_t log_to,
                        uint32_t *out);                                             /* utils.rs:23-33 */
uint32_t orc_f64_to_torus(double v);                                               /* utils.rs:36-41 */
void orc_teoplitz(const uint32_t *p, size_t n, uint32_t *matrix /* n*n */);        /* utils.rs:113-153 */
void orc_poly_mul(const uint32_t *p0, const uint32_t *p1, size_t n, uint32_t *res); /* utils.rs:155-160 */
void orc_poly_dot_product(const uint32_t *p0, const uint32_t *p1, size_t rows, size_t n,
                          size_t stride0, size_t stride1, uint32_t *res);          /* utils.rs:163-173 */
void orc_poly_mul_monomial(const uint32_t *p0, size_t n, int64_t monomial_index,
                           uint32_t *out);                                          /* utils.rs:183-207 */
void orc_school_book_negacyclic_mul(const uint32_t *p0, const uint32_t *p1, size_t n,
                                    uint32_t *res);                                 /* utils.rs:221-236 */

/* ---- decomposer.rs ---- */
uint32_t orc_round_value(uint32_t value, uint32_t log_base, uint32_t levels);      /* :27-40 */
void orc_decompose(uint32_t value, uint32_t log_base, uint32_t levels, uint32_t *out /* levels */); /* :42-80 */
uint32_t orc_recompose(const uint32_t *legs, uint32_t log_base, uint32_t levels);  /* :83-95 */

/* ---- glwe.rs ---- */
void orc_glwe_mul_monomial(const orc_params *p, const uint32_t *ct, int64_t index, uint32_t *out); /* :20-34 */
void orc_decompose_glwe_ciphertext(const orc_params *p, const uint32_t *ct,
                                   uint32_t *out /* (k+1)*l x N */);               /* :69-108 */
int  orc_glwe_encode_message(const orc_params *p, const uint32_t *msg, size_t len, uint32_t *out /* N */); /* :141-151 */
void orc_trivial_encrypt_glwe(const orc_params *p, const uint32_t *pt, uint32_t *ct); /* :232-243 */

/* ---- ggsw.rs ---- */
void orc_external_product(const orc_params *p, const uint32_t *ggsw, const uint32_t *glwe,
                          uint32_t *out);                                           /* :132-161 */
void orc_cmux(const orc_params *p, const uint32_t *ggsw, const uint32_t *ct0, uint32_t *ct1 /* clobbered */,
              uint32_t *out);                                                       /* :164-178 */

/* ---- bootstrapping.rs / key_switching.rs ---- */
void orc_sample_extract(const orc_params *p, const uint32_t *glwe, size_t sample_index,
                        uint32_t *lwe_out /* kN+1 */);                              /* bootstrapping.rs:122-156 */
void orc_key_switch_lwe(const orc_params *p, const uint32_t *lwe_in /* kN+1 */, const uint32_t *ksk,
                        uint32_t *lwe_out /* n+1 */);                               /* key_switching.rs:63-103 */
/* blind rotation only (bootstrapping.rs:67-105): returns the accumulator GLWE [(k+1), N] */
int  orc_blind_rotate(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk,
                      const uint32_t *test_vector /* N, unencoded */, uint32_t *acc_out);
int  orc_bootstrap(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk, const uint32_t *ksk,
                   const uint32_t *test_vector /* N, unencoded */, uint32_t *lwe_out); /* bootstrapping.rs:58-120 */
/* B independent bootstraps, one per worker (pthread) (the reference itself is single-threaded). */
int  orc_bootstrap_batch(const orc_params *p, const uint32_t *lwe_in, size_t batch, const uint32_t *bsk,
                         const uint32_t *ksk, const uint32_t *test_vector, uint32_t *lwe_out, int nthreads);

/* ---- test_vector.rs ---- */
int  orc_test_vector_from_lut(const orc_params *p, const uint32_t *lut, size_t lut_len, uint32_t *tv); /* :38-67 */
void orc_test_vector_identity(const orc_params *p, uint32_t *tv);                  /* :23-35 */
/* op: 0 AND, 1 OR, 2 XOR (f(0,0)=0 gates only) -- :5-20 */
int  orc_test_vector_boolean(const orc_params *p, int op, uint32_t *tv);

/* ---- lwe.rs ---- */
int      orc_lwe_encode(const orc_params *p, uint32_t m, uint32_t *out);           /* :83-88 */
uint32_t orc_lwe_decode(const orc_params *p, uint32_t pt);                         /* :102-107 (floor) */
uint32_t orc_lwe_decrypt(const uint32_t *sk, size_t n, const uint32_t *ct);        /* :162-173 */
void     orc_lwe_add(const uint32_t *a, const uint32_t *b, size_t len, uint32_t *out); /* :9-15 */
void     orc_lwe_mul_scalar(const uint32_t *a, uint32_t s, size_t len, uint32_t *out); /* :17-23 */

/* ---- boolean.rs :9-53 (+ gates defined by this build) ----
 * op: 0 AND, 1 OR, 2 XOR, 3 NAND (= trivial(1) - AND), 4 NOR, 5 XNOR */
int  orc_gate(const orc_params *p, int op, const uint32_t *ct0, const uint32_t *ct1, const uint32_t *bsk,
              const uint32_t *ksk, uint32_t *out);

/* ---- seeded key generation / encryption (own RNG; thread_rng is not reproducible) ----
 * RNG spec (shared with the product's host keygen so that keys can be diffed bit for bit):
 * xoshiro256** seeded by splitmix64(seed ^ domain*0x9E3779B97F4A7C15 ^ index*0xD1B54A32D192ED03).
 * domains: 1 GGSW i, 2 KSK block s_index, 3 lwe_sk, 4 glwe_sk, 5 client encryption index. */
void orc_keygen(const orc_params *p, uint64_t seed, uint32_t *lwe_sk /* n */, uint32_t *glwe_sk /* k*N */,
                uint32_t *bsk /* n*(k+1)l*(k+1)*N */, uint32_t *ksk /* kN*l_ks*(n+1) */);
void orc_lwe_encrypt(const orc_params *p, const uint32_t *sk, size_t n, uint32_t plaintext, uint64_t seed,
                     uint64_t index, uint32_t *ct /* n+1 */);                       /* lwe.rs:138-160 */

#ifdef __cplusplus
}
#endif
#endif
