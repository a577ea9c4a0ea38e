// host_api.cpp -- host-side (C++) mirror of the reference's non-hot-path API: parameters
// (lib.rs:23-124), test-vector construction (test_vector.rs:5-67), client-side LWE encode/decode/
// encrypt/decrypt (lwe.rs:83-173) and key generation (bootstrapping.rs:23-56, ggsw.rs:76-130,
// key_switching.rs:20-60).  No CUDA here; no dependency on oracle/.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "api_internal.hpp"

namespace tfhe_host {

// ---- seeded RNG: xoshiro256** seeded by splitmix64(seed ^ domain*0x9E3779B97F4A7C15 ^ index*0xD1B54A32D192ED03);
// domains: 1 GGSW i, 2 KSK block s_index, 3 lwe_sk, 4 glwe_sk, 5 client encryption index, 6 GGSW g of a BMMP key triple.
// This seeded generator is a TEST HARNESS (reproducible keys and inputs for parity tests and benchmarks; the reference
// itself uses thread_rng), NOT a CSPRNG: do not generate production keys with it. ----
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t domain, uint64_t index) {
        uint64_t x = seed ^ (domain * 0x9E3779B97F4A7C15ULL) ^ (index * 0xD1B54A32D192ED03ULL);
        for (auto &v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t u64() {  // xoshiro256**
        const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t u32() { return (uint32_t)(u64() >> 32); }
    double gauss(double std_dev) {  // Box-Muller, cosine branch
        const double u1 = (double)((u64() >> 11) + 1) * (1.0 / 9007199254740992.0);
        const double u2 = (double)(u64() >> 11) * (1.0 / 9007199254740992.0);
        return std_dev * sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
    }
    // utils.rs:68-93: bits of successive random bytes, least significant first
    void binary(uint32_t *out, size_t len) {
        uint8_t cur = (uint8_t)u32();
        int bit = 0;
        for (size_t i = 0; i < len; i++) {
            out[i] = (cur >> bit) & 1u;
            if (++bit == 8) { cur = (uint8_t)u32(); bit = 0; }
        }
    }
};

// utils.rs:36-41: fractional part to the torus; Rust's `as u32` saturates, negatives become 0 (H4)
static uint32_t f64_to_torus(double v) {
    double frac = v - round(v);
    frac = round(frac * 4294967296.0);
    if (!(frac > 0.0)) return 0;
    if (frac >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)frac;
}

// a (*) s in Z_{2^32}[X]/(X^N+1) for a BINARY s: sum of negacyclic rotations of a (utils.rs:155-173
// gives the same value; this form needs no multiplications).
static void negacyclic_mul_binary(const uint32_t *a, const uint32_t *s, size_t N, uint32_t *acc /* += */) {
    for (size_t j = 0; j < N; j++) {
        if (!s[j]) continue;
        // res[i] += a[i-j] for i >= j ; res[i] -= a[N+i-j] for i < j
        for (size_t i = j; i < N; i++) acc[i] += a[i - j];
        for (size_t i = 0; i < j; i++) acc[i] -= a[N + i - j];
    }
}

// glwe.rs:190-209: uniform masks, body = sum a_i*s_i + e
static void encrypt_glwe_zero(const tfhe_params &p, const uint32_t *glwe_sk, Rng &r, uint32_t *ct) {
    const size_t N = (size_t)1 << p.glwe_poly_degree, k = p.glwe_dimension;
    for (size_t i = 0; i < k * N; i++) ct[i] = r.u32();
    uint32_t *body = ct + k * N;
    memset(body, 0, N * sizeof(uint32_t));
    for (size_t i = 0; i < k; i++) negacyclic_mul_binary(ct + i * N, glwe_sk + i * N, N, body);
    for (size_t j = 0; j < N; j++) body[j] += f64_to_torus(r.gauss(p.glwe_std_dev));
}

// lwe.rs:117-136: error first, then the mask
static void encrypt_lwe_zero(const uint32_t *sk, size_t n, double std_dev, Rng &r, uint32_t *ct) {
    const uint32_t error = f64_to_torus(r.gauss(std_dev));
    uint32_t a_s = 0;
    for (size_t i = 0; i < n; i++) {
        ct[i] = r.u32();
        a_s += sk[i] * ct[i];
    }
    ct[n] = a_s + error;
}

// ggsw.rs:76-130: GGSW encryption of the plaintext m (a small integer): (k+1)*l GLWE encryptions of zero, plus
// m * B^(l_full - (lev+1)) on coefficient 0 of polynomial `poly` of row poly*l + lev (ggsw.rs:96-103)
static void encrypt_ggsw(const tfhe_params &p, const uint32_t *glwe_sk, uint32_t m, Rng &r, uint32_t *g) {
    const size_t N = (size_t)1 << p.glwe_poly_degree, k = p.glwe_dimension, l = p.pbs_levels, glwe_sz = (k + 1) * N;
    const uint32_t lq = p.log_q / p.pbs_log_base;
    for (size_t poly = 0; poly < k + 1; poly++)
        for (size_t lev = 0; lev < l; lev++) {
            uint32_t *row = g + (poly * l + lev) * glwe_sz;
            encrypt_glwe_zero(p, glwe_sk, r, row);
            if (m != 0) row[poly * N] += m * (1u << (p.pbs_log_base * (lq - (lev + 1))));
        }
}
static unsigned keygen_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return hw ? (hw > 32 ? 32 : hw) : 4;
}
template <class F>
static void parallel_for(size_t count, F f) {
    std::atomic<size_t> next{0};
    auto work = [&]() { for (size_t i; (i = next.fetch_add(1)) < count;) f(i); };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < keygen_threads(); t++) th.emplace_back(work);
    for (auto &t : th) t.join();
}
static void gen_secret_keys(const tfhe_params &p, uint64_t seed, uint32_t *lwe_sk, uint32_t *glwe_sk) {
    const size_t N = (size_t)1 << p.glwe_poly_degree;
    { Rng r(seed, 3, 0); r.binary(lwe_sk, p.lwe_dimension); }            // lwe.rs:54-58
    { Rng r(seed, 4, 0); r.binary(glwe_sk, p.glwe_dimension * N); }      // glwe.rs:177-181
}
// bootstrapping.rs:41-51 + key_switching.rs:20-60: KSK from the flattened GLWE key to the LWE key
static void gen_ksk(const tfhe_params &p, uint64_t seed, const uint32_t *lwe_sk, const uint32_t *glwe_sk, uint32_t *ksk) {
    const size_t N = (size_t)1 << p.glwe_poly_degree, n = p.lwe_dimension, from_n = p.glwe_dimension * N, lks = p.ks_levels;
    const uint32_t lfull = p.log_q / p.ks_log_base;
    parallel_for(from_n, [&](size_t s) {
        Rng r(seed, 2, s);
        for (size_t lev = 0; lev < lks; lev++) {
            uint32_t *row = ksk + (s * lks + lev) * (n + 1);
            encrypt_lwe_zero(lwe_sk, n, p.lwe_std_dev, r, row);
            row[n] += (1u << (p.ks_log_base * (lfull - (lev + 1)))) * glwe_sk[s];  // key_switching.rs:41-48
        }
    });
}

}  // namespace tfhe_host

using namespace tfhe_host;

extern "C" {

int tfhe_params_default(int test_cfg, tfhe_params *o) {
    if (!o) return TFHE_E_PARAM;
    o->glwe_dimension = 2;
    o->glwe_poly_degree = 9;
    o->lwe_dimension = test_cfg ? 4 : 722;
    o->padding_bits = 1;
    o->log_p = 2;
    o->log_q = 32;
    o->ks_log_base = 4;
    o->ks_levels = 5;
    o->pbs_log_base = 4;
    o->pbs_levels = 6;
    o->lwe_std_dev = 0.000013071021089943935;
    o->glwe_std_dev = 0.00000004990272175010415;
    return TFHE_OK;
}

int tfhe_params_preset(const char *name, tfhe_params *o) {
    if (!name || !o) return TFHE_E_PARAM;
    if (!strcmp(name, "P0")) return tfhe_params_default(0, o);
    if (!strcmp(name, "P0t")) return tfhe_params_default(1, o);
    tfhe_params_default(0, o);
    if (!strcmp(name, "P1")) {  // BASELINE config #2: N=1024, n=630
        o->glwe_dimension = 1; o->glwe_poly_degree = 10; o->lwe_dimension = 630;
        o->pbs_log_base = 8; o->pbs_levels = 3; o->ks_log_base = 2; o->ks_levels = 8;
        o->log_p = 2; o->padding_bits = 1;
        o->lwe_std_dev = 3.0517578125e-05;       /* 2^-15 */
        o->glwe_std_dev = 2.9802322387695312e-08; /* 2^-25 */
        return TFHE_OK;
    }
    if (!strcmp(name, "P2")) {  // BASELINE config #3: N=2048, 4-bit message space
        o->glwe_dimension = 1; o->glwe_poly_degree = 11; o->lwe_dimension = 742;
        o->pbs_log_base = 8; o->pbs_levels = 3; o->ks_log_base = 4; o->ks_levels = 5;
        o->log_p = 4; o->padding_bits = 1;
        o->lwe_std_dev = 7.069849454709433e-06;
        o->glwe_std_dev = 4.656612873077393e-10;  /* 2^-31 */
        return TFHE_OK;
    }
    return TFHE_E_PARAM;
}

int tfhe_params_validate(const tfhe_params *p) {
    if (!p) return TFHE_E_PARAM;
    if (p->log_q != 32) return TFHE_E_PARAM;
    for (int which = 0; which < 2; which++) {
        const uint32_t lb = which ? p->ks_log_base : p->pbs_log_base, lv = which ? p->ks_levels : p->pbs_levels;
        if (lb == 0 || lv == 0 || 32 % lb != 0 || lb * lv > 32) return TFHE_E_PARAM;
    }
    if (p->glwe_poly_degree > 11 || p->log_p + p->padding_bits >= 32) return TFHE_E_PARAM;
    if (p->log_p > p->glwe_poly_degree) return TFHE_E_PARAM;  // 2^log_p must divide N (test_vector.rs:47)
    if (p->lwe_dimension == 0 || p->lwe_dimension > 65534) return TFHE_E_PARAM;
    if (tfhe_host::pbs_config_id(*p) < 0 || tfhe_host::ks_config_id(*p) < 0) return TFHE_E_PARAM;
    return TFHE_OK;
}

// test_vector.rs:38-67
int tfhe_test_vector_from_lut(const tfhe_params *p, const uint32_t *lut, size_t lut_len, uint32_t *tv) {
    if (!p || !lut || !tv) return TFHE_E_PARAM;
    if (p->glwe_poly_degree > 31 || p->log_p > p->glwe_poly_degree) return TFHE_E_PARAM;   // 2^log_p must divide N (rep = N / 2^log_p >= 1)
    const uint32_t pm = 1u << p->log_p;
    if (lut_len != pm) return TFHE_E_ASSERT;  // assert! test_vector.rs:41
    const size_t N = (size_t)1 << p->glwe_poly_degree;
    const size_t rep = N / pm, half = rep / 2;
    // value at position x of the un-rotated vector: lut[x / rep], "negated" inside the message bits
    // for the first rep/2 entries; then rotate left by rep/2
    for (size_t i = 0; i < N; i++) {
        const size_t x = (i + half) % N;
        uint32_t v = lut[x / rep];
        if (x < half && v != 0) v = pm - v;
        tv[i] = v;
    }
    return TFHE_OK;
}

int tfhe_test_vector_identity(const tfhe_params *p, uint32_t *tv) {
    if (!p || p->glwe_poly_degree > 31 || p->log_p > p->glwe_poly_degree) return TFHE_E_PARAM;
    std::vector<uint32_t> lut(1u << p->log_p);
    for (uint32_t i = 0; i < lut.size(); i++) lut[i] = i;
    return tfhe_test_vector_from_lut(p, lut.data(), lut.size(), tv);
}

int tfhe_test_vector_boolean(const tfhe_params *p, int gate, uint32_t *tv) {
    if (!p || gate < TFHE_AND || gate > TFHE_XOR) return TFHE_E_PARAM;
    if (p->glwe_poly_degree > 31 || p->log_p > p->glwe_poly_degree) return TFHE_E_PARAM;
    std::vector<uint32_t> lut(1u << p->log_p);
    for (uint32_t i = 0; i < lut.size(); i++) {  // test_vector.rs:14-17: left = bit 1, right = bit 0
        const uint32_t l = (i >> 1) & 1u, r = i & 1u;
        lut[i] = gate == TFHE_AND ? (l & r) : gate == TFHE_OR ? (l | r) : (l ^ r);
    }
    return tfhe_test_vector_from_lut(p, lut.data(), lut.size(), tv);
}

int tfhe_lwe_encode(const tfhe_params *p, uint32_t m, uint32_t *out) {
    if (!p || !out || p->log_p + p->padding_bits > p->log_q || p->log_q > 32 || p->log_p > 31) return TFHE_E_PARAM;
    if (!(m < (1u << p->log_p))) return TFHE_E_ASSERT;  // lwe.rs:84
    *out = m << (p->log_q - (p->log_p + p->padding_bits));
    return TFHE_OK;
}
int tfhe_lwe_decode(const tfhe_params *p, uint32_t pt, uint32_t *out) {
    if (!p || !out) return TFHE_E_PARAM;
    *out = pt >> (p->log_q - (p->log_p + p->padding_bits));  // lwe.rs:104-105: floor, no mask (H5)
    return TFHE_OK;
}
int tfhe_lwe_encrypt(const tfhe_params *p, const uint32_t *sk, size_t n, uint32_t plaintext, uint64_t seed, uint64_t index,
                     uint32_t *ct) {
    if (!p || !sk || !ct) return TFHE_E_PARAM;
    Rng r(seed, 5, index);
    encrypt_lwe_zero(sk, n, p->lwe_std_dev, r, ct);
    ct[n] += plaintext;  // lwe.rs:151-152
    return TFHE_OK;
}
int tfhe_lwe_decrypt(const uint32_t *sk, size_t n, const uint32_t *ct, uint32_t *out) {
    if (!sk || !ct || !out) return TFHE_E_PARAM;
    uint32_t a_s = 0;
    for (size_t i = 0; i < n; i++) a_s += sk[i] * ct[i];
    *out = ct[n] - a_s;  // lwe.rs:169-170
    return TFHE_OK;
}

int tfhe_keygen(const tfhe_params *pp, uint64_t seed, uint32_t *lwe_sk, uint32_t *glwe_sk, uint32_t *bsk, uint32_t *ksk) {
    if (!pp || !lwe_sk || !glwe_sk || !bsk || !ksk) return TFHE_E_PARAM;
    const tfhe_params p = *pp;
    if (p.log_q != 32 || p.pbs_log_base == 0 || p.ks_log_base == 0) return TFHE_E_PARAM;
    const size_t N = (size_t)1 << p.glwe_poly_degree, k = p.glwe_dimension, n = p.lwe_dimension, l = p.pbs_levels;
    const size_t ggsw_sz = (k + 1) * l * (k + 1) * N;
    gen_secret_keys(p, seed, lwe_sk, glwe_sk);
    // bootstrapping.rs:32-38: GGSW(s_i) for every LWE secret bit
    parallel_for(n, [&](size_t i) {
        Rng r(seed, 1, i);
        encrypt_ggsw(p, glwe_sk, lwe_sk[i], r, bsk + i * ggsw_sz);
    });
    gen_ksk(p, seed, lwe_sk, glwe_sk, ksk);
    return TFHE_OK;
}

// notes/BMMP Bootstrapping.md:21-25: the unrolled-by-2 bootstrapping key, 3 GGSWs per pair of LWE secret bits:
//   bk[3i] = GGSW(s_2i * s_2i+1), bk[3i+1] = GGSW(s_2i * (1 - s_2i+1)), bk[3i+2] = GGSW(s_2i+1 * (1 - s_2i)).
// Secret keys and KSK are the ones tfhe_keygen derives from the same seed.
int tfhe_keygen_bmmp(const tfhe_params *pp, uint64_t seed, uint32_t *lwe_sk, uint32_t *glwe_sk, uint32_t *bsk3, uint32_t *ksk) {
    if (!pp || !lwe_sk || !glwe_sk || !bsk3 || !ksk) return TFHE_E_PARAM;
    const tfhe_params p = *pp;
    if (p.log_q != 32 || p.pbs_log_base == 0 || p.ks_log_base == 0 || (p.lwe_dimension & 1u)) return TFHE_E_PARAM;
    const size_t N = (size_t)1 << p.glwe_poly_degree, k = p.glwe_dimension, n = p.lwe_dimension, l = p.pbs_levels;
    const size_t ggsw_sz = (k + 1) * l * (k + 1) * N;
    gen_secret_keys(p, seed, lwe_sk, glwe_sk);
    parallel_for(3 * (n / 2), [&](size_t g) {
        const size_t i = g / 3, which = g % 3;
        const uint32_t s0 = lwe_sk[2 * i], s1 = lwe_sk[2 * i + 1];
        const uint32_t m = which == 0 ? s0 * s1 : which == 1 ? s0 * (1u - s1) : s1 * (1u - s0);
        Rng r(seed, 6, g);   // own domain: domain 5 is the client-encryption stream (a shared stream would leak the noise)
        encrypt_ggsw(p, glwe_sk, m, r, bsk3 + g * ggsw_sz);
    });
    gen_ksk(p, seed, lwe_sk, glwe_sk, ksk);
    return TFHE_OK;
}

// ---- flat wire / on-disk format: header + little-endian u32 words
int tfhe_file_write(const char *path, int kind, const tfhe_params *p, const uint32_t *words, uint64_t count) {
    if (!path || !p || (!words && count)) return TFHE_E_PARAM;
    static_assert(sizeof(tfhe_file_header) == 80, "wire header layout");
    tfhe_file_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "TFHEB200", 8);
    h.version = 1;
    h.kind = (uint32_t)kind;
    h.params = *p;
    h.count = count;
    FILE *f = fopen(path, "wb");
    if (!f) return TFHE_E_PARAM;
    const bool ok = fwrite(&h, sizeof h, 1, f) == 1 && (count == 0 || fwrite(words, 4, count, f) == count);
    return (fclose(f) == 0 && ok) ? TFHE_OK : TFHE_E_PARAM;
}
int tfhe_file_read(const char *path, tfhe_file_header *hdr_out, uint32_t *words, uint64_t capacity) {
    if (!path || !hdr_out) return TFHE_E_PARAM;
    FILE *f = fopen(path, "rb");
    if (!f) return TFHE_E_PARAM;
    tfhe_file_header h;
    int rc = TFHE_OK;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "TFHEB200", 8) != 0 || h.version != 1) rc = TFHE_E_PARAM;
    if (rc == TFHE_OK) {
        *hdr_out = h;
        if (words) {
            if (h.count > capacity) rc = TFHE_E_PARAM;
            else if (h.count && fread(words, 4, h.count, f) != h.count) rc = TFHE_E_PARAM;
        }
    }
    fclose(f);
    return rc;
}

}  // extern "C"
