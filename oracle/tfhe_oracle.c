/*
 * tfhe_oracle.c -- CPU restatement of Janmajayamall/tfhe-research.  TEST INFRASTRUCTURE ONLY
 * (see tfhe_oracle.h for the rules and for how parity is pinned).  Every function cites the
 * reference lines it follows (paths relative to /root/reference/src/).  u32 wrapping everywhere.
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static int g_faithful_toeplitz = 0;
void orc_set_faithful_toeplitz(int on) { g_faithful_toeplitz = on; }

/* ------------------------------------------------------------------ params: lib.rs:76-124 */
void orc_params_default(int test_cfg, orc_params *o) {
    o->glwe_dimension = 2;
    o->glwe_poly_degree = 9;
    o->lwe_dimension = test_cfg ? 4 : 722; /* lib.rs:82 vs :106 */
    o->log_p = 2;
    o->log_q = 32;
    o->ks_log_base = 4;
    o->ks_levels = 5;
    o->pbs_log_base = 4;
    o->pbs_levels = 6;
    o->padding_bits = 1;
    o->lwe_std_dev = 0.000013071021089943935;
    o->glwe_std_dev = 0.00000004990272175010415;
}

static inline size_t P_N(const orc_params *p) { return (size_t)1 << p->glwe_poly_degree; }
static inline size_t P_K(const orc_params *p) { return p->glwe_dimension; }

/* ------------------------------------------------------------------ utils.rs */
/* utils.rs:13-18 */
uint32_t orc_integer_division(uint32_t a, uint32_t divisor) {
    uint32_t rational = a / divisor;
    uint32_t fractional = a % divisor;
    return rational + ((fractional + (divisor >> 1)) / divisor);
}

/* utils.rs:23-33 */
void orc_switch_modulus(const uint32_t *v, size_t len, uint32_t log_from, uint32_t log_to, uint32_t *out) {
    for (size_t i = 0; i < len; i++) {
        uint32_t x = orc_integer_division(v[i], 1u << (log_from - log_to));
        out[i] = x % (1u << log_to);
    }
}

/* utils.rs:36-41; `frac as u32` is a saturating cast in Rust (negatives -> 0): SURVEY 9-B H4 */
uint32_t orc_f64_to_torus(double v) {
    double frac = v - round(v);
    frac *= 4294967296.0;
    frac = round(frac);
    if (!(frac > 0.0)) return 0; /* negative, -0.0 and NaN all cast to 0 */
    if (frac >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)frac;
}

/* utils.rs:113-153: row i = [p[i], p[i-1], .., p[0], -p[n-1], .., -p[i+1]] */
void orc_teoplitz(const uint32_t *p, size_t n, uint32_t *m) {
    for (size_t i = 0; i < n; i++) {
        uint32_t *row = m + i * n;
        size_t c = 0;
        for (size_t j = i + 1; j-- > 0;) row[c++] = p[j];
        for (size_t j = n; j-- > i + 1;) row[c++] = 0u - p[j];
    }
}

/* utils.rs:155-160: res = teoplitz(p0) . p1 (ndarray dot on u32 = wrapping in release) */
void orc_poly_mul(const uint32_t *p0, const uint32_t *p1, size_t n, uint32_t *res) {
    if (g_faithful_toeplitz) {
        uint32_t *m = (uint32_t *)malloc(n * n * sizeof(uint32_t));
        orc_teoplitz(p0, n, m);
        for (size_t i = 0; i < n; i++) {
            const uint32_t *row = m + i * n;
            uint32_t s = 0;
            for (size_t j = 0; j < n; j++) s += row[j] * p1[j];
            res[i] = s;
        }
        free(m);
        return;
    }
    /* same sums, no temporary: res[i] = sum_{j<=i} p0[i-j] p1[j] - sum_{j>i} p0[n+i-j] p1[j] */
    for (size_t i = 0; i < n; i++) {
        uint32_t s = 0;
        for (size_t j = 0; j <= i; j++) s += p0[i - j] * p1[j];
        for (size_t j = i + 1; j < n; j++) s -= p0[n + i - j] * p1[j];
        res[i] = s;
    }
}

/* utils.rs:163-173 (row r of p0 at p0 + r*stride0, of p1 at p1 + r*stride1) */
void orc_poly_dot_product(const uint32_t *p0, const uint32_t *p1, size_t rows, size_t n, size_t stride0,
                          size_t stride1, uint32_t *res) {
    uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
    orc_poly_mul(p0, p1, n, res);
    for (size_t r = 1; r < rows; r++) {
        orc_poly_mul(p0 + r * stride0, p1 + r * stride1, n, tmp);
        for (size_t i = 0; i < n; i++) res[i] += tmp[i];
    }
    free(tmp);
}

/* utils.rs:183-207 */
void orc_poly_mul_monomial(const uint32_t *p0, size_t n, int64_t monomial_index, uint32_t *out) {
    size_t idx = (size_t)(((uint64_t)monomial_index) % (uint64_t)(2 * n)); /* `as usize % (2*n)` :186 */
    size_t flip_sign = idx / n;
    size_t degree = idx % n;
    uint32_t mul = flip_sign ? 0xFFFFFFFFu : 1u; /* (u32::MAX).pow(flip_sign) :195 */
    /* rotate_right(degree): new[i] = old[(i - degree) mod n] :199 */
    for (size_t i = 0; i < n; i++) {
        size_t src = (i + n - degree) % n;
        out[i] = p0[src] * mul;
    }
    for (size_t i = 0; i < degree; i++) out[i] = 0u - out[i]; /* :202-204 */
}

/* utils.rs:221-236 */
void orc_school_book_negacyclic_mul(const uint32_t *p0, const uint32_t *p1, size_t n, uint32_t *res) {
    for (size_t i = 0; i < n; i++) {
        uint32_t s = 0;
        for (size_t j = 0; j < i + 1; j++) s += p0[j] * p1[i - j];
        for (size_t j = i + 1; j < n; j++) s -= p0[j] * p1[n - (j - i)];
        res[i] = s;
    }
}

/* ------------------------------------------------------------------ decomposer.rs */
/* decomposer.rs:27-40 */
uint32_t orc_round_value(uint32_t value, uint32_t log_base, uint32_t levels) {
    uint32_t ignored_bits = 32 - log_base * levels;
    if (ignored_bits == 0) return value;
    uint32_t ignored_mask = (1u << ignored_bits) - 1;
    uint32_t ignored_value = value & ignored_mask;
    uint32_t ignored_msb = ignored_value >> (ignored_bits - 1);
    return ((value >> ignored_bits) + ignored_msb) << ignored_bits; /* wraps to 0 at the top (H11) */
}

/* decomposer.rs:42-80 -- digit set {-B/2..B/2-1} u {B} */
void orc_decompose(uint32_t value, uint32_t log_base, uint32_t levels, uint32_t *out) {
    uint32_t tmp[32];
    uint32_t nd = 32 / log_base;
    value = orc_round_value(value, log_base, levels);
    uint32_t base_mask = (1u << log_base) - 1;
    uint32_t base_by_2_mask = 1u << (log_base - 1);
    uint32_t carry = 0;
    for (uint32_t l = 0; l < nd; l++) {
        uint32_t res = ((value >> (log_base * l)) & base_mask) + carry;
        uint32_t carry_mask = res & base_by_2_mask;
        res = res - (carry_mask << 1);
        carry = carry_mask >> (log_base - 1);
        tmp[l] = res;
    }
    /* reverse to big endian, keep the first `levels` (:67-77) */
    for (uint32_t i = 0; i < levels; i++) out[i] = tmp[nd - 1 - i];
}

/* decomposer.rs:83-95 */
uint32_t orc_recompose(const uint32_t *legs, uint32_t log_base, uint32_t levels) {
    uint32_t value = 0;
    for (uint32_t i = 0; i < levels; i++) value += legs[i] << (log_base * (levels - 1 - i));
    uint32_t ignored_bits = 32 - log_base * levels;
    return ignored_bits >= 32 ? 0 : value << ignored_bits;
}

/* ------------------------------------------------------------------ glwe.rs */
/* glwe.rs:20-34 */
void orc_glwe_mul_monomial(const orc_params *p, const uint32_t *ct, int64_t index, uint32_t *out) {
    size_t N = P_N(p);
    for (size_t r = 0; r < P_K(p) + 1; r++) orc_poly_mul_monomial(ct + r * N, N, index, out + r * N);
}

/* glwe.rs:69-108: out row (poly*l + level), level 0 most significant */
void orc_decompose_glwe_ciphertext(const orc_params *p, const uint32_t *ct, uint32_t *out) {
    size_t N = P_N(p), l = p->pbs_levels;
    uint32_t d[32];
    for (size_t r = 0; r < P_K(p) + 1; r++)
        for (size_t j = 0; j < N; j++) {
            orc_decompose(ct[r * N + j], p->pbs_log_base, p->pbs_levels, d);
            for (size_t lev = 0; lev < l; lev++) out[(r * l + lev) * N + j] = d[lev];
        }
}

/* glwe.rs:141-151 (assert -> -1) */
int orc_glwe_encode_message(const orc_params *p, const uint32_t *msg, size_t len, uint32_t *out) {
    size_t N = P_N(p);
    memset(out, 0, N * sizeof(uint32_t));
    for (size_t i = 0; i < len && i < N; i++) {
        if (!(msg[i] < (1u << p->log_p))) return -1;
        out[i] = msg[i] << (p->log_q - (p->log_p + p->padding_bits));
    }
    return 0;
}

/* glwe.rs:232-243 */
void orc_trivial_encrypt_glwe(const orc_params *p, const uint32_t *pt, uint32_t *ct) {
    size_t N = P_N(p), k = P_K(p);
    memset(ct, 0, (k + 1) * N * sizeof(uint32_t));
    memcpy(ct + k * N, pt, N * sizeof(uint32_t));
}

/* ------------------------------------------------------------------ ggsw.rs */
/* ggsw.rs:132-161 */
void orc_external_product(const orc_params *p, const uint32_t *ggsw, const uint32_t *glwe, uint32_t *out) {
    size_t N = P_N(p), k = P_K(p), l = p->pbs_levels, rows = (k + 1) * l;
    uint32_t *dec = (uint32_t *)malloc(rows * N * sizeof(uint32_t));
    orc_decompose_glwe_ciphertext(p, glwe, dec);
    for (size_t col = 0; col < k + 1; col++) /* ggsw[:, col, :] : row stride (k+1)*N */
        orc_poly_dot_product(dec, ggsw + col * N, rows, N, N, (k + 1) * N, out + col * N);
    free(dec);
}

/* ggsw.rs:164-178 (ct1 is mutated in place, H8) */
void orc_cmux(const orc_params *p, const uint32_t *ggsw, const uint32_t *ct0, uint32_t *ct1, uint32_t *out) {
    size_t sz = (P_K(p) + 1) * P_N(p);
    for (size_t i = 0; i < sz; i++) ct1[i] -= ct0[i];
    orc_external_product(p, ggsw, ct1, out);
    for (size_t i = 0; i < sz; i++) out[i] += ct0[i];
}

/* ------------------------------------------------------------------ bootstrapping.rs */
/* bootstrapping.rs:122-156 */
void orc_sample_extract(const orc_params *p, const uint32_t *glwe, size_t sample_index, uint32_t *lwe_out) {
    size_t N = P_N(p), k = P_K(p), c = 0;
    uint32_t lwe_b = glwe[k * N + sample_index];
    for (size_t r = 0; r < k; r++) {
        const uint32_t *poly = glwe + r * N;
        for (size_t i = sample_index + 1; i-- > 0;) lwe_out[c++] = poly[i];
        for (size_t i = N; i-- > sample_index + 1;) lwe_out[c++] = 0u - poly[i];
    }
    lwe_out[c] = lwe_b;
}

/* key_switching.rs:63-103 */
void orc_key_switch_lwe(const orc_params *p, const uint32_t *lwe_in, const uint32_t *ksk, uint32_t *lwe_out) {
    size_t from_n = P_K(p) * P_N(p), to_n = p->lwe_dimension, l = p->ks_levels;
    uint32_t d[32];
    memset(lwe_out, 0, (to_n + 1) * sizeof(uint32_t));
    for (size_t i = 0; i < from_n; i++) {
        orc_decompose(lwe_in[i], p->ks_log_base, p->ks_levels, d);
        for (size_t lev = 0; lev < l; lev++) {
            const uint32_t *row = ksk + (i * l + lev) * (to_n + 1);
            uint32_t a = d[lev];
            for (size_t c = 0; c < to_n + 1; c++) lwe_out[c] += a * row[c]; /* scaled_add :88 */
        }
    }
    for (size_t c = 0; c < to_n + 1; c++) lwe_out[c] = 0u - lwe_out[c];
    lwe_out[to_n] += lwe_in[from_n];
}

/* bootstrapping.rs:67-105 */
int orc_blind_rotate(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk, const uint32_t *tv,
                     uint32_t *acc) {
    size_t N = P_N(p), k = P_K(p), n = p->lwe_dimension, l = p->pbs_levels;
    size_t glwe_sz = (k + 1) * N, ggsw_sz = (k + 1) * l * glwe_sz;
    uint32_t *approx = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint32_t *pt = (uint32_t *)malloc(N * sizeof(uint32_t));
    uint32_t *vx = (uint32_t *)malloc(glwe_sz * sizeof(uint32_t));
    uint32_t *c1 = (uint32_t *)malloc(glwe_sz * sizeof(uint32_t));
    uint32_t *nxt = (uint32_t *)malloc(glwe_sz * sizeof(uint32_t));
    int rc = 0;
    orc_switch_modulus(lwe_in, n + 1, p->log_q, p->glwe_poly_degree + 1, approx); /* :67-71 */
    if (orc_glwe_encode_message(p, tv, N, pt) != 0) { rc = -1; goto done; }       /* :84, assert glwe.rs:144 */
    orc_trivial_encrypt_glwe(p, pt, vx);                                          /* :82 */
    orc_glwe_mul_monomial(p, vx, -(int64_t)approx[n], acc);                       /* :79-86 */
    for (size_t i = 0; i < n; i++) {                                              /* :90-105 */
        orc_glwe_mul_monomial(p, acc, (int64_t)approx[i], c1);
        orc_cmux(p, bsk + i * ggsw_sz, acc, c1, nxt);
        memcpy(acc, nxt, glwe_sz * sizeof(uint32_t));
    }
done:
    free(approx); free(pt); free(vx); free(c1); free(nxt);
    return rc;
}

/* notes/BMMP Bootstrapping.md:13-25 -- blind rotation unrolled by two (the reference holds NO code for it: this is a
 * restatement of the note with the crate's own building blocks; parity against the Rust crate is UNPINNED).
 * Per pair (a, a') = (a~_2i, a~_2i+1) and key triple bk[3i..3i+2]:
 *   bundle = (X^(a+a') - 1) bk[3i] + (X^a - 1) bk[3i+1] + (X^a' - 1) bk[3i+2]   (3 GGSW scalings by a plaintext polynomial,
 *            2 GGSW additions; exact wrapping u32 arithmetic on every GGSW polynomial)
 *   acc    = external_product(bundle, acc) + acc                                (1 GLWE addition)                        */
int orc_blind_rotate_bmmp(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk3, const uint32_t *tv, uint32_t *acc) {
    size_t N = P_N(p), k = P_K(p), n = p->lwe_dimension, l = p->pbs_levels;
    size_t glwe_sz = (k + 1) * N, ggsw_sz = (k + 1) * l * glwe_sz, polys = (k + 1) * l * (k + 1);
    if (n & 1) return -2;
    uint32_t *approx = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint32_t *pt = (uint32_t *)malloc(N * sizeof(uint32_t));
    uint32_t *vx = (uint32_t *)malloc(glwe_sz * sizeof(uint32_t));
    uint32_t *bundle = (uint32_t *)malloc(ggsw_sz * sizeof(uint32_t));
    uint32_t *rot = (uint32_t *)malloc(N * sizeof(uint32_t));
    uint32_t *prod = (uint32_t *)malloc(glwe_sz * sizeof(uint32_t));
    int rc = 0;
    orc_switch_modulus(lwe_in, n + 1, p->log_q, p->glwe_poly_degree + 1, approx);
    if (orc_glwe_encode_message(p, tv, N, pt) != 0) { rc = -1; goto done; }
    orc_trivial_encrypt_glwe(p, pt, vx);
    orc_glwe_mul_monomial(p, vx, -(int64_t)approx[n], acc);
    for (size_t i = 0; i < n / 2; i++) {
        const int64_t e[3] = {(int64_t)approx[2 * i] + (int64_t)approx[2 * i + 1], (int64_t)approx[2 * i], (int64_t)approx[2 * i + 1]};
        memset(bundle, 0, ggsw_sz * sizeof(uint32_t));
        for (int which = 0; which < 3; which++) {
            const uint32_t *g = bsk3 + (3 * i + which) * ggsw_sz;
            for (size_t q = 0; q < polys; q++) {
                orc_poly_mul_monomial(g + q * N, N, e[which], rot);              /* X^e * poly   (utils.rs:183-207) */
                for (size_t j = 0; j < N; j++) bundle[q * N + j] += rot[j] - g[q * N + j];
            }
        }
        orc_external_product(p, bundle, acc, prod);                               /* ggsw.rs:132-161 */
        for (size_t j = 0; j < glwe_sz; j++) acc[j] += prod[j];                   /* glwe.rs:37-41   */
    }
done:
    free(approx); free(pt); free(vx); free(bundle); free(rot); free(prod);
    return rc;
}
int orc_bootstrap_bmmp(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk3, const uint32_t *ksk, const uint32_t *tv,
                       uint32_t *lwe_out) {
    size_t N = P_N(p), k = P_K(p);
    uint32_t *acc = (uint32_t *)malloc((k + 1) * N * sizeof(uint32_t));
    uint32_t *ext = (uint32_t *)malloc((k * N + 1) * sizeof(uint32_t));
    int rc = orc_blind_rotate_bmmp(p, lwe_in, bsk3, tv, acc);
    if (rc == 0) {
        orc_sample_extract(p, acc, 0, ext);
        orc_key_switch_lwe(p, ext, ksk, lwe_out);
    }
    free(acc); free(ext);
    return rc;
}

/* bootstrapping.rs:58-120 */
int orc_bootstrap(const orc_params *p, const uint32_t *lwe_in, const uint32_t *bsk, const uint32_t *ksk,
                  const uint32_t *tv, uint32_t *lwe_out) {
    size_t N = P_N(p), k = P_K(p);
    uint32_t *acc = (uint32_t *)malloc((k + 1) * N * sizeof(uint32_t));
    uint32_t *ext = (uint32_t *)malloc((k * N + 1) * sizeof(uint32_t));
    int rc = orc_blind_rotate(p, lwe_in, bsk, tv, acc);
    if (rc == 0) {
        orc_sample_extract(p, acc, 0, ext);      /* :108 */
        orc_key_switch_lwe(p, ext, ksk, lwe_out); /* :111-117 */
    }
    free(acc); free(ext);
    return rc;
}

typedef struct {
    const orc_params *p; const uint32_t *in, *bsk, *ksk, *tv; uint32_t *out;
    size_t batch; size_t *next; pthread_mutex_t *mu; int rc;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    size_t n1 = j->p->lwe_dimension + 1;
    for (;;) {
        pthread_mutex_lock(j->mu);
        size_t b = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (b >= j->batch) break;
        int r = orc_bootstrap(j->p, j->in + b * n1, j->bsk, j->ksk, j->tv, j->out + b * n1);
        if (r != 0) j->rc = r;
    }
    return NULL;
}

/* one PBS per worker thread; each PBS is the reference's single-threaded code path */
int orc_bootstrap_batch(const orc_params *p, const uint32_t *lwe_in, size_t batch, const uint32_t *bsk,
                        const uint32_t *ksk, const uint32_t *tv, uint32_t *lwe_out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    size_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        batch_job j = {p, lwe_in, bsk, ksk, tv, lwe_out, batch, &next, &mu, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc != 0) rc = jobs[t].rc;
    }
    return rc;
}

/* ------------------------------------------------------------------ test_vector.rs */
/* test_vector.rs:38-67 */
int orc_test_vector_from_lut(const orc_params *p, const uint32_t *lut, size_t lut_len, uint32_t *tv) {
    uint32_t plaintext_modulus = 1u << p->log_p;
    if (lut_len != plaintext_modulus) return -1; /* assert :41 */
    size_t N = P_N(p);
    size_t repetition = N / plaintext_modulus;
    uint32_t *t = (uint32_t *)malloc(N * sizeof(uint32_t));
    size_t c = 0;
    for (size_t v = 0; v < lut_len; v++)
        for (size_t r = 0; r < repetition; r++) t[c++] = lut[v];
    for (size_t i = 0; i < repetition / 2; i++)
        if (t[i] != 0) t[i] = plaintext_modulus - t[i];
    size_t rot = repetition / 2; /* rotate_left :64 */
    for (size_t i = 0; i < c; i++) tv[i] = t[(i + rot) % c];
    for (size_t i = c; i < N; i++) tv[i] = 0;
    free(t);
    return 0;
}

/* test_vector.rs:23-35 */
void orc_test_vector_identity(const orc_params *p, uint32_t *tv) {
    uint32_t pm = 1u << p->log_p;
    uint32_t *lut = (uint32_t *)malloc(pm * sizeof(uint32_t));
    for (uint32_t i = 0; i < pm; i++) lut[i] = i;
    orc_test_vector_from_lut(p, lut, pm, tv);
    free(lut);
}

static uint32_t gate_f(int op, uint32_t l, uint32_t r) {
    switch (op) {
    case 0: return l & r;
    case 1: return l | r;
    default: return l ^ r;
    }
}

/* test_vector.rs:5-20: lut[i] = f((i>>1)&1, i&1) */
int orc_test_vector_boolean(const orc_params *p, int op, uint32_t *tv) {
    if (op < 0 || op > 2) return -1;
    uint32_t pm = 1u << p->log_p;
    uint32_t *lut = (uint32_t *)malloc(pm * sizeof(uint32_t));
    for (uint32_t i = 0; i < pm; i++) lut[i] = gate_f(op, (i >> 1) & 1, i & 1);
    int rc = orc_test_vector_from_lut(p, lut, pm, tv);
    free(lut);
    return rc;
}

/* ------------------------------------------------------------------ lwe.rs */
/* lwe.rs:83-88 */
int orc_lwe_encode(const orc_params *p, uint32_t m, uint32_t *out) {
    if (!(m < (1u << p->log_p))) return -1;
    *out = m << (p->log_q - (p->log_p + p->padding_bits));
    return 0;
}
/* lwe.rs:102-107: bare right shift (floor, no mask) H5 */
uint32_t orc_lwe_decode(const orc_params *p, uint32_t pt) {
    return pt >> (p->log_q - (p->log_p + p->padding_bits));
}
/* lwe.rs:162-173 */
uint32_t orc_lwe_decrypt(const uint32_t *sk, size_t n, const uint32_t *ct) {
    uint32_t a_s = 0;
    for (size_t i = 0; i < n; i++) a_s += sk[i] * ct[i];
    return ct[n] - a_s;
}
/* lwe.rs:9-15 */
void orc_lwe_add(const uint32_t *a, const uint32_t *b, size_t len, uint32_t *out) {
    for (size_t i = 0; i < len; i++) out[i] = a[i] + b[i];
}
/* lwe.rs:17-23 */
void orc_lwe_mul_scalar(const uint32_t *a, uint32_t s, size_t len, uint32_t *out) {
    for (size_t i = 0; i < len; i++) out[i] = a[i] * s;
}

/* ------------------------------------------------------------------ boolean.rs */
/* boolean.rs:9-53: ct_in = 2*ct1 + ct0, bootstrap with gate LUT.  NAND/NOR/XNOR are
 * trivial(1) - {AND,OR,XOR} (a naive f(0,0)!=0 LUT is wrong in the reference). */
int orc_gate(const orc_params *p, int op, const uint32_t *ct0, const uint32_t *ct1, const uint32_t *bsk,
             const uint32_t *ksk, uint32_t *out) {
    if (op < 0 || op > 5) return -1;
    size_t n1 = p->lwe_dimension + 1, N = P_N(p);
    uint32_t *tv = (uint32_t *)malloc(N * sizeof(uint32_t));
    uint32_t *t2 = (uint32_t *)malloc(n1 * sizeof(uint32_t));
    uint32_t *cin = (uint32_t *)malloc(n1 * sizeof(uint32_t));
    int rc = orc_test_vector_boolean(p, op % 3, tv);
    if (rc == 0) {
        orc_lwe_mul_scalar(ct1, 2u, n1, t2); /* boolean.rs:18 */
        orc_lwe_add(t2, ct0, n1, cin);
        rc = orc_bootstrap(p, cin, bsk, ksk, tv, out);
    }
    if (rc == 0 && op >= 3) {
        uint32_t one;
        orc_lwe_encode(p, 1, &one);
        for (size_t j = 0; j < n1; j++) out[j] = 0u - out[j];
        out[n1 - 1] += one;
    }
    free(tv); free(t2); free(cin);
    return rc;
}

/* ------------------------------------------------------------------ seeded RNG + keygen */
typedef struct { uint64_t s[4]; } orc_rng;
static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_init(orc_rng *r, uint64_t seed, uint64_t domain, uint64_t index) {
    uint64_t x = seed ^ (domain * 0x9E3779B97F4A7C15ULL) ^ (index * 0xD1B54A32D192ED03ULL);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_u64(orc_rng *r) { /* xoshiro256** */
    uint64_t *s = r->s;
    uint64_t result = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static uint32_t rng_u32(orc_rng *r) { return (uint32_t)(rng_u64(r) >> 32); }
static double rng_gauss(orc_rng *r, double std_dev) { /* Box-Muller, cosine branch */
    double u1 = (double)((rng_u64(r) >> 11) + 1) * (1.0 / 9007199254740992.0);
    double u2 = (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0);
    return std_dev * sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
/* utils.rs:68-93 sample_binary_array: bits of successive random bytes, LSB first */
static void rng_binary(orc_rng *r, uint32_t *out, size_t len) {
    uint8_t cur = (uint8_t)rng_u32(r);
    int bit = 0;
    for (size_t i = 0; i < len; i++) {
        out[i] = (cur >> bit) & 1;
        if (++bit == 8) { cur = (uint8_t)rng_u32(r); bit = 0; }
    }
}

/* glwe.rs:190-209 encrypt_glwe_zero */
static void encrypt_glwe_zero(const orc_params *p, const uint32_t *glwe_sk, orc_rng *r, uint32_t *ct) {
    size_t N = P_N(p), k = P_K(p);
    for (size_t i = 0; i < k * N; i++) ct[i] = rng_u32(r);          /* :195 sample_uniform_array */
    orc_poly_dot_product(ct, glwe_sk, k, N, N, N, ct + k * N);      /* :197 */
    for (size_t j = 0; j < N; j++)                                   /* :200-203 */
        ct[k * N + j] += orc_f64_to_torus(rng_gauss(r, p->glwe_std_dev));
}

/* lwe.rs:117-136 encrypt_lwe_zero (error sampled first, then the mask) */
static void encrypt_lwe_zero(const uint32_t *sk, size_t n, double std_dev, orc_rng *r, uint32_t *ct) {
    uint32_t error = orc_f64_to_torus(rng_gauss(r, std_dev));
    uint32_t a_s = 0;
    for (size_t i = 0; i < n; i++) { ct[i] = rng_u32(r); a_s += sk[i] * ct[i]; }
    ct[n] = a_s + error;
}

/* bootstrapping.rs:23-56 + ggsw.rs:76-130 + key_switching.rs:20-60 */
void orc_keygen(const orc_params *p, uint64_t seed, uint32_t *lwe_sk, uint32_t *glwe_sk, uint32_t *bsk,
                uint32_t *ksk) {
    size_t N = P_N(p), k = P_K(p), n = p->lwe_dimension, l = p->pbs_levels;
    size_t glwe_sz = (k + 1) * N, ggsw_sz = (k + 1) * l * glwe_sz;
    orc_rng r;
    rng_init(&r, seed, 3, 0);
    rng_binary(&r, lwe_sk, n);          /* lwe.rs:54-58 */
    rng_init(&r, seed, 4, 0);
    rng_binary(&r, glwe_sk, k * N);     /* glwe.rs:177-181 */
    uint32_t log_q_by_log_base = p->log_q / p->pbs_log_base; /* ggsw.rs:90-91 */
    for (size_t i = 0; i < n; i++) {    /* bootstrapping.rs:32-38 */
        rng_init(&r, seed, 1, i);
        uint32_t m = lwe_sk[i];
        uint32_t *g = bsk + i * ggsw_sz;
        for (size_t poly = 0; poly < k + 1; poly++)          /* ggsw.rs:83 */
            for (size_t lev = 0; lev < l; lev++) {           /* ggsw.rs:92 */
                uint32_t *row = g + (poly * l + lev) * glwe_sz;
                encrypt_glwe_zero(p, glwe_sk, &r, row);
                if (m != 0) {                                /* ggsw.rs:96-103 */
                    uint32_t factor = m * (1u << (p->pbs_log_base * (log_q_by_log_base - (lev + 1))));
                    row[poly * N + 0] += factor;
                }
            }
    }
    /* LweSecretKey::from(&glwe_sk) = row-major flatten (lwe.rs:62-73) == glwe_sk itself */
    size_t from_n = k * N, lks = p->ks_levels;
    uint32_t lfull = p->log_q / p->ks_log_base; /* key_switching.rs:39 */
    for (size_t s = 0; s < from_n; s++) {
        rng_init(&r, seed, 2, s);
        for (size_t lev = 0; lev < lks; lev++) {
            uint32_t factor = (1u << (p->ks_log_base * (lfull - (lev + 1)))) * glwe_sk[s]; /* :41-43 */
            uint32_t *row = ksk + (s * lks + lev) * (n + 1);
            encrypt_lwe_zero(lwe_sk, n, p->lwe_std_dev, &r, row);
            row[n] += factor; /* :47-48 */
        }
    }
}

/* lwe.rs:138-160 */
void orc_lwe_encrypt(const orc_params *p, const uint32_t *sk, size_t n, uint32_t plaintext, uint64_t seed,
                     uint64_t index, uint32_t *ct) {
    orc_rng r;
    rng_init(&r, seed, 5, index);
    encrypt_lwe_zero(sk, n, p->lwe_std_dev, &r, ct);
    ct[n] += plaintext;
}
