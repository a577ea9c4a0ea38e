"""Pins the CPU oracle (oracle/tfhe_oracle.c) -- no GPU.

(1) the reference's own deterministic tests, (2) the known-answer vectors of SURVEY.md 9-C,
(3) C restatement vs independent numpy restatement on seeded random inputs, (4) the reference's
functional tests (decrypt-level) with a rounding decoder in the harness.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import np_oracle as npo
from oracle import orc

L = orc.lib()
M32 = 0xFFFFFFFF


def s32(x):
    return [(int(v) ^ 0x80000000) - 0x80000000 for v in x]


def dec(v, lb, lv):
    out = orc.z(lv)
    L.orc_decompose(v, lb, lv, out)
    return s32(out)


# ---- (1) reference deterministic tests ---------------------------------------------------------
def test_poly_mul_works_utils_rs_265():
    v0 = np.array([12, 4, 123, 43, 3, 2, 3], dtype=np.uint32)
    v1 = np.array([12, 232, 5, 3, 2, 4, 2], dtype=np.uint32)
    for faithful in (0, 1):
        L.orc_set_faithful_toeplitz(faithful)
        a, b = orc.z(7), orc.z(7)
        L.orc_poly_mul(v0, v1, 7, a)
        L.orc_school_book_negacyclic_mul(v0, v1, 7, b)
        assert a.tolist() == b.tolist() == [4294966139, 2387, 2353, 29088, 10647, 1354, 930]
    L.orc_set_faithful_toeplitz(0)
    assert npo.poly_mul(v0, v1).tolist() == a.tolist()
    assert npo.school_book_negacyclic_mul(v0, v1).tolist() == a.tolist()


def test_decomposition_decomposer_rs_103():
    # reference: i in 0..1e8 with (4, 7); here a 2e5 prefix + 2e5 seeded samples over the full range
    rng = np.random.default_rng(7)
    vals = list(range(200_000)) + rng.integers(0, 1 << 32, 200_000, dtype=np.uint64).tolist()
    out = orc.z(7)
    for lb, lv in ((4, 7), (4, 6), (4, 5), (8, 3), (2, 8), (8, 4), (4, 8)):
        for v in vals[:: 1 if (lb, lv) == (4, 7) else 50]:
            L.orc_decompose(v, lb, lv, out[:lv])
            assert L.orc_recompose(out[:lv], lb, lv) == L.orc_round_value(v, lb, lv)


# ---- (2) SURVEY 9-C KATs -----------------------------------------------------------------------
def test_kat_monomial():
    p = np.array([1, 2, 3, 4], dtype=np.uint32)
    exp = {1: [M32 - 3, 1, 2, 3], 5: [4, M32, M32 - 1, M32 - 2], -1: [2, 3, 4, M32], -5: [M32 - 1, M32 - 2, M32 - 3, 1]}
    for idx, e in exp.items():
        out = orc.z(4)
        L.orc_poly_mul_monomial(p, 4, idx, out)
        assert out.tolist() == e
        assert npo.poly_mul_monomial(p, idx).tolist() == e


def test_kat_decompose():
    assert dec(0x12345678, 4, 6) == [1, 2, 3, 4, 5, 6]
    assert L.orc_round_value(0x12345678, 4, 6) == 0x12345600
    assert dec(0x0000F800, 4, 6) == [0, 0, 0, 0, 16, -8]
    assert dec(0x0FF80000, 4, 6) == [1, -1, 16, -8, 0, 0]
    assert dec(0xF8F8F8F8, 4, 6) == [16, -8, 16, -8, 16, -7]
    assert L.orc_round_value(0xF8F8F8F8, 4, 6) == 0xF8F8F900
    assert dec(0xFFFFFFFF, 4, 6) == [0] * 6
    assert L.orc_round_value(0xFFFFFFFF, 4, 6) == 0
    assert dec(0x7FFFFF80, 4, 6) == [-8, 0, 0, 0, 0, 0]
    assert dec(0x00000080, 4, 6) == [0, 0, 0, 0, 0, 1]
    assert dec(0xABCDEF12, 4, 5) == [-5, -4, -3, -2, -1]
    assert L.orc_round_value(0xABCDEF12, 4, 5) == 0xABCDF000


def test_kat_switch_modulus():
    v = np.array([0x001FFFFF, 0x00200000, 0x005FFFFF, 0x00600000, 0x80000000, 0xFFDFFFFF, 0xFFE00000, 0xFFFFFFFF],
                 dtype=np.uint32)
    out = orc.z(8)
    L.orc_switch_modulus(v, 8, 32, 10, out)
    assert out.tolist() == [0, 1, 1, 2, 512, 1023, 0, 0]
    assert npo.switch_modulus(v, 32, 10).tolist() == out.tolist()


def rle(a):
    out = []
    for v in a:
        if out and out[-1][0] == v:
            out[-1][1] += 1
        else:
            out.append([int(v), 1])
    return [tuple(x) for x in out]


def test_kat_test_vectors():
    p = orc.params()
    assert rle(orc.test_vector_identity(p)) == [(0, 64), (1, 128), (2, 128), (3, 128), (0, 64)]
    assert rle(orc.test_vector_boolean(p, 0)) == [(0, 320), (1, 128), (0, 64)]
    q = npo.Params()
    assert npo.test_vector_identity(q).tolist() == orc.test_vector_identity(p).tolist()
    for op, f in ((0, lambda a, b: a & b), (1, lambda a, b: a | b), (2, lambda a, b: a ^ b)):
        assert npo.test_vector_boolean(q, f).tolist() == orc.test_vector_boolean(p, op).tolist()
    # H6: a naive NAND LUT would be [1*320, 0*128, 3*64]
    assert rle(orc.test_vector_from_lut(p, [1, 1, 1, 0])) == [(1, 320), (0, 128), (3, 64)]
    with pytest.raises(AssertionError):
        orc.test_vector_from_lut(p, [0, 1, 2])


def test_kat_f64_to_torus_saturates_h4():
    assert L.orc_f64_to_torus(-1e-6) == 0
    assert L.orc_f64_to_torus(0.25) == 1 << 30
    assert L.orc_f64_to_torus(1.25) == 1 << 30
    assert L.orc_f64_to_torus(0.75) == 0  # 0.75 - round(0.75) = -0.25 -> clamps to 0


def test_defaults_lib_rs_76():
    p, t = orc.params(), orc.params(True)
    assert (p.k, p.N, p.n, p.log_p, p.padding_bits, p.log_q) == (2, 512, 722, 2, 1, 32)
    assert (p.pbs_log_base, p.pbs_levels, p.ks_log_base, p.ks_levels) == (4, 6, 4, 5)
    assert t.n == 4 and t.N == 512


# ---- (3) C restatement vs numpy restatement ----------------------------------------------------
SMALL = [
    dict(glwe_dimension=2, glwe_poly_degree=5, lwe_dimension=3, pbs_log_base=4, pbs_levels=6, ks_log_base=4, ks_levels=5, log_p=2),
    dict(glwe_dimension=1, glwe_poly_degree=6, lwe_dimension=2, pbs_log_base=8, pbs_levels=3, ks_log_base=2, ks_levels=8, log_p=2),
    dict(glwe_dimension=1, glwe_poly_degree=5, lwe_dimension=2, pbs_log_base=8, pbs_levels=4, ks_log_base=4, ks_levels=8, log_p=3),
]


@pytest.mark.parametrize("cfg", SMALL)
def test_c_vs_numpy_random(cfg):
    p = orc.params(**cfg)
    q = npo.Params(**cfg)
    rng = np.random.default_rng(11)
    N, k, l = p.N, p.k, p.pbs_levels
    r32 = lambda *s: rng.integers(0, 1 << 32, s, dtype=np.uint64).astype(np.uint32)
    a, b = r32(N), r32(N)
    o = orc.z(N)
    L.orc_poly_mul(a, b, N, o)
    assert o.tolist() == npo.poly_mul(a, b).tolist()
    L.orc_set_faithful_toeplitz(1)
    o2 = orc.z(N)
    L.orc_poly_mul(a, b, N, o2)
    L.orc_set_faithful_toeplitz(0)
    assert o.tolist() == o2.tolist()
    t = orc.z(N * N)
    L.orc_teoplitz(a, N, t)
    assert t.reshape(N, N).tolist() == npo.teoplitz(a).tolist()
    for idx in (0, 1, N - 1, N, N + 3, 2 * N - 1, -1, -N, -(2 * N - 1), 5 * N + 2):
        L.orc_poly_mul_monomial(a, N, idx, o)
        assert o.tolist() == npo.poly_mul_monomial(a, idx).tolist()
    glwe, glwe1, ggsw = r32(k + 1, N), r32(k + 1, N), r32((k + 1) * l, k + 1, N)
    d = orc.z((k + 1) * l * N)
    L.orc_decompose_glwe_ciphertext(C.byref(p), glwe.reshape(-1), d)
    assert d.reshape(-1, N).tolist() == npo.decompose_glwe_ciphertext(q, glwe).tolist()
    e = orc.z((k + 1) * N)
    L.orc_external_product(C.byref(p), ggsw.reshape(-1), glwe.reshape(-1), e)
    assert e.reshape(k + 1, N).tolist() == npo.external_product(q, ggsw, glwe).tolist()
    c1 = glwe1.copy().reshape(-1)
    L.orc_cmux(C.byref(p), ggsw.reshape(-1), glwe.reshape(-1), c1, e)
    assert e.reshape(k + 1, N).tolist() == npo.cmux(q, ggsw, glwe, glwe1).tolist()
    assert c1.tolist() == (glwe1 - glwe).reshape(-1).tolist()  # H8: ct1 clobbered with the difference
    se = orc.z(k * N + 1)
    for si in (0, 1, N - 1):
        L.orc_sample_extract(C.byref(p), glwe.reshape(-1), si, se)
        assert se.tolist() == npo.sample_extract(q, glwe, si).tolist()
    ksk = r32(k * N * p.ks_levels, p.n + 1)
    big = r32(k * N + 1)
    ko = orc.z(p.n + 1)
    L.orc_key_switch_lwe(C.byref(p), big, ksk.reshape(-1), ko)
    assert ko.tolist() == npo.key_switch_lwe(q, big, ksk).tolist()
    # full bootstrap on random (meaningless) keys: still a deterministic function of its inputs
    bsk = r32(p.n, (k + 1) * l, k + 1, N)
    lwe = r32(p.n + 1)
    tv = rng.integers(0, 1 << p.log_p, N).astype(np.uint32)
    assert orc.bootstrap(p, lwe, bsk.reshape(-1), ksk.reshape(-1), tv).tolist() == npo.bootstrap(q, lwe, bsk, ksk, tv).tolist()
    assert orc.blind_rotate(p, lwe, bsk.reshape(-1), tv).tolist() == npo.blind_rotate(q, lwe, bsk, tv).tolist()


def test_encode_assert_h10():
    p = orc.params(True)
    tv = np.full(p.N, 4, dtype=np.uint32)
    s = p.sizes()
    with pytest.raises(AssertionError):
        orc.bootstrap(p, orc.z(p.n + 1), orc.z(s["bsk"]), orc.z(s["ksk"]), tv)


# ---- (4) the reference's functional tests, decrypt-level ---------------------------------------
@pytest.fixture(scope="module")
def keys_test_cfg():
    p = orc.params(True)  # cfg(test): n = 4
    return p, orc.keygen(p, 0xB200)


def test_encrypt_and_decrypt_lwe_lwe_rs_183(keys_test_cfg):
    p, (lwe_sk, *_r) = keys_test_cfg
    for m in range(4):
        ct = orc.lwe_encrypt(p, lwe_sk, m, 1, m)
        assert orc.lwe_decrypt_round(p, lwe_sk, ct) == m


def test_key_switching_works_key_switching_rs_118(keys_test_cfg):
    p, (lwe_sk, glwe_sk, bsk, ksk) = keys_test_cfg
    for m in range(4):
        big = orc.lwe_encrypt(p, glwe_sk, m, 2, m)  # under the flattened GLWE key (lwe.rs:62-73)
        out = orc.z(p.n + 1)
        L.orc_key_switch_lwe(C.byref(p), big, ksk, out)
        assert orc.lwe_decrypt_round(p, lwe_sk, out) == m


def test_bootstrapping_works_bootstrapping_rs_194(keys_test_cfg):
    p, (lwe_sk, glwe_sk, bsk, ksk) = keys_test_cfg
    tv = orc.test_vector_identity(p)
    for m in range(4):
        ct = orc.lwe_encrypt(p, lwe_sk, m, 3, m)
        out = orc.bootstrap(p, ct, bsk, ksk, tv)
        assert orc.lwe_decrypt_round(p, lwe_sk, out) == m


def test_boolean_gates_work_boolean_rs_67(keys_test_cfg):
    p, (lwe_sk, glwe_sk, bsk, ksk) = keys_test_cfg
    fs = [lambda a, b: a & b, lambda a, b: a | b, lambda a, b: a ^ b,
          lambda a, b: 1 - (a & b), lambda a, b: 1 - (a | b), lambda a, b: 1 - (a ^ b)]
    for i in range(4):
        lhs, rhs = (i >> 1) & 1, i & 1
        ct1 = orc.lwe_encrypt(p, lwe_sk, lhs, 4, 2 * i)
        ct0 = orc.lwe_encrypt(p, lwe_sk, rhs, 4, 2 * i + 1)
        for op, f in enumerate(fs):
            out = orc.gate(p, op, ct0, ct1, bsk, ksk)
            assert orc.lwe_decrypt_round(p, lwe_sk, out) == f(lhs, rhs), (op, lhs, rhs)


def test_keygen_layout_ggsw_rs_83(keys_test_cfg):
    # row r = poly*l + level of GGSW_i decrypts (GLWE) to s_i * 2^(4*(8-level-1)) at coeff 0 of poly `poly`
    p, (lwe_sk, glwe_sk, bsk, ksk) = keys_test_cfg
    N, k, l = p.N, p.k, p.pbs_levels
    g = bsk.reshape(p.n, (k + 1) * l, k + 1, N)
    sk = glwe_sk.reshape(k, N)
    for i in range(p.n):
        for row in (0, l - 1, l, (k + 1) * l - 1):
            poly, lev = divmod(row, l)
            body = g[i, row, k].copy()
            o = orc.z(N)
            for r in range(k):
                L.orc_poly_mul(np.ascontiguousarray(g[i, row, r]), np.ascontiguousarray(sk[r]), N, o)
                body = body - o
            factor = int(lwe_sk[i]) << (4 * (8 - lev - 1))
            # phase = e - factor*s_poly (mask polys) or e + factor (body poly); only check the body rows exactly
            if poly == k:
                err = (int(body[0]) - factor) & M32
                assert err < (1 << 12) or err > M32 - (1 << 12)


def test_bmmp_unrolled_blind_rotation_decrypts():
    """SURVEY 8(f) N1 (notes/BMMP Bootstrapping.md; no reference code, parity UNPINNED): the oracle's restatement of the
    unrolled-by-two blind rotation, on key triples from the host keygen, bootstraps every message correctly."""
    import tfhe_research_b200 as T
    for preset, n in (("P0", 4), ("P1", 6)):
        p = T.TfheParams.preset(preset, lwe_dimension=n)
        o = orc.params(**{f: getattr(p, f) for f, _ in T.TfheParams._fields_})
        lwe_sk, glwe_sk, bsk3, ksk = T.bootstrapping_key_gen_bmmp(p, 0xB200)
        lwe_sk2, glwe_sk2, bsk, ksk2 = T.bootstrapping_key_gen(p, 0xB200)
        assert np.array_equal(lwe_sk, lwe_sk2) and np.array_equal(glwe_sk, glwe_sk2) and np.array_equal(ksk, ksk2)
        assert bsk3.size == 3 * (n // 2) * p.ggsw_words
        tv = T.construct_identity_test_vector(p)
        for m in range(1 << p.log_p):
            ct = T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, m), 9, m)
            out = orc.bootstrap_bmmp(o, ct, bsk3, ksk, tv)
            assert T.decode_rounded(p, T.decrypt_lwe(lwe_sk, out)) == m, (preset, m)
            # same ciphertext through the standard chain decrypts to the same message (different key material, different bits)
            assert T.decode_rounded(p, T.decrypt_lwe(lwe_sk, orc.bootstrap(o, ct, bsk, ksk, tv))) == m


def test_key_switch_byte_plane_identity():
    """The tensor-core key switch (kernels.cuh K4-MMA) rests on: sum_r d_r * w_r = sum_pl 2^(8 pl) * (sum_r d_r * byte_pl(w_r))
    mod 2^32, with every inner sum an exact 32-bit integer.  Checked here in plain integers against the oracle's
    key_switching.rs:63-103 restatement, digits from the oracle's decomposer (incl. the +B digit), extreme key words."""
    import ctypes as C
    o = orc.params(True)                       # k=2, N=512, n=4, ks (4, 5)
    L = orc.lib()
    rng = np.random.default_rng(23)
    kN, n, lev, logb = 2 * 512, 4, 5, 4
    ksk = rng.integers(0, 1 << 32, (kN * lev, n + 1), dtype=np.uint64).astype(np.uint32)
    ksk[:7] = 0xFFFFFFFF
    ksk[7:9] = 0x80000000
    lwe = rng.integers(0, 1 << 32, kN + 1, dtype=np.uint64).astype(np.uint32)
    lwe[:16] = 0xF8F8F8F8                      # windows of B-1 plus carry: digit +B
    exp = orc.z(n + 1)
    L.orc_key_switch_lwe(C.byref(o), lwe, ksk, exp)
    digits = np.zeros(kN * lev, dtype=np.int64)                                                   # signed, MSB first
    d = orc.z(lev)
    for i in range(kN):
        L.orc_decompose(int(lwe[i]), logb, lev, d)
        digits[i * lev:(i + 1) * lev] = d.view(np.int32)
    assert digits.max() == 16 and digits.min() >= -8
    planes = [((ksk.astype(np.int64) >> (8 * pl)) & 255) for pl in range(4)]
    inner = [digits @ pb for pb in planes]                                                        # exact integers
    assert max(int(np.abs(x).max()) for x in inner) < 2 ** 31                                     # fit the s32 accumulators
    total = sum(int(1 << (8 * pl)) * inner[pl] for pl in range(4))
    got = (-total) % (1 << 32)
    got[n] = (got[n] + int(lwe[kN])) % (1 << 32)
    assert np.array_equal(got.astype(np.uint32), exp)
