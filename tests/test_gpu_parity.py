"""GPU parity tests: every entry point of the C-ABI vs the CPU oracle, bit for bit (-m gpu).

Sizes are chosen so the oracle (reference algorithm, O(N^2) products) finishes in seconds: full
polynomial sizes, small LWE dimension n (every CMUX step runs the same code; n only sets the loop
count), plus a few full-n cases.  Full-size batches are covered by size-independent properties
(decrypt-level correctness, batch-invariance, determinism) in test_gpu_fullsize.py.
"""
import numpy as np
import pytest

import tfhe_research_b200 as T
from oracle import orc

pytestmark = pytest.mark.gpu

FIELDS = [f for f, _ in T.TfheParams._fields_]


def oparams(p):
    return orc.params(**{f: getattr(p, f) for f in FIELDS})


class Env:
    """preset may carry the arithmetic path: "P1" (2-prime NTT) or "P1:fft" (exact FP64 FFT, limb-split key)."""

    def __init__(self, preset, n, seed=0xB200):
        preset, _, path = preset.partition(":")
        self.p = T.TfheParams.preset(preset, lwe_dimension=n)
        self.o = oparams(self.p)
        self.lwe_sk, self.glwe_sk, self.bsk, self.ksk = T.bootstrapping_key_gen(self.p, seed)
        self.ctx = T.Context(self.p, 0, path={"fft": T.PATH_FFT, "ntt": T.PATH_NTT}.get(path))   # None: the library's default
        if path == "fft":
            self.ctx.set_fft_check(True)   # also record the distance-to-integer of every rounded value
        self.bk = self.ctx.upload_key(self.bsk, self.ksk)

    def enc(self, m, idx, seed=1):
        return T.encrypt_lwe_plaintext(self.p, self.lwe_sk, T.encode_message(self.p, m), seed, idx)

    def dec(self, ct):
        return T.decode_rounded(self.p, T.decrypt_lwe(self.lwe_sk, ct))


_envs = {}


def env(preset, n):
    key = (preset, n)
    if key not in _envs:
        _envs[key] = Env(preset, n)
    return _envs[key]


CASES = [("P0:ntt", 4), ("P1:ntt", 3), ("P2:ntt", 2), ("P1:fft", 3), ("P1:fft", 21), ("P0:fft", 4), ("P0:fft", 7), ("P2:fft", 2), ("P2:fft", 5)]


def r32(rng, *shape):
    return rng.integers(0, 1 << 32, shape, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("preset,n", CASES)
def test_switch_modulus_decompose_monomial(preset, n):
    e = env(preset, n)
    rng = np.random.default_rng(1)
    L = orc.lib()
    import ctypes as C
    edge = np.array([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x0000F800, 0x0FF80000, 0xF8F8F8F8, 0x7FFFFF80, 0x80, 0xABCDEF12,
                     0x001FFFFF, 0x00200000, 0xFFDFFFFF, 0xFFE00000], dtype=np.uint32)
    v = np.concatenate([edge, r32(rng, 5000)])
    exp = orc.z(len(v))
    L.orc_switch_modulus(v, len(v), 32, e.p.glwe_poly_degree + 1, exp)
    assert np.array_equal(e.ctx.switch_modulus(v), exp)
    for which, (lb, lv) in enumerate(((e.p.pbs_log_base, e.p.pbs_levels), (e.p.ks_log_base, e.p.ks_levels))):
        got = e.ctx.decompose(v, which)
        d = orc.z(lv)
        for i in range(0, len(v), 7):
            L.orc_decompose(int(v[i]), lb, lv, d)
            assert np.array_equal(got[i], d), hex(int(v[i]))
    N = e.p.N
    glwe = r32(rng, 6, e.p.k + 1, N)
    idx = np.array([0, 1, N, 2 * N - 1, -1, -(N + 3)], dtype=np.int64)
    got = e.ctx.glwe_mul_monomial(glwe, idx)
    for b in range(6):
        o = orc.z((e.p.k + 1) * N)
        L.orc_glwe_mul_monomial(C.byref(e.o), glwe[b].reshape(-1), int(idx[b]), o)
        assert np.array_equal(got[b].reshape(-1), o)


@pytest.mark.parametrize("preset,n", CASES)
def test_external_product_and_cmux(preset, n):
    import ctypes as C
    e = env(preset, n)
    rng = np.random.default_rng(2)
    L = orc.lib()
    B, gsz, gg = 5, e.p.glwe_words, e.p.ggsw_words
    glwe0, glwe1 = r32(rng, B, e.p.k + 1, e.p.N), r32(rng, B, e.p.k + 1, e.p.N)
    glwe0[0, 0, :8] = [0xFFFFFFFF, 0x7FFFFF80, 0x0000F800, 0xF8F8F8F8, 0, 0x80000000, 0x0FF80000, 0x00FFFFFF]
    gi = np.array([i % n for i in range(B)], dtype=np.uint32)
    got = e.ctx.external_product(e.bk, gi, glwe0)
    for b in range(B):
        o = orc.z(gsz)
        L.orc_external_product(C.byref(e.o), e.bsk[int(gi[b]) * gg:(int(gi[b]) + 1) * gg], glwe0[b].reshape(-1), o)
        assert np.array_equal(got[b].reshape(-1), o), b
    got = e.ctx.cmux(e.bk, gi, glwe0, glwe1)
    for b in range(B):
        o = orc.z(gsz)
        c1 = glwe1[b].copy().reshape(-1)
        L.orc_cmux(C.byref(e.o), e.bsk[int(gi[b]) * gg:(int(gi[b]) + 1) * gg], glwe0[b].reshape(-1), c1, o)
        assert np.array_equal(got[b].reshape(-1), o), b


@pytest.mark.parametrize("preset,n", CASES)
def test_blind_rotate_extract_keyswitch_bootstrap(preset, n):
    import ctypes as C
    e = env(preset, n)
    L = orc.lib()
    rng = np.random.default_rng(3)
    pm = 1 << e.p.log_p
    B = 9
    cts = np.stack([e.enc(i % pm, i) for i in range(B)])
    cts[B - 1] = r32(rng, n + 1)                    # arbitrary (non-ciphertext) words are legal inputs too
    cts[B - 2, :n] = 0                               # every a~_i == 0: all CMUX steps skipped
    tvs = np.stack([T.construct_identity_test_vector(e.p), rng.integers(0, pm, e.p.N).astype(np.uint32)])
    idx = np.array([b % 2 for b in range(B)], dtype=np.uint32)
    acc = e.ctx.blind_rotate(e.bk, cts, tvs, idx)
    ext = e.ctx.sample_extract(acc)
    ks = e.ctx.key_switch(e.bk, ext)
    out = e.ctx.bootstrap(e.bk, cts, tvs, idx)
    for b in range(B):
        exp_acc = orc.blind_rotate(e.o, cts[b], e.bsk, tvs[idx[b]])
        assert np.array_equal(acc[b], exp_acc), b
        exp_ext = orc.z(e.p.k * e.p.N + 1)
        L.orc_sample_extract(C.byref(e.o), exp_acc.reshape(-1), 0, exp_ext)
        assert np.array_equal(ext[b], exp_ext), b
        exp_ks = orc.z(n + 1)
        L.orc_key_switch_lwe(C.byref(e.o), exp_ext, e.ksk, exp_ks)
        assert np.array_equal(ks[b], exp_ks), b
        assert np.array_equal(out[b], orc.bootstrap(e.o, cts[b], e.bsk, e.ksk, tvs[idx[b]])), b
    for b in range(0, B - 2, 2):                     # identity LUT ciphertexts decrypt to their message
        assert e.dec(out[b]) == b % pm
    for nb in (1, 3, 8):                             # small batches take the split-row (atomic) key-switch path
        assert np.array_equal(e.ctx.key_switch(e.bk, ext[:nb]), ks[:nb]), nb
        assert np.array_equal(e.ctx.bootstrap(e.bk, cts[:nb], tvs, idx[:nb]), out[:nb]), nb


@pytest.mark.parametrize("preset,n", [("P1:fft", 21), ("P0:fft", 7), ("P2:fft", 5), ("P1:fft", 630)])
def test_key_switch_tensor_core_path_vs_imad_vs_oracle(preset, n):
    """key_switching.rs:63-103 on arbitrary input words: the integer-tensor-core products (byte planes; KS_MMA = mma.sync,
    KS_TCGEN05 = tcgen05.mma with the accumulator in tensor memory) and the 32-bit multiply-add product (KS_IMAD) give the
    oracle's bits, for ragged batch sizes and extreme words."""
    import ctypes as C
    e = env(preset, n)
    L = orc.lib()
    rng = np.random.default_rng(17)
    kN = e.p.k * e.p.N
    for B in (9, 130, 301):
        lwe = r32(rng, B, kN + 1)
        lwe[0, :] = 0xFFFFFFFF                        # rounding wraps to 0 (decomposer.rs:39)
        lwe[1, :] = 0x7FFFFFFF
        lwe[2, :kN:2] = 0x80000000
        lwe[3, :] = 0xF8F8F8F8                        # windows of B-1 with carries: digit +B
        e.ctx.set_ks_path(T.KS_MMA)
        got_mma = e.ctx.key_switch(e.bk, lwe)
        e.ctx.set_ks_path(T.KS_IMAD)
        got_imad = e.ctx.key_switch(e.bk, lwe)
        e.ctx.set_ks_path(T.KS_TCGEN05)               # tcgen05.mma kind::i8 + TMEM + TMA (falls back to MMA where it does not apply)
        got_tc5 = e.ctx.key_switch(e.bk, lwe)
        e.ctx.set_ks_path(T.KS_MMA)
        assert np.array_equal(got_mma, got_imad), B
        assert np.array_equal(got_tc5, got_imad), B
        for b in (0, 1, 2, 3, B - 1):
            exp = orc.z(n + 1)
            L.orc_key_switch_lwe(C.byref(e.o), lwe[b], e.ksk, exp)
            assert np.array_equal(got_mma[b], exp), (B, b)


@pytest.mark.parametrize("preset,n", [("P0:ntt", 4), ("P1:ntt", 3), ("P1:fft", 3), ("P0:fft", 4)])
def test_gates_boolean_rs(preset, n):
    e = env(preset, n)
    fs = [lambda a, b: a & b, lambda a, b: a | b, lambda a, b: a ^ b,
          lambda a, b: 1 - (a & b), lambda a, b: 1 - (a | b), lambda a, b: 1 - (a ^ b)]
    ct0 = np.stack([e.enc(i & 1, 100 + i) for i in range(4)])
    ct1 = np.stack([e.enc((i >> 1) & 1, 200 + i) for i in range(4)])
    lin = e.ctx.gate_linear(ct0, ct1)
    assert np.array_equal(lin, (ct1 * np.uint32(2) + ct0).astype(np.uint32))
    for op, f in enumerate(fs):
        out = e.ctx.gate(e.bk, op, ct0, ct1)
        for i in range(4):
            assert np.array_equal(out[i], orc.gate(e.o, op, ct0[i], ct1[i], e.bsk, e.ksk)), (op, i)
            assert e.dec(out[i]) == f((i >> 1) & 1, i & 1), (op, i)
    ops = np.array([0, 3, 2, 5], dtype=np.uint8)     # mixed gates in one batch
    out = e.ctx.gate(e.bk, ops, ct0, ct1)
    for i in range(4):
        assert np.array_equal(out[i], orc.gate(e.o, int(ops[i]), ct0[i], ct1[i], e.bsk, e.ksk))


def test_encode_assert_and_errors():
    e = env("P0:ntt", 4)
    bad = np.full(e.p.N, 4, dtype=np.uint32)        # >= 2^log_p: assert! glwe.rs:144
    with pytest.raises(T.TfheError) as ei:
        e.ctx.bootstrap(e.bk, np.zeros((1, 5), dtype=np.uint32), bad)
    assert ei.value.code == T.TFHE_E_ASSERT
    with pytest.raises(T.TfheError) as ei:
        e.ctx.external_product(e.bk, np.array([99], dtype=np.uint32), np.zeros((1, 3, 512), dtype=np.uint32))
    assert ei.value.code == T.TFHE_E_PARAM
    # empty batch is a no-op
    out = e.ctx.bootstrap(e.bk, np.zeros((0, 5), dtype=np.uint32), T.construct_identity_test_vector(e.p))
    assert out.shape == (0, 5)


def test_device_pointers_torch():
    import torch
    e = env("P0", 4)
    cts = np.stack([e.enc(i % 4, 300 + i) for i in range(8)])
    tv = T.construct_identity_test_vector(e.p)
    ref = e.ctx.bootstrap(e.bk, cts, tv)
    d_in = torch.from_numpy(cts.view(np.int32)).cuda()
    d_tv = torch.from_numpy(tv.view(np.int32)).cuda()
    d_out = e.ctx.bootstrap(e.bk, d_in, d_tv)
    assert d_out.is_cuda
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), ref)


def test_reference_defaults_full_n_one_gate():
    """BASELINE config #1: one bootstrapped NAND with the reference's non-test defaults (n = 722)."""
    e = env("P0", 722)
    ct0, ct1 = e.enc(1, 1)[None], e.enc(1, 2)[None]
    out = e.ctx.gate(e.bk, T.NAND, ct0, ct1)
    assert np.array_equal(out[0], orc.gate(e.o, 3, ct0[0], ct1[0], e.bsk, e.ksk))
    assert e.dec(out[0]) == 0


@pytest.mark.parametrize("preset,n", CASES)
def test_negacyclic_mul_utils_rs_155(preset, n):
    """poly_mul (Toeplitz product, utils.rs:113-160) of signed digit-range polynomials with arbitrary u32 polynomials."""
    e = env(preset, n)
    L = orc.lib()
    N = e.p.N
    rng = np.random.default_rng(9)
    B = 6
    a = rng.integers(-1024, 1025, (B, N)).astype(np.int32)
    a[0, :4] = [1024, -1024, 0, 1]
    a[1] = 0
    a[1, 1] = 1                                     # times X: a pure negacyclic shift
    g = r32(rng, B, N)
    g[2] = 0xFFFFFFFF
    got = e.ctx.negacyclic_mul(a, g)
    for b in range(B):
        exp = orc.z(N)
        L.orc_poly_mul(a[b].astype(np.uint32), g[b], N, exp)
        assert np.array_equal(got[b], exp), b
    assert np.array_equal(got[1], np.concatenate([(-g[1, -1:].astype(np.int64) & 0xFFFFFFFF).astype(np.uint32), g[1, :-1]]))
    # arbitrary u32 x u32, as the reference's poly_mul takes them (extreme words included)
    big = r32(rng, 4, N)
    big[0, :3] = [0xFFFFFFFF, 0x80000000, 0x7FFFFFFF]
    big[1] = 0xFFFFFFFF
    got = e.ctx.negacyclic_mul(big.view(np.int32), g[:4])
    for b in range(4):
        exp = orc.z(N)
        L.orc_poly_mul(big[b], g[b], N, exp)
        assert np.array_equal(got[b], exp), b


@pytest.mark.parametrize("preset,n", [("P0:ntt", 4), ("P1:ntt", 3), ("P1:fft", 3), ("P0:fft", 4)])
def test_ks_first_ordering_notes_tfhe_md_365(preset, n):
    """key switch FIRST, then blind rotation + sample extraction (input/output under the kN-dim key)."""
    import ctypes as C
    e = env(preset, n)
    L = orc.lib()
    p = e.p
    kN = p.k * p.N
    big_sk = T.lwe_secret_key_from_glwe(e.glwe_sk)                 # lwe.rs:62-73
    pm = 1 << p.log_p
    B = 6
    cts = np.stack([T.encrypt_lwe_plaintext(p, big_sk, T.encode_message(p, i % pm), 5, i) for i in range(B)])
    tv = T.construct_identity_test_vector(p)
    out = e.ctx.bootstrap_ks_first(e.bk, cts, tv)
    assert out.shape == (B, kN + 1)
    for b in range(B):
        ks = orc.z(n + 1)
        L.orc_key_switch_lwe(C.byref(e.o), cts[b], e.ksk, ks)
        acc = orc.blind_rotate(e.o, ks, e.bsk, tv)
        ext = orc.z(kN + 1)
        L.orc_sample_extract(C.byref(e.o), acc.reshape(-1), 0, ext)
        assert np.array_equal(out[b], ext), b
        assert T.decode_rounded(p, T.decrypt_lwe(big_sk, out[b])) == b % pm


@pytest.mark.parametrize("preset,n", [("P2", 2)])
def test_k_input_gates_notes_boolean_gates_md(preset, n):
    """3- and 4-input gates on a 4-bit message space; ct_in = sum 2^i c_i (Horner steps of boolean.rs:18)."""
    e = env(preset, n)
    p = e.p
    pm = 1 << p.log_p
    delta = 1 << (p.log_q - p.log_p - p.padding_bits)
    for k, tt in ((3, 0b11101000), (3, 0b00000001), (4, 0x6996), (4, 0x8001)):   # majority3, NOR3, parity4, (all-equal)4
        B = 1 << k
        cts = [np.stack([e.enc((j >> i) & 1, 300 + 16 * i + j) for j in range(B)]) for i in range(k)]
        out = e.ctx.gate_k(e.bk, tt, cts)
        neg = tt & 1
        lut = np.zeros(pm, dtype=np.uint32)
        for j in range(1 << k):
            lut[j] = ((tt >> j) & 1) ^ neg
        tv = T.construct_test_from_lut(p, lut)
        for j in range(B):
            lin = np.zeros(n + 1, dtype=np.uint32)
            for i in range(k):
                lin = (lin + (cts[i][j].astype(np.uint64) << np.uint64(i)).astype(np.uint32)).astype(np.uint32)
            exp = orc.bootstrap(e.o, lin, e.bsk, e.ksk, tv)
            if neg:
                exp = (0 - exp.astype(np.int64)).astype(np.uint32)
                exp[n] = np.uint32((int(exp[n]) + delta) & 0xFFFFFFFF)
            assert np.array_equal(out[j], exp), (k, hex(tt), j)
            assert e.dec(out[j]) == (tt >> j) & 1, (k, hex(tt), j)
    with pytest.raises(T.TfheError):
        e.ctx.gate_k(e.bk, 0b10, [cts[0]] * 5)      # k > log_p


def test_bmmp_variant_bit_exact_vs_oracle():
    """SURVEY 8(f) N1 / BASELINE config #5: unrolled-by-two blind rotation with key triples, through the ordinary entry
    points, bit for bit against oracle/tfhe_oracle.c orc_blind_rotate_bmmp (the reference has prose only: parity with
    the Rust crate is unpinned; this pins the kernel against the note's restatement)."""
    n = 6
    p = T.TfheParams.preset("P1", lwe_dimension=n)
    o = oparams(p)
    lwe_sk, glwe_sk, bsk3, ksk = T.bootstrapping_key_gen_bmmp(p, 0xB200)
    ctx = T.Context(p, 0, path=T.PATH_FFT)
    ctx.set_fft_check(True)
    bk = ctx.upload_key_bmmp(bsk3, ksk)
    rng = np.random.default_rng(11)
    pm = 1 << p.log_p
    B = 7                                            # 3 + 3 + 1: a partially filled CTA too
    cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(B)])
    cts[5] = r32(rng, n + 1)
    cts[6, :n] = 0                                   # all pairs (0, 0): every step skipped
    cts[4, 0] = 0                                    # a = 0, a' != 0 in the first pair
    tvs = np.stack([T.construct_identity_test_vector(p), rng.integers(0, pm, p.N).astype(np.uint32)])
    idx = np.array([b % 2 for b in range(B)], dtype=np.uint32)
    acc = ctx.blind_rotate(bk, cts, tvs, idx)
    out = ctx.bootstrap(bk, cts, tvs, idx)
    for b in range(B):
        assert np.array_equal(acc[b], orc.blind_rotate_bmmp(o, cts[b], bsk3, tvs[idx[b]])), b
        assert np.array_equal(out[b], orc.bootstrap_bmmp(o, cts[b], bsk3, ksk, tvs[idx[b]])), b
    for b in (0, 2):
        assert T.decode_rounded(p, T.decrypt_lwe(lwe_sk, out[b])) == b % pm
    assert 0.0 < ctx.fft_rounding_margin() < 2.0 ** -6
    with pytest.raises(T.TfheError):                 # a BMMP key holds products of key bits: no single external products
        ctx.external_product(bk, np.array([0], dtype=np.uint32), np.zeros((1, 2, 1024), dtype=np.uint32))
    with pytest.raises(T.TfheError):                 # odd n
        T.bootstrapping_key_gen_bmmp(T.TfheParams.preset("P1", lwe_dimension=5), 1)


def test_fft_path_selection_and_limits():
    """The FFT path is the default where instantiated; beyond its shared-memory bound on n the NTT path serves."""
    p = T.TfheParams.preset("P1", lwe_dimension=8)
    ctx = T.Context(p, 0)
    assert ctx.pbs_path == T.PATH_FFT
    assert ctx.fft_rounding_margin() == 0.0              # nothing recorded unless checking is switched on
    ctx.close()
    big = T.TfheParams.preset("P1", lwe_dimension=1200)  # mask of 3 resident ciphertexts no longer fits
    ctx = T.Context(big, 0)
    assert ctx.pbs_path == T.PATH_NTT
    with pytest.raises(T.TfheError) as ei:
        ctx.set_pbs_path(T.PATH_FFT)
    assert ei.value.code == T.TFHE_E_PARAM
    ctx.close()
    p2 = T.TfheParams.preset("P2", lwe_dimension=2)      # N = 2048: FFT path with half-row key slots, no BMMP variant
    ctx = T.Context(p2, 0)
    assert ctx.pbs_path == T.PATH_FFT
    with pytest.raises(T.TfheError):
        ctx.upload_key_bmmp(np.zeros(3 * p2.ggsw_words, dtype=np.uint32), np.zeros(p2.ksk_words, dtype=np.uint32))
    ctx.close()
