"""GPU tests at BASELINE.json's full sizes (-m gpu): size-independent properties + a few oracle spot checks.

Full-n oracle runs cost seconds each (reference algorithm), so they are limited to a handful of
ciphertexts; everything else is checked through properties the domain offers: decrypt-level
correctness of every sampled output, batch invariance (a ciphertext's result does not depend on its
position or on the batch size), determinism, trivial-ciphertext closed forms, and plain evaluation of
the same boolean circuit.
"""
import numpy as np
import pytest
import torch

import tfhe_research_b200 as T
from oracle import orc
from tfhe_research_b200 import circuit

pytestmark = pytest.mark.gpu
FIELDS = [f for f, _ in T.TfheParams._fields_]


class Env:
    def __init__(self, preset):
        preset, _, path = preset.partition(":")   # "P1" = 2-prime NTT path, "P1:fft" = exact FP64-FFT path
        self.p = T.TfheParams.preset(preset)
        self.o = orc.params(**{f: getattr(self.p, f) for f in FIELDS})
        self.lwe_sk, self.glwe_sk, self.bsk, self.ksk = T.bootstrapping_key_gen(self.p, 0xB200)
        self.ctx = T.Context(self.p, 0, path={"fft": T.PATH_FFT, "ntt": T.PATH_NTT}.get(path))   # None: the library's default
        if path == "fft":
            self.ctx.set_fft_check(True)   # also record the distance-to-integer of every rounded value
        self.bk = self.ctx.upload_key(self.bsk, self.ksk)

    def enc(self, m, idx):
        return T.encrypt_lwe_plaintext(self.p, self.lwe_sk, T.encode_message(self.p, m), 1, idx)

    def dec(self, ct):
        return T.decode_rounded(self.p, T.decrypt_lwe(self.lwe_sk, ct))


_E = {}


def env(preset):
    if preset not in _E:
        _E[preset] = Env(preset)
    return _E[preset]


def make_batch(e, B, n_unique=128):
    pm = 1 << e.p.log_p
    uniq = np.stack([e.enc(i % pm, i) for i in range(n_unique)])
    return np.tile(uniq, ((B + n_unique - 1) // n_unique, 1))[:B].copy(), n_unique


@pytest.mark.parametrize("which", ["P1:ntt", "P1:fft"])
def test_p1_batch_4096_identity(which):
    """BASELINE config #2: batch of 4096 PBS, N=1024, n=630, identity test vector."""
    e = env(which)
    pm = 1 << e.p.log_p
    cts, nu = make_batch(e, 4096)
    tv = T.construct_identity_test_vector(e.p)
    out = e.ctx.bootstrap(e.bk, cts, tv)
    for i in range(0, 4096, 37):
        assert e.dec(out[i]) == (i % nu) % pm
    # batch invariance + determinism: same ciphertext, different position / batch size -> identical bits
    assert np.array_equal(out[:nu], out[nu:2 * nu])
    small = e.ctx.bootstrap(e.bk, cts[5:8], tv)
    assert np.array_equal(small, out[5:8])
    # 445 ciphertexts spread over two waves of CTAs holding 1 or 2 each (4096: 2 or 3 each): same bits per ciphertext
    assert np.array_equal(e.ctx.bootstrap(e.bk, cts[:445], tv), out[:445])
    # oracle spot check at full n (reference algorithm, ~1 s each)
    for i in (0, 77):
        assert np.array_equal(out[i], orc.bootstrap(e.o, cts[i], e.bsk, e.ksk, tv))
    if which.endswith(":fft"):
        # a-posteriori exactness certificate: no value was further than 2^-6 from an integer before rounding
        # (a-priori bound 2^-9); and the two arithmetic paths agree bit for bit on the whole batch
        m = e.ctx.fft_rounding_margin()
        assert 0.0 < m < 2.0 ** -6, m
        assert np.array_equal(out, env("P1:ntt").ctx.bootstrap(env("P1:ntt").bk, cts, tv))
        e.ctx.set_fft_check(False)          # the production kernel (no margin recording) gives the same bits
        assert np.array_equal(out, e.ctx.bootstrap(e.bk, cts, tv))
        e.ctx.set_fft_check(True)


@pytest.mark.parametrize("which", ["P1:ntt", "P1:fft"])
def test_p1_trivial_ciphertexts_closed_form(which):
    """a = 0 => every a~_i = 0 => all CMUX steps are skipped and acc = X^{-b~} * v(X) exactly."""
    e = env(which)
    p = e.p
    N, n = p.N, p.n
    tv = T.construct_identity_test_vector(p)
    rng = np.random.default_rng(5)
    cts = np.zeros((64, n + 1), dtype=np.uint32)
    cts[:, n] = rng.integers(0, 1 << 32, 64, dtype=np.uint64).astype(np.uint32)
    acc = e.ctx.blind_rotate(e.bk, cts, tv)
    enc = (tv.astype(np.uint64) << np.uint64(p.log_q - p.log_p - p.padding_bits)).astype(np.uint32)
    for b in range(64):
        bt = ((int(cts[b, n]) + (1 << (31 - p.glwe_poly_degree - 1))) >> (31 - p.glwe_poly_degree)) % (2 * N)
        j = np.arange(N)
        src = (j + bt) % (2 * N)
        exp = enc[src % N].astype(np.int64)
        exp = np.where(src >= N, -exp, exp) & 0xFFFFFFFF
        assert np.array_equal(acc[b, p.k], exp.astype(np.uint32))
        assert not acc[b, :p.k].any()


def test_p1_skipped_steps_in_shared_ctas():
    """Steps with a~_i = 0 are skipped per ciphertext while the other ciphertexts of the same CTA (which share the key
    stream, and on the FFT path relay key rows to each other through TMEM) run them: zero some mask words of some
    ciphertexts of a batch that fills every CTA with 2-3 ciphertexts; both arithmetic paths must agree bit for bit."""
    e, en = env("P1:fft"), env("P1:ntt")
    n = e.p.n
    tv = T.construct_identity_test_vector(e.p)
    cts, nu = make_batch(e, 400)
    rng = np.random.default_rng(11)
    for b in range(0, 400, 3):                      # every third ciphertext: ~20 % of its steps skipped
        cts[b, :n][rng.random(n) < 0.2] = 0
    cts[7, :n] = 0                                  # one trivial ciphertext: all steps skipped
    cts[8, :n // 2] = 0
    acc_f = e.ctx.blind_rotate(e.bk, cts, tv)
    acc_n = en.ctx.blind_rotate(en.bk, cts, tv)
    assert np.array_equal(acc_f, acc_n)
    assert np.array_equal(acc_f[:3], e.ctx.blind_rotate(e.bk, cts[:3], tv))   # batch invariance: 1 ciphertext per CTA
    assert np.array_equal(acc_f[7], e.ctx.blind_rotate(e.bk, cts[7:8], tv)[0])


def test_p2_programmable_lut_batch_16k():
    """BASELINE config #3: programmable LUT bootstrap, N=2048, 4-bit messages, batch 16k."""
    e = env("P2")
    pm = 1 << e.p.log_p
    rng = np.random.default_rng(2)
    lut = rng.integers(0, pm, pm).astype(np.uint32)
    lut[0] = 0  # the reference's LUT construction is only sound for f(0) = 0
    tvs = np.stack([T.construct_test_from_lut(e.p, lut), T.construct_identity_test_vector(e.p)])
    B = 16384
    cts, nu = make_batch(e, B)
    idx = (np.arange(B) // nu % 2).astype(np.uint32)
    out = e.ctx.bootstrap(e.bk, cts, tvs, idx)
    for i in range(0, B, 151):
        m = (i % nu) % pm
        assert e.dec(out[i]) == (int(lut[m]) if idx[i] == 0 else m), i
    assert np.array_equal(out[3], orc.bootstrap(e.o, cts[3], e.bsk, e.ksk, tvs[0]))  # one full-n oracle run


def test_p0_layered_circuit_and_gate_batch():
    """BASELINE config #4 (single GPU slice): mixed NAND/AND/XOR gates, depth-1 and layered."""
    e = env("P0")
    rng = np.random.default_rng(3)
    n_in = 64
    bits = rng.integers(0, 2, n_in)
    wires = np.stack([e.enc(int(b), 1000 + i) for i, b in enumerate(bits)])
    levels = circuit.random_layered_circuit(n_in, [256, 256, 128], seed=3)
    d_w = torch.from_numpy(wires.view(np.int32)).cuda()

    def gate_fn(ops, ct0, ct1):
        return e.ctx.gate(e.bk, np.ascontiguousarray(ops), ct0, ct1)

    res = circuit.evaluate_encrypted(levels, d_w, gate_fn).cpu().numpy().view(np.uint32)
    exp = circuit.evaluate_plain(levels, bits)
    assert [e.dec(r) for r in res] == exp.tolist()
    # depth-1: 4096 independent gates over the three opcodes; one oracle check per opcode
    B = 4096
    ops = rng.choice(np.array([T.NAND, T.AND, T.XOR], dtype=np.uint8), B)
    i0, i1 = rng.integers(0, n_in, B), rng.integers(0, n_in, B)
    out = e.ctx.gate(e.bk, ops, wires[i0], wires[i1])
    f = {T.NAND: lambda l, r: 1 - (l & r), T.AND: lambda l, r: l & r, T.XOR: lambda l, r: l ^ r}
    for b in range(0, B, 29):
        assert e.dec(out[b]) == f[int(ops[b])](int(bits[i1[b]]), int(bits[i0[b]])), b
    for op in (T.NAND, T.AND, T.XOR):
        b = int(np.argmax(ops == op))
        assert np.array_equal(out[b], orc.gate(e.o, int(op), wires[i0[b]], wires[i1[b]], e.bsk, e.ksk))


def test_p1_bmmp_batch_full_n():
    """BASELINE config #5 (single-GPU slice): BMMP-style unrolled bootstrapping with key switching at n = 630."""
    p = T.TfheParams.preset("P1")
    lwe_sk, glwe_sk, bsk3, ksk = T.bootstrapping_key_gen_bmmp(p, 0xB200)
    ctx = T.Context(p, 0, path=T.PATH_FFT)
    ctx.set_fft_check(True)
    bk = ctx.upload_key_bmmp(bsk3, ksk)
    pm = 1 << p.log_p
    nu, B = 64, 1024
    uniq = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(nu)])
    cts = np.tile(uniq, (B // nu, 1))
    tv = T.construct_identity_test_vector(p)
    out = ctx.bootstrap(bk, cts, tv)
    for i in range(0, B, 17):
        assert T.decode_rounded(p, T.decrypt_lwe(lwe_sk, out[i])) == (i % nu) % pm, i
    assert np.array_equal(out[:nu], out[nu:2 * nu])                      # batch invariance
    assert 0.0 < ctx.fft_rounding_margin() < 2.0 ** -6
    o = orc.params(**{f: getattr(p, f) for f in FIELDS})
    assert np.array_equal(out[3], orc.bootstrap_bmmp(o, cts[3], bsk3, ksk, tv))   # one full-n oracle run
