"""numpy restatement of Janmajayamall/tfhe-research's hot path -- TEST INFRASTRUCTURE ONLY.

Second, independent restatement of the reference (the first is oracle/tfhe_oracle.c); the two are
diffed against each other in tests/test_oracle.py to harden a parity claim that cannot be pinned
against the Rust binary (no cargo in this image, no golden vectors in the reference).
Written from the Rust sources, not from the C file.  Citations are into /root/reference/src/.
All arithmetic is numpy uint32 (wrapping).  Only small sizes: pure-numpy O(N^2) products.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

U32 = np.uint32


@dataclass
class Params:  # lib.rs:23-34
    glwe_dimension: int = 2
    glwe_poly_degree: int = 9  # log2 N
    lwe_dimension: int = 722
    padding_bits: int = 1
    log_p: int = 2
    log_q: int = 32
    ks_log_base: int = 4
    ks_levels: int = 5
    pbs_log_base: int = 4
    pbs_levels: int = 6
    lwe_std_dev: float = 0.000013071021089943935
    glwe_std_dev: float = 0.00000004990272175010415

    @property
    def N(self) -> int:
        return 1 << self.glwe_poly_degree

    @property
    def k(self) -> int:
        return self.glwe_dimension


def switch_modulus(values, log_from, log_to):
    """utils.rs:13-33: round(v / 2^(log_from-log_to)) mod 2^log_to, half-up."""
    v = np.asarray(values, dtype=np.uint64)
    d = np.uint64(1 << (log_from - log_to))
    rational = v // d
    fractional = v % d
    res = rational + (fractional + (d >> np.uint64(1))) // d
    return (res % np.uint64(1 << log_to)).astype(U32)


def teoplitz(p):
    """utils.rs:113-153: matrix[i][c] = p[i-c] for c<=i, -p[n+i-c] for c>i."""
    p = np.asarray(p, dtype=U32)
    n = p.shape[0]
    i = np.arange(n)[:, None]
    c = np.arange(n)[None, :]
    idx = (i - c) % n
    m = p[idx]
    neg = (c > i)
    return np.where(neg, (-m.astype(np.int64)).astype(U32), m).astype(U32)


def poly_mul(p0, p1):
    """utils.rs:155-160 (ndarray dot on u32 wraps in release builds)."""
    t = teoplitz(p0).astype(np.uint64)
    return ((t @ np.asarray(p1, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)).astype(U32)


def poly_dot_product(a, b):
    """utils.rs:163-173."""
    res = poly_mul(a[0], b[0])
    for r0, r1 in zip(a[1:], b[1:]):
        res = res + poly_mul(r0, r1)
    return res.astype(U32)


def school_book_negacyclic_mul(p0, p1):
    """utils.rs:221-236."""
    n = len(p0)
    res = np.zeros(n, dtype=U32)
    a = [int(x) for x in p0]
    b = [int(x) for x in p1]
    for i in range(n):
        s = 0
        for j in range(i + 1):
            s += a[j] * b[i - j]
        for j in range(i + 1, n):
            s -= a[j] * b[n - (j - i)]
        res[i] = s & 0xFFFFFFFF
    return res


def poly_mul_monomial(p0, monomial_index):
    """utils.rs:183-207."""
    p0 = np.asarray(p0, dtype=U32)
    n = p0.shape[0]
    idx = (monomial_index % (1 << 64)) % (2 * n)  # `as usize % (2 * n)`
    flip_sign, degree = divmod(idx, n)
    out = p0.copy()
    if flip_sign:
        out = (-out.astype(np.int64)).astype(U32)
    out = np.roll(out, degree)  # rotate_right
    out[:degree] = (-out[:degree].astype(np.int64)).astype(U32)
    return out


def round_value(value, log_base, levels):
    """decomposer.rs:27-40."""
    value &= 0xFFFFFFFF
    ignored_bits = 32 - log_base * levels
    if ignored_bits == 0:
        return value
    ignored_value = value & ((1 << ignored_bits) - 1)
    ignored_msb = ignored_value >> (ignored_bits - 1)
    return (((value >> ignored_bits) + ignored_msb) << ignored_bits) & 0xFFFFFFFF


def decompose(value, log_base, levels):
    """decomposer.rs:42-80 (pure python ints, one value)."""
    value = round_value(value, log_base, levels)
    base_mask = (1 << log_base) - 1
    half = 1 << (log_base - 1)
    carry = 0
    dec = []
    for l in range(32 // log_base):
        res = ((value >> (log_base * l)) & base_mask) + carry
        carry_mask = res & half
        res = (res - (carry_mask << 1)) & 0xFFFFFFFF
        carry = carry_mask >> (log_base - 1)
        dec.append(res)
    dec.reverse()
    return dec[:levels]


def recompose(legs, log_base, levels):
    """decomposer.rs:83-95."""
    value = 0
    for index, leg in enumerate(legs):
        value = (value + (leg << (log_base * (levels - 1 - index)))) & 0xFFFFFFFF
    return (value << (32 - log_base * levels)) & 0xFFFFFFFF


def decompose_glwe_ciphertext(p: Params, ct):
    """glwe.rs:69-108 -> [(k+1)*l, N], row = poly*l + level."""
    l = p.pbs_levels
    out = np.zeros(((p.k + 1) * l, p.N), dtype=U32)
    for r in range(p.k + 1):
        for j in range(p.N):
            out[r * l:(r + 1) * l, j] = decompose(int(ct[r, j]), p.pbs_log_base, l)
    return out


def glwe_mul_monomial(ct, index):
    """glwe.rs:20-34."""
    return np.stack([poly_mul_monomial(row, index) for row in ct])


def external_product(p: Params, ggsw, glwe):
    """ggsw.rs:132-161; ggsw is [(k+1)l, k+1, N]."""
    dec = decompose_glwe_ciphertext(p, glwe)
    return np.stack([poly_dot_product(dec, ggsw[:, col, :]) for col in range(p.k + 1)]).astype(U32)


def cmux(p: Params, ggsw, ct0, ct1):
    """ggsw.rs:164-178."""
    diff = (ct1 - ct0).astype(U32)
    return (external_product(p, ggsw, diff) + ct0).astype(U32)


def sample_extract(p: Params, glwe, sample_index=0):
    """bootstrapping.rs:122-156."""
    out = []
    for poly in glwe[:-1]:
        out.extend(poly[sample_index::-1].tolist())
        out.extend(((-poly[:sample_index:-1].astype(np.int64)) & 0xFFFFFFFF).tolist())
    out.append(int(glwe[p.k, sample_index]))
    return np.array(out, dtype=U32)


def key_switch_lwe(p: Params, lwe, ksk):
    """key_switching.rs:63-103; ksk is [kN*l_ks, n+1]."""
    from_n = p.k * p.N
    digits = []
    for a in lwe[:from_n]:
        digits.extend(decompose(int(a), p.ks_log_base, p.ks_levels))
    d = np.array(digits, dtype=np.uint64)
    s = (d[:, None] * ksk.astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    total = (s.sum(axis=0) & np.uint64(0xFFFFFFFF)).astype(U32)
    total = (-total.astype(np.int64)).astype(U32)
    total[p.lwe_dimension] = (int(total[p.lwe_dimension]) + int(lwe[from_n])) & 0xFFFFFFFF
    return total


def encode_glwe(p: Params, msg):
    """glwe.rs:141-151."""
    msg = np.asarray(msg, dtype=U32)
    assert (msg < (1 << p.log_p)).all()
    return (msg << U32(p.log_q - (p.log_p + p.padding_bits))).astype(U32)


def blind_rotate(p: Params, lwe, bsk, tv):
    """bootstrapping.rs:67-105; bsk is [n, (k+1)l, k+1, N]."""
    approx = switch_modulus(lwe, p.log_q, p.glwe_poly_degree + 1)
    vx = np.zeros((p.k + 1, p.N), dtype=U32)
    vx[p.k] = encode_glwe(p, tv)
    acc = glwe_mul_monomial(vx, -int(approx[p.lwe_dimension]))
    for i in range(p.lwe_dimension):
        c1 = glwe_mul_monomial(acc, int(approx[i]))
        acc = cmux(p, bsk[i], acc, c1)
    return acc


def bootstrap(p: Params, lwe, bsk, ksk, tv):
    """bootstrapping.rs:58-120."""
    acc = blind_rotate(p, lwe, bsk, tv)
    return key_switch_lwe(p, sample_extract(p, acc, 0), ksk)


def test_vector_from_lut(p: Params, lut):
    """test_vector.rs:38-67."""
    pm = 1 << p.log_p
    assert len(lut) == pm
    rep = p.N // pm
    tv = []
    for v in lut:
        tv.extend([int(v)] * rep)
    for i in range(rep // 2):
        if tv[i] != 0:
            tv[i] = pm - tv[i]
    tv = tv[rep // 2:] + tv[:rep // 2]
    return np.array(tv, dtype=U32)


def test_vector_boolean(p: Params, f):
    """test_vector.rs:5-20."""
    return test_vector_from_lut(p, [f((i >> 1) & 1, i & 1) for i in range(1 << p.log_p)])


def test_vector_identity(p: Params):
    """test_vector.rs:23-35."""
    return test_vector_from_lut(p, list(range(1 << p.log_p)))


# keep pytest from collecting the reference-named helpers above
test_vector_from_lut.__test__ = False
test_vector_boolean.__test__ = False
test_vector_identity.__test__ = False
