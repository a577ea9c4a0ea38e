"""ctypes loader for the C oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


class OrcParams(C.Structure):
    _fields_ = [
        ("glwe_dimension", C.c_uint32),
        ("glwe_poly_degree", C.c_uint32),
        ("lwe_dimension", C.c_uint32),
        ("padding_bits", C.c_uint32),
        ("log_p", C.c_uint32),
        ("log_q", C.c_uint32),
        ("ks_log_base", C.c_uint32),
        ("ks_levels", C.c_uint32),
        ("pbs_log_base", C.c_uint32),
        ("pbs_levels", C.c_uint32),
        ("lwe_std_dev", C.c_double),
        ("glwe_std_dev", C.c_double),
    ]

    @property
    def N(self):
        return 1 << self.glwe_poly_degree

    @property
    def k(self):
        return self.glwe_dimension

    @property
    def n(self):
        return self.lwe_dimension

    def sizes(self):
        N, k, n, l, lks = self.N, self.k, self.n, self.pbs_levels, self.ks_levels
        return dict(glwe=(k + 1) * N, ggsw=(k + 1) * l * (k + 1) * N, bsk=n * (k + 1) * l * (k + 1) * N,
                    ksk=k * N * lks * (n + 1), lwe=n + 1, lwe_big=k * N + 1)


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "tfhe_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return path


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    PP = C.POINTER(OrcParams)
    sigs = {
        "orc_params_default": (None, [C.c_int, PP]),
        "orc_set_faithful_toeplitz": (None, [C.c_int]),
        "orc_integer_division": (C.c_uint32, [C.c_uint32, C.c_uint32]),
        "orc_switch_modulus": (None, [u32p, C.c_size_t, C.c_uint32, C.c_uint32, u32p]),
        "orc_f64_to_torus": (C.c_uint32, [C.c_double]),
        "orc_teoplitz": (None, [u32p, C.c_size_t, u32p]),
        "orc_poly_mul": (None, [u32p, u32p, C.c_size_t, u32p]),
        "orc_poly_mul_monomial": (None, [u32p, C.c_size_t, C.c_int64, u32p]),
        "orc_school_book_negacyclic_mul": (None, [u32p, u32p, C.c_size_t, u32p]),
        "orc_round_value": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_decompose": (None, [C.c_uint32, C.c_uint32, C.c_uint32, u32p]),
        "orc_recompose": (C.c_uint32, [u32p, C.c_uint32, C.c_uint32]),
        "orc_glwe_mul_monomial": (None, [PP, u32p, C.c_int64, u32p]),
        "orc_decompose_glwe_ciphertext": (None, [PP, u32p, u32p]),
        "orc_glwe_encode_message": (C.c_int, [PP, u32p, C.c_size_t, u32p]),
        "orc_trivial_encrypt_glwe": (None, [PP, u32p, u32p]),
        "orc_external_product": (None, [PP, u32p, u32p, u32p]),
        "orc_cmux": (None, [PP, u32p, u32p, u32p, u32p]),
        "orc_sample_extract": (None, [PP, u32p, C.c_size_t, u32p]),
        "orc_key_switch_lwe": (None, [PP, u32p, u32p, u32p]),
        "orc_blind_rotate": (C.c_int, [PP, u32p, u32p, u32p, u32p]),
        "orc_bootstrap": (C.c_int, [PP, u32p, u32p, u32p, u32p, u32p]),
        "orc_blind_rotate_bmmp": (C.c_int, [PP, u32p, u32p, u32p, u32p]),
        "orc_bootstrap_bmmp": (C.c_int, [PP, u32p, u32p, u32p, u32p, u32p]),
        "orc_bootstrap_batch": (C.c_int, [PP, u32p, C.c_size_t, u32p, u32p, u32p, u32p, C.c_int]),
        "orc_test_vector_from_lut": (C.c_int, [PP, u32p, C.c_size_t, u32p]),
        "orc_test_vector_identity": (None, [PP, u32p]),
        "orc_test_vector_boolean": (C.c_int, [PP, C.c_int, u32p]),
        "orc_lwe_encode": (C.c_int, [PP, C.c_uint32, C.POINTER(C.c_uint32)]),
        "orc_lwe_decode": (C.c_uint32, [PP, C.c_uint32]),
        "orc_lwe_decrypt": (C.c_uint32, [u32p, C.c_size_t, u32p]),
        "orc_lwe_add": (None, [u32p, u32p, C.c_size_t, u32p]),
        "orc_lwe_mul_scalar": (None, [u32p, C.c_uint32, C.c_size_t, u32p]),
        "orc_gate": (C.c_int, [PP, C.c_int, u32p, u32p, u32p, u32p, u32p]),
        "orc_keygen": (None, [PP, C.c_uint64, u32p, u32p, u32p, u32p]),
        "orc_lwe_encrypt": (None, [PP, u32p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint64, u32p]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def params(test_cfg: bool = False, **over) -> OrcParams:
    p = OrcParams()
    lib().orc_params_default(1 if test_cfg else 0, C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def z(*shape):
    return np.zeros(shape, dtype=np.uint32)


# ---- thin numpy-friendly wrappers ------------------------------------------------------------
def keygen(p: OrcParams, seed: int):
    s = p.sizes()
    lwe_sk, glwe_sk, bsk, ksk = z(p.n), z(p.k * p.N), z(s["bsk"]), z(s["ksk"])
    lib().orc_keygen(C.byref(p), seed, lwe_sk, glwe_sk, bsk, ksk)
    return lwe_sk, glwe_sk, bsk, ksk


def lwe_encrypt(p: OrcParams, sk, message: int, seed: int, index: int):
    pt = C.c_uint32()
    assert lib().orc_lwe_encode(C.byref(p), message, C.byref(pt)) == 0
    ct = z(len(sk) + 1)
    lib().orc_lwe_encrypt(C.byref(p), np.ascontiguousarray(sk), len(sk), pt.value, seed, index, ct)
    return ct


def lwe_decrypt_round(p: OrcParams, sk, ct) -> int:
    """decrypt + ROUNDING decode (harness decoder; the reference's own decode is a floor, H5)."""
    pt = lib().orc_lwe_decrypt(np.ascontiguousarray(sk), len(sk), np.ascontiguousarray(ct))
    shift = p.log_q - (p.log_p + p.padding_bits)
    return ((pt + (1 << (shift - 1))) >> shift) & ((1 << (p.log_p + p.padding_bits)) - 1)


def bootstrap(p: OrcParams, lwe_in, bsk, ksk, tv):
    out = z(p.n + 1)
    rc = lib().orc_bootstrap(C.byref(p), np.ascontiguousarray(lwe_in), bsk, ksk, np.ascontiguousarray(tv), out)
    if rc != 0:
        raise AssertionError("reference assert!: test vector entry >= 2^log_p (glwe.rs:144)")
    return out


def blind_rotate(p: OrcParams, lwe_in, bsk, tv):
    out = z((p.k + 1) * p.N)
    rc = lib().orc_blind_rotate(C.byref(p), np.ascontiguousarray(lwe_in), bsk, np.ascontiguousarray(tv), out)
    if rc != 0:
        raise AssertionError("reference assert!: test vector entry >= 2^log_p (glwe.rs:144)")
    return out.reshape(p.k + 1, p.N)


def blind_rotate_bmmp(p: OrcParams, lwe_in, bsk3, tv):
    """notes/BMMP Bootstrapping.md:13-25 (unrolled by two; no reference code: parity unpinned)."""
    out = z(p.k + 1, p.N)
    rc = lib().orc_blind_rotate_bmmp(C.byref(p), np.ascontiguousarray(lwe_in), bsk3, np.ascontiguousarray(tv), out.reshape(-1))
    assert rc == 0, rc
    return out


def bootstrap_bmmp(p: OrcParams, lwe_in, bsk3, ksk, tv):
    out = z(p.n + 1)
    rc = lib().orc_bootstrap_bmmp(C.byref(p), np.ascontiguousarray(lwe_in), bsk3, ksk, np.ascontiguousarray(tv), out)
    assert rc == 0, rc
    return out


def bootstrap_batch(p: OrcParams, lwe_in, bsk, ksk, tv, nthreads: int):
    lwe_in = np.ascontiguousarray(lwe_in, dtype=np.uint32)
    B = lwe_in.shape[0]
    out = z(B, p.n + 1)
    rc = lib().orc_bootstrap_batch(C.byref(p), lwe_in.reshape(-1), B, bsk, ksk, np.ascontiguousarray(tv),
                                   out.reshape(-1), nthreads)
    assert rc == 0
    return out


def gate(p: OrcParams, op: int, ct0, ct1, bsk, ksk):
    out = z(p.n + 1)
    assert lib().orc_gate(C.byref(p), op, np.ascontiguousarray(ct0), np.ascontiguousarray(ct1), bsk, ksk, out) == 0
    return out


def test_vector_identity(p: OrcParams):
    tv = z(p.N)
    lib().orc_test_vector_identity(C.byref(p), tv)
    return tv


def test_vector_boolean(p: OrcParams, op: int):
    tv = z(p.N)
    assert lib().orc_test_vector_boolean(C.byref(p), op, tv) == 0
    return tv


def test_vector_from_lut(p: OrcParams, lut):
    tv = z(p.N)
    lut = np.ascontiguousarray(lut, dtype=np.uint32)
    rc = lib().orc_test_vector_from_lut(C.byref(p), lut, len(lut), tv)
    if rc != 0:
        raise AssertionError("reference assert!: lut.len() == 2^log_p (test_vector.rs:41)")
    return tv


for _f in (test_vector_identity, test_vector_boolean, test_vector_from_lut):
    _f.__test__ = False
