"""Host-side (C++) half of the C-ABI library vs the oracle -- no GPU, no device calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tfhe_research_b200 as T
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void \*|void|uint64_t|size_t|const char \*|tfhe_ctx \*)\s*(tfhe_[a-z0-9_]+)\(", hdr, re.M))
    assert declared == set(T.EXPORTS), declared ^ set(T.EXPORTS)
    L = C.CDLL(T.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def same_params(a, b):
    return all(getattr(a, f) == getattr(b, f) for f, _ in T.TfheParams._fields_)


def test_params_default_and_presets():
    assert same_params(T.TfheParams.default(), orc.params())
    assert same_params(T.TfheParams.default(True), orc.params(True))
    assert same_params(T.TfheParams.preset("P0"), orc.params())
    p1, p2 = T.TfheParams.preset("P1"), T.TfheParams.preset("P2")
    assert (p1.k, p1.N, p1.n, p1.pbs_log_base, p1.pbs_levels, p1.ks_log_base, p1.ks_levels) == (1, 1024, 630, 8, 3, 2, 8)
    assert (p2.k, p2.N, p2.n, p2.log_p) == (1, 2048, 742, 4)
    for p in (T.TfheParams.default(), T.TfheParams.default(True), p1, p2):
        p.validate()
    with pytest.raises(T.TfheError):
        T.TfheParams.preset("nope")


@pytest.mark.parametrize("over", [dict(log_q=64), dict(pbs_log_base=5), dict(pbs_levels=9), dict(ks_log_base=3),
                                  dict(log_p=10), dict(glwe_dimension=3), dict(pbs_log_base=0)])
def test_params_validate_rejects(over):
    p = T.TfheParams.preset("P0", **over)
    with pytest.raises(T.TfheError) as e:
        p.validate()
    assert e.value.code == T.TFHE_E_PARAM


@pytest.mark.parametrize("preset", ["P0", "P1", "P2"])
def test_test_vectors_match_oracle(preset):
    p = T.TfheParams.preset(preset)
    o = orc.params(**{f: getattr(p, f) for f, _ in T.TfheParams._fields_})
    assert np.array_equal(T.construct_identity_test_vector(p), orc.test_vector_identity(o))
    for g in (T.AND, T.OR, T.XOR):
        assert np.array_equal(T.construct_test_vector_boolean(p, g), orc.test_vector_boolean(o, g))
    rng = np.random.default_rng(5)
    lut = rng.integers(0, 1 << p.log_p, 1 << p.log_p).astype(np.uint32)
    assert np.array_equal(T.construct_test_from_lut(p, lut), orc.test_vector_from_lut(o, lut))
    with pytest.raises(T.TfheError) as e:
        T.construct_test_from_lut(p, lut[:-1])
    assert e.value.code == T.TFHE_E_ASSERT  # assert! test_vector.rs:41


def test_encode_decode_lwe_rs_83_107():
    p = T.TfheParams.default()
    assert T.encode_message(p, 3) == 3 << 29
    with pytest.raises(T.TfheError) as e:
        T.encode_message(p, 4)
    assert e.value.code == T.TFHE_E_ASSERT
    assert T.decode(p, (3 << 29) - 1) == 2          # floor, H5
    assert T.decode(p, 0xFFFFFFFF) == 7             # no mask, H5
    assert T.decode_rounded(p, (3 << 29) - 1) == 3


SMALL = [dict(), dict(glwe_dimension=1, glwe_poly_degree=10, lwe_dimension=5, pbs_log_base=8, pbs_levels=3, ks_log_base=2, ks_levels=8)]


@pytest.mark.parametrize("over", SMALL)
def test_keygen_and_encrypt_bit_identical_to_oracle(over):
    p = T.TfheParams.default(True)
    for k_, v in over.items():
        setattr(p, k_, v)
    o = orc.params(**{f: getattr(p, f) for f, _ in T.TfheParams._fields_})
    mine = T.bootstrapping_key_gen(p, 0x1234)
    ref = orc.keygen(o, 0x1234)
    for a, b, name in zip(mine, ref, ("lwe_sk", "glwe_sk", "bsk", "ksk")):
        assert np.array_equal(a, b), name
    for i in range(8):
        pt = T.encode_message(p, i % 4)
        assert np.array_equal(T.encrypt_lwe_plaintext(p, mine[0], pt, 9, i), orc.lwe_encrypt(o, ref[0], i % 4, 9, i))
        assert T.decode_rounded(p, T.decrypt_lwe(mine[0], T.encrypt_lwe_plaintext(p, mine[0], pt, 9, i))) == i % 4


def test_no_cpu_fallback_without_gpu():
    """Device entry points must fail loudly when there is no GPU (this container has none)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TfheError) as e:
        T.Context(T.TfheParams.default(True), 0)
    assert e.value.code == T.TFHE_E_CUDA
    with pytest.raises(T.TfheError) as e:          # the all-GPUs handle has no fallback either
        T.MultiGpuContext(T.TfheParams.default(True), 2)
    assert e.value.code == T.TFHE_E_CUDA


def test_mgpu_argument_checks_without_gpu():
    import ctypes as C
    L, p, h = T.lib(), T.TfheParams.default(True), C.c_void_p()
    assert L.tfhe_mgpu_create(C.byref(p), 0, None, C.byref(h)) == T.TFHE_E_PARAM          # n_gpus < 1
    bad = T.TfheParams.preset("P0", log_q=64)
    assert L.tfhe_mgpu_create(C.byref(bad), 1, None, C.byref(h)) == T.TFHE_E_PARAM        # invalid parameter set
    assert L.tfhe_mgpu_n_gpus(None) == T.TFHE_E_PARAM
    assert L.tfhe_mgpu_last_error(None) == b"null tfhe_mgpu"
    L.tfhe_mgpu_destroy(None)
    L.tfhe_mgpu_bk_free(None)
    assert L.tfhe_mgpu_bootstrap_batch(None, None, None, None, 0, None, 0, None) == T.TFHE_E_PARAM
    assert L.tfhe_ctx_get_stream(None) is None and L.tfhe_bk_get_path(None) == T.TFHE_E_PARAM


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tfhe-research_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liborc" not in src and "import orc" not in src and "tfhe_oracle.h" not in src, f


def test_wire_format_round_trip(tmp_path):
    """flat little-endian u32 files (header + words in the reference's ndarray layouts)."""
    p = T.TfheParams.default(test_cfg=True)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 7)
    for kind, arr in ((T.FILE_BSK, bsk), (T.FILE_KSK, ksk), (T.FILE_LWE_SK, lwe_sk), (T.FILE_GLWE_SK, glwe_sk)):
        path = str(tmp_path / f"k{kind}.tfhe")
        T.save_words(path, kind, p, arr)
        k2, p2, w = T.load_words(path)
        assert k2 == kind and same_params(p, p2) and np.array_equal(w, arr)
        raw = open(path, "rb").read()
        assert raw[:8] == b"TFHEB200" and len(raw) == 80 + 4 * arr.size
        assert np.array_equal(np.frombuffer(raw[80:], dtype="<u4"), arr)     # plain LE words, readable without this library
    # a truncated or foreign file is rejected, not mis-read
    bad = tmp_path / "bad.tfhe"
    bad.write_bytes(b"NOTATFHE" + bytes(100))
    with pytest.raises(T.TfheError):
        T.load_words(str(bad))
