"""Golden vectors (tests/golden/pbs_golden.json, frozen oracle outputs; generator committed next to them).

CPU part: the oracle still reproduces them and the product's seeded keygen derives the same keys.
GPU part (-m gpu): the CUDA path reproduces them WITHOUT executing the oracle.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import tfhe_research_b200 as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "pbs_golden.json")))


def product_keys(case):
    p = T.TfheParams.default(True)
    for f, v in case["params"].items():
        setattr(p, f, v)
    keys = T.bootstrapping_key_gen(p, case["key_seed"])
    lwe_sk, glwe_sk, bsk, ksk = keys
    sha = hashlib.sha256(bsk.tobytes() + ksk.tobytes() + lwe_sk.tobytes() + glwe_sk.tobytes()).hexdigest()
    assert sha == case["key_sha256"], "seeded keygen no longer reproduces the golden key material"
    return p, keys


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_reproduces_golden(case):
    from oracle import orc
    p, (lwe_sk, glwe_sk, bsk, ksk) = product_keys(case)
    o = orc.params(**case["params"])
    tvs = [orc.test_vector_identity(o), orc.test_vector_from_lut(o, case["lut"])]
    for ct, g in zip(case["lwe_in"], case["bootstrap"]):
        ct = np.array(ct, dtype=np.uint32)
        assert orc.bootstrap(o, ct, bsk, ksk, tvs[g["tv"]]).tolist() == g["out"]
        assert hashlib.sha256(orc.blind_rotate(o, ct, bsk, tvs[g["tv"]]).tobytes()).hexdigest() == g["acc_sha256"]
    for g in case["gates"]:
        c0, c1 = (np.array(case["lwe_in"][g[k]], dtype=np.uint32) for k in ("ct0", "ct1"))
        assert orc.gate(o, g["op"], c0, c1, bsk, ksk).tolist() == g["out"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fft", "ntt"])
@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_gpu_reproduces_golden(case, path):
    """Both arithmetic paths (exact FP64 FFT with a limb-split key; 2-prime NTT) reproduce the frozen vectors."""
    p, (lwe_sk, glwe_sk, bsk, ksk) = product_keys(case)
    ctx = T.Context(p, 0, path=T.PATH_FFT if path == "fft" else T.PATH_NTT)
    bk = ctx.upload_key(bsk, ksk)
    tvs = np.stack([T.construct_identity_test_vector(p), T.construct_test_from_lut(p, case["lut"])])
    cts = np.array(case["lwe_in"], dtype=np.uint32)
    idx = np.array([g["tv"] for g in case["bootstrap"]], dtype=np.uint32)
    out = ctx.bootstrap(bk, cts, tvs, idx)
    acc = ctx.blind_rotate(bk, cts, tvs, idx)
    for i, g in enumerate(case["bootstrap"]):
        assert out[i].tolist() == g["out"], i
        assert hashlib.sha256(np.ascontiguousarray(acc[i]).tobytes()).hexdigest() == g["acc_sha256"], i
    for g in case["gates"]:
        r = ctx.gate(bk, g["op"], cts[g["ct0"]][None], cts[g["ct1"]][None])
        assert r[0].tolist() == g["out"], g["op"]
    bk.free()
    ctx.close()
