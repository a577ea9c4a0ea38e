#!/usr/bin/env python
"""Check a golden-vector dump of the REAL Rust crate against the C oracle and the CUDA path, word for word.

  # on a machine with cargo, in a checkout of Janmajayamall/tfhe-research:
  git apply <this repo>/rust/reference_dump.patch
  TFHE_DUMP_DIR=/tmp/tfhe_dump cargo test --release dump_golden -- --nocapture
  # here:
  python tools/check_reference_dump.py /tmp/tfhe_dump            # oracle always; the GPU path when a CUDA device is present
  python tools/check_reference_dump.py --self-test               # writes the same file set from the ORACLE and checks it
                                                                 # (exercises the format and this script; proves nothing about Rust)

This is what un-caps "parity unpinned against the Rust binary" (DESIGN.md section 2): the dump holds keys, inputs and
the outputs of the crate's own `bootstrap` (bootstrapping.rs:58-120) and `and` / `or` (boolean.rs:9-53); `bootstrap` is a
deterministic function of (ciphertext, BSK, KSK, test vector), so the oracle and the kernels must reproduce every output
word.  File format: tools/golden_format.md.  Exit code 0 = every word equal.
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FILES = ["lwe_sk", "glwe_sk", "bsk", "ksk", "tv_identity", "pbs_in", "pbs_out_identity", "tv_and", "tv_or",
         "gate_ct0", "gate_ct1", "gate_and_out", "gate_or_out"]
KIND = {"lwe_sk": 5, "glwe_sk": 6, "bsk": 3, "ksk": 4, "tv_identity": 7, "tv_and": 7, "tv_or": 7}


def load(T, d):
    out, params = {}, None
    for name in FILES:
        kind, p, w = T.load_words(os.path.join(d, name + ".bin"))
        assert kind == KIND.get(name, 1), f"{name}: file kind {kind}"
        if params is None:
            params = p
        assert all(getattr(p, f) == getattr(params, f) for f, _ in T.TfheParams._fields_), f"{name}: parameter header differs"
        out[name] = w
    return params, out


def self_test_dump(T, d):
    """The file set rust/reference_dump.patch writes, produced by the oracle instead (n = 4 like cfg(test))."""
    from oracle import orc
    p = T.TfheParams.default(test_cfg=True)
    o = orc.params(True)
    lwe_sk, glwe_sk, bsk, ksk = orc.keygen(o, 0x5E1F)
    pm = 1 << p.log_p
    tv = orc.test_vector_identity(o)
    cin = np.stack([orc.lwe_encrypt(o, lwe_sk, i % pm, 11, i) for i in range(2 * pm)])
    cout = np.stack([orc.bootstrap(o, c, bsk, ksk, tv) for c in cin])
    c1 = np.stack([orc.lwe_encrypt(o, lwe_sk, (i >> 1) & 1, 12, i) for i in range(4)])
    c0 = np.stack([orc.lwe_encrypt(o, lwe_sk, i & 1, 13, i) for i in range(4)])
    data = {"lwe_sk": lwe_sk, "glwe_sk": glwe_sk, "bsk": bsk, "ksk": ksk, "tv_identity": tv, "pbs_in": cin, "pbs_out_identity": cout,
            "tv_and": orc.test_vector_boolean(o, 0), "tv_or": orc.test_vector_boolean(o, 1), "gate_ct0": c0, "gate_ct1": c1,
            "gate_and_out": np.stack([orc.gate(o, 0, c0[i], c1[i], bsk, ksk) for i in range(4)]),
            "gate_or_out": np.stack([orc.gate(o, 1, c0[i], c1[i], bsk, ksk) for i in range(4)])}
    for name, w in data.items():
        T.save_words(os.path.join(d, name + ".bin"), KIND.get(name, 1), p, w)


def check(d, want_gpu=None, quiet=False):
    import tfhe_research_b200 as T
    from oracle import orc
    say = (lambda *a: None) if quiet else print
    p, w = load(T, d)
    p.validate()
    o = orc.params(**{f: getattr(p, f) for f, _ in T.TfheParams._fields_})
    n, row = p.n, p.n + 1
    assert w["lwe_sk"].size == n and w["glwe_sk"].size == p.k * p.N and w["bsk"].size == p.bsk_words and w["ksk"].size == p.ksk_words
    cin, cout = w["pbs_in"].reshape(-1, row), w["pbs_out_identity"].reshape(-1, row)
    c0, c1 = w["gate_ct0"].reshape(-1, row), w["gate_ct1"].reshape(-1, row)
    g_and, g_or = w["gate_and_out"].reshape(-1, row), w["gate_or_out"].reshape(-1, row)
    bad = 0
    # host-side functions: test vectors (test_vector.rs:5-67)
    for name, mine in (("tv_identity", orc.test_vector_identity(o)), ("tv_and", orc.test_vector_boolean(o, 0)), ("tv_or", orc.test_vector_boolean(o, 1))):
        ok = np.array_equal(mine, w[name])
        bad += not ok
        say(f"oracle {name:<18} {'==' if ok else '!='} reference")
    for name, mine in (("tv_identity", T.construct_identity_test_vector(p)), ("tv_and", T.construct_test_vector_boolean(p, T.AND)), ("tv_or", T.construct_test_vector_boolean(p, T.OR))):
        ok = np.array_equal(mine, w[name])
        bad += not ok
        say(f"host   {name:<18} {'==' if ok else '!='} reference")
    # the oracle against the reference's outputs
    pbs = np.stack([orc.bootstrap(o, c, w["bsk"], w["ksk"], w["tv_identity"]) for c in cin])
    ands = np.stack([orc.gate(o, 0, c0[i], c1[i], w["bsk"], w["ksk"]) for i in range(len(c0))])
    ors = np.stack([orc.gate(o, 1, c0[i], c1[i], w["bsk"], w["ksk"]) for i in range(len(c0))])
    for name, mine, ref in (("bootstrap", pbs, cout), ("and", ands, g_and), ("or", ors, g_or)):
        ok = np.array_equal(mine, ref)
        bad += not ok
        say(f"oracle {name:<18} {'==' if ok else '!='} reference   ({ref.shape[0]} ciphertexts, {ref.size} words)")
    # decrypt-level sanity of the dump itself (rounding decoder)
    pm = 1 << p.log_p
    dec = [orc.lwe_decrypt_round(o, w["lwe_sk"], c) for c in cout]
    say(f"reference bootstrap outputs decrypt to {dec} (inputs were messages i mod {pm})")
    # the CUDA path
    import torch
    have_gpu = torch.cuda.is_available()
    if want_gpu is True and not have_gpu:
        raise SystemExit("no CUDA device")
    if have_gpu and want_gpu is not False:
        for path, pname in ((T.PATH_FFT, "fft"), (T.PATH_NTT, "ntt")):
            ctx = T.Context(p, 0, path=path)
            bk = ctx.upload_key(w["bsk"], w["ksk"])
            res = (("bootstrap", ctx.bootstrap(bk, cin, w["tv_identity"]), cout), ("and", ctx.gate(bk, T.AND, c0, c1), g_and), ("or", ctx.gate(bk, T.OR, c0, c1), g_or))
            for name, mine, ref in res:
                ok = np.array_equal(mine, ref)
                bad += not ok
                say(f"cuda[{pname}] {name:<15} {'==' if ok else '!='} reference")
            bk.free()
            ctx.close()
    else:
        say("no CUDA device: GPU comparison skipped")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir", nargs="?")
    ap.add_argument("--self-test", action="store_true")
    a = ap.parse_args()
    if a.self_test:
        import tfhe_research_b200 as T
        with tempfile.TemporaryDirectory() as d:
            self_test_dump(T, d)
            bad = check(d)
    else:
        if not a.dir:
            ap.error("give the dump directory (or --self-test)")
        bad = check(a.dir)
    print("ALL EQUAL" if bad == 0 else f"{bad} MISMATCHING GROUPS")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
