#!/usr/bin/env python
"""Generates tests/golden/pbs_golden.json from the CPU oracle (run in the build container).

The reference holds no golden ciphertexts (thread_rng everywhere) and cannot be run here, so these are
ORACLE outputs frozen at a known-good commit: they guard the oracle against regressions (CPU suite) and
let the GPU suite check the kernels without executing the oracle.  Keys are not stored: they are
re-derived from `key_seed` by the seeded keygen (bit-identical in oracle and product); a SHA-256 of the
key material pins them.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

CASES = [
    ("P0t", dict(), 0xB200),                                                                                         # lib.rs:77-99 cfg(test)
    ("P1n3", dict(glwe_dimension=1, glwe_poly_degree=10, lwe_dimension=3, pbs_log_base=8, pbs_levels=3, ks_log_base=2, ks_levels=8), 0xB201),
    ("P2n2", dict(glwe_dimension=1, glwe_poly_degree=11, lwe_dimension=2, pbs_log_base=8, pbs_levels=3, ks_log_base=4, ks_levels=5, log_p=4), 0xB202),
]


def main():
    out = {"generator": "tests/golden/make_golden.py", "cases": []}
    for name, over, seed in CASES:
        p = orc.params(True, **over)
        lwe_sk, glwe_sk, bsk, ksk = orc.keygen(p, seed)
        pm = 1 << p.log_p
        rng = np.random.default_rng(seed)
        cts = [orc.lwe_encrypt(p, lwe_sk, m % pm, 7, m) for m in range(6)]
        cts.append(rng.integers(0, 1 << 32, p.n + 1, dtype=np.uint64).astype(np.uint32))   # arbitrary words
        lut = rng.integers(0, pm, pm).astype(np.uint32); lut[0] = 0
        tvs = [orc.test_vector_identity(p), orc.test_vector_from_lut(p, lut)]
        case = {"name": name, "params": {f: getattr(p, f) for f, _ in orc.OrcParams._fields_}, "key_seed": seed,
                "key_sha256": hashlib.sha256(bsk.tobytes() + ksk.tobytes() + lwe_sk.tobytes() + glwe_sk.tobytes()).hexdigest(),
                "lut": lut.tolist(), "lwe_in": [c.tolist() for c in cts], "bootstrap": [], "gates": []}
        for i, c in enumerate(cts):
            tv = tvs[i % 2]
            case["bootstrap"].append({"tv": i % 2, "out": orc.bootstrap(p, c, bsk, ksk, tv).tolist(),
                                      "acc_sha256": hashlib.sha256(orc.blind_rotate(p, c, bsk, tv).tobytes()).hexdigest()})
        if p.log_p == 2:
            for op in range(6):
                case["gates"].append({"op": op, "ct0": 1, "ct1": 3, "out": orc.gate(p, op, cts[1], cts[3], bsk, ksk).tolist()})
        out["cases"].append(case)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbs_golden.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
