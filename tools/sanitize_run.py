#!/usr/bin/env python
"""Tiny workload for compute-sanitizer (memcheck / racecheck): a few PBS + gates on the smallest configs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_research_b200 as T

for preset, n in (("P0", 2), ("P1", 2), ("P2", 1)):
    p = T.TfheParams.preset(preset, lwe_dimension=n)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 1)
    ctx = T.Context(p, 0)
    bk = ctx.upload_key(bsk, ksk)
    cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % 4), 1, i) for i in range(3)])
    out = ctx.bootstrap(bk, cts, T.construct_identity_test_vector(p))
    if p.log_p == 2:
        ctx.gate(bk, T.NAND, cts, cts[::-1].copy())
    g = ctx.external_product(bk, np.zeros(3, dtype=np.uint32), np.ones((3, p.k + 1, p.N), dtype=np.uint32))
    print(preset, "ok", [T.decode_rounded(p, T.decrypt_lwe(lwe_sk, o)) for o in out])
    bk.free(); ctx.close()
