import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import tfhe_research_b200 as T
for preset in ("P1", "P0"):
    p = T.TfheParams.preset(preset)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    ctx = T.Context(p, 0); bk = ctx.upload_key(bsk, ksk)
    ctx.set_latency_config(int(os.environ.get('LATMODE', '3')))
    cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % 4), 1, i) for i in range(8)])
    tv = T.construct_identity_test_vector(p)
    for B in (1, 8):
        for _ in range(2):
            out = ctx.bootstrap(bk, cts[:B], tv)
        print(preset, B, ctx.last_timing(), file=sys.stderr)
