#!/usr/bin/env python
"""BASELINE config #5 (single-GPU slice): BMMP-style unrolled-by-two bootstrapping with key switching vs the standard
chain on the same parameter set (P1).  One JSON line; device-resident inputs, CUDA events inside the library."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import tfhe_research_b200 as T
    p = T.TfheParams.preset("P1")
    lwe_sk, glwe_sk, bsk3, ksk = T.bootstrapping_key_gen_bmmp(p, 0xB200)
    ctx = T.Context(p, 0, path=T.PATH_FFT)
    bk = ctx.upload_key_bmmp(bsk3, ksk)
    pm = 1 << p.log_p
    nu = 64
    uniq = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(nu)])
    cts = np.tile(uniq, ((a.batch + nu - 1) // nu, 1))[:a.batch]
    d_in = torch.from_numpy(cts.view(np.int32).copy()).cuda()
    d_tv = torch.from_numpy(T.construct_identity_test_vector(p).view(np.int32).copy()).cuda()
    d_out = torch.empty((a.batch, p.n + 1), dtype=torch.int32, device="cuda")
    br, ks = [], []
    for s in range(1 + a.steps):
        ctx.bootstrap(bk, d_in, d_tv, out=d_out)
        if s:
            t = ctx.last_timing(); br.append(t["blind_rotate_ms"]); ks.append(t["key_switch_ms"])
    res = d_out.cpu().numpy().view(np.uint32)
    ok = all(T.decode_rounded(p, T.decrypt_lwe(lwe_sk, res[i])) == (i % nu) % pm for i in range(0, a.batch, max(1, a.batch // 64)))
    print(json.dumps({"config": "BMMP unrolled-by-two + key switch, P1 (k=1 N=1024 n=630), FFT path", "batch": a.batch,
                      "blind_rotate_ms": min(br), "key_switch_ms": min(ks), "pbs_per_s": a.batch / ((min(br) + min(ks)) * 1e-3),
                      "key_bytes": int(3 * (p.n // 2) * (p.k + 1) * p.pbs_levels * 2 * (p.k + 1) * (p.N // 2) * 16), "decrypts_ok": ok}))


if __name__ == "__main__":
    main()
