#!/usr/bin/env python
"""Small driver for profiling / A-B runs: one preset, one batch, device-resident inputs.

  python tools/prof_run.py --preset P1 --batch 592 --steps 1 [--lib path/to/variant.so] [--check]
Prints one JSON line with the blind-rotation and key-switch kernel times (CUDA events inside the lib).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="P1")
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--n", type=int, default=0, help="override lwe_dimension")
    ap.add_argument("--lib", default="")
    ap.add_argument("--check", action="store_true", help="decrypt-check the outputs")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import tfhe_research_b200 as T
    if a.lib:
        T.LIB_PATH = os.path.abspath(a.lib)
    import torch
    over = {"lwe_dimension": a.n} if a.n else {}
    p = T.TfheParams.preset(a.preset, **over)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    ctx = T.Context(p, 0)
    bk = ctx.upload_key(bsk, ksk)
    pm = 1 << p.log_p
    nu = min(a.batch, 64)
    uniq = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(nu)])
    cts = np.tile(uniq, ((a.batch + nu - 1) // nu, 1))[:a.batch]
    d_in = torch.from_numpy(cts.view(np.int32).copy()).cuda()
    d_tv = torch.from_numpy(T.construct_identity_test_vector(p).view(np.int32).copy()).cuda()
    d_out = torch.empty((a.batch, p.n + 1), dtype=torch.int32, device="cuda")
    br, ks = [], []
    for s in range(a.warmup + a.steps):
        ctx.bootstrap(bk, d_in, d_tv, out=d_out)
        if s >= a.warmup:
            t = ctx.last_timing()
            br.append(t["blind_rotate_ms"]); ks.append(t["key_switch_ms"])
    ok = None
    if a.check:
        res = d_out.cpu().numpy().view(np.uint32)
        ok = all(T.decode_rounded(p, T.decrypt_lwe(lwe_sk, res[i])) == (i % nu) % pm for i in range(0, a.batch, max(1, a.batch // 32)))
    print(json.dumps({"tag": a.tag or os.path.basename(a.lib or "default"), "preset": a.preset, "batch": a.batch, "n": p.n,
                      "blind_rotate_ms": min(br), "key_switch_ms": min(ks), "pbs_per_s": a.batch / ((min(br) + min(ks)) * 1e-3), "ok": ok}))


if __name__ == "__main__":
    main()
