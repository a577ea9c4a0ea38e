#!/usr/bin/env python
"""Single-ciphertext / small-batch latency of the PBS call (BASELINE metric "per-PBS latency"), both kernel configurations.
  python tools/latency_run.py   -> one JSON line: for each preset and batch in {1, 8, 148}: total call ms (device pointers)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import tfhe_research_b200 as T
    out = {}
    for preset in ("P1", "P0", "P2"):
        p = T.TfheParams.preset(preset)
        lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
        ctx = T.Context(p, 0)
        bk = ctx.upload_key(bsk, ksk)
        pm = 1 << p.log_p
        cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(148)])
        d_in = torch.from_numpy(cts.view(np.int32)).cuda()
        d_tv = torch.from_numpy(T.construct_identity_test_vector(p).view(np.int32).copy()).cuda()
        rec = {}
        for B in (1, 8, 24, 148):
            for name, on in (("cluster_split", 4), ("cluster", 3), ("all_teams", 2), ("deep_ring", 1), ("throughput_cfg", 0)):
                ctx.set_latency_config(on)
                x = d_in[:B].contiguous()
                o = torch.empty((B, p.n + 1), dtype=torch.int32, device="cuda")
                ts = []
                for _ in range(5):
                    ctx.bootstrap(bk, x, d_tv, out=o)
                    t = ctx.last_timing()
                    ts.append((t["total_ms"], t["blind_rotate_ms"], t["key_switch_ms"]))
                best = min(ts[1:])
                rec[f"b{B}_{name}"] = {"total_ms": best[0], "blind_rotate_ms": best[1], "key_switch_ms": best[2]}
                res = o.cpu().numpy().view(np.uint32)
                assert all(T.decode_rounded(p, T.decrypt_lwe(lwe_sk, res[i])) == i % pm for i in range(B))
        if preset == "P0":       # config #1: one bootstrapped NAND
            c0 = torch.from_numpy(T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, 1), 1, 9001).view(np.int32)[None].copy()).cuda()
            c1 = torch.from_numpy(T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, 1), 1, 9002).view(np.int32)[None].copy()).cuda()
            for name, on in (("cluster_split", 4), ("cluster", 3), ("all_teams", 2), ("deep_ring", 1), ("throughput_cfg", 0)):
                ctx.set_latency_config(on)
                ts = []
                for _ in range(5):
                    ctx.gate(bk, T.NAND, c0, c1)
                    ts.append(ctx.last_timing()["total_ms"])
                rec[f"single_nand_{name}_ms"] = min(ts[1:])
        out[preset] = rec
        bk.free()
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
