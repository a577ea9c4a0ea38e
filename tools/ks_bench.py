#!/usr/bin/env python
"""Key-switch kernels side by side (device-resident inputs, CUDA-event time of the key-switch part of tfhe_key_switch):
  python tools/ks_bench.py -> JSON: per preset and batch, ms for KS_IMAD / KS_MMA / KS_TCGEN05, outputs compared bit for bit."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tfhe_research_b200 as T
out = {}
for preset, B in (("P1", 4096), ("P0", 4096), ("P2", 16384)):
    p = T.TfheParams.preset(preset)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    ctx = T.Context(p, 0)
    bk = ctx.upload_key(bsk, ksk)
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.integers(0, 1 << 32, (B, p.k * p.N + 1), dtype=np.uint64).astype(np.uint32).view(np.int32)).cuda()
    res, rec = {}, {}
    for name, path in (("imad", T.KS_IMAD), ("mma_sync", T.KS_MMA), ("tcgen05", T.KS_TCGEN05)):
        ctx.set_ks_path(path)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(torch.cuda.current_stream())
            y = ctx.key_switch(bk, x)
            e1.record(torch.cuda.current_stream())
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = y.cpu().numpy()
        rec[name + "_call_ms"] = min(ts[1:])
    rec["identical"] = bool(np.array_equal(res["imad"], res["mma_sync"]) and np.array_equal(res["imad"], res["tcgen05"]))
    # kernel-only time through the bootstrap call's own events
    cts = torch.from_numpy(rng.integers(0, 1 << 32, (B, p.n + 1), dtype=np.uint64).astype(np.uint32).view(np.int32)).cuda()
    tv = torch.from_numpy(T.construct_identity_test_vector(p).view(np.int32).copy()).cuda()
    for name, path in (("imad", T.KS_IMAD), ("mma_sync", T.KS_MMA), ("tcgen05", T.KS_TCGEN05)):
        ctx.set_ks_path(path)
        ks = []
        for _ in range(3):
            ctx.bootstrap(bk, cts, tv)
            ks.append(ctx.last_timing()["key_switch_ms"])
        rec[name + "_kernels_ms"] = min(ks[1:])
    out[f"{preset}_b{B}"] = rec
    bk.free(); ctx.close()
print(json.dumps(out))
