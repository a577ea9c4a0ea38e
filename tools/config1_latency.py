#!/usr/bin/env python
"""BASELINE config #1: ONE bootstrapped NAND gate with the reference's default parameters (lib.rs:101-123):
latency on the GPU (batch 1) next to the C port of the reference's CPU path on one host core (faithful
Toeplitz algorithm); outputs compared bit for bit."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_research_b200 as T
from oracle import orc

p = T.TfheParams.default()
o = orc.params()
lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
ctx = T.Context(p, 0)
bk = ctx.upload_key(bsk, ksk)
ct0 = T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, 1), 1, 0)[None]
ct1 = T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, 1), 1, 1)[None]
lat = []
for _ in range(5):
    t0 = time.perf_counter(); out = ctx.gate(bk, T.NAND, ct0, ct1); lat.append((time.perf_counter() - t0) * 1e3)
dev = ctx.last_timing()
orc.lib().orc_set_faithful_toeplitz(1)
t0 = time.perf_counter(); exp = orc.gate(o, 3, ct0[0], ct1[0], bsk, ksk); cpu_s = time.perf_counter() - t0
orc.lib().orc_set_faithful_toeplitz(0)
print(json.dumps({"config": "single bootstrapped NAND, reference defaults k=2 N=512 n=722", "bit_exact": bool(np.array_equal(out[0], exp)),
                  "decrypts_to": T.decode_rounded(p, T.decrypt_lwe(lwe_sk, out[0])),
                  "gpu_latency_ms_host_call": min(lat[1:]), "gpu_device_ms": dev, "cpu_port_1core_s": cpu_s,
                  "cpu_note": "C restatement of the reference's Rust path (Toeplitz O(N^2)), one host core"}))
