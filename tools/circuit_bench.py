#!/usr/bin/env python
"""BASELINE config #4: boolean-circuit workload on bootstrapped gates (boolean.rs), sharded across GPUs.

  python tools/circuit_bench.py [--gates 65536] [--levels 16 --width 4096]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/circuit_bench.py ...

(i)  depth-1: `--gates` independent gates uniform over {NAND, AND, XOR}: inputs scattered from rank 0,
     results gathered to rank 0 (NCCL scatter / gather, keys replicated);
(ii) layered: `--levels` x `--width` gates, each reading two random wires of the previous level; one
     NCCL all-gather per level.
Both are verified on rank 0 by decrypting every output wire and comparing with the plain evaluation.
Prints one JSON line (gates/s; device time, max over ranks).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="P0")
    ap.add_argument("--gates", type=int, default=65536)
    ap.add_argument("--levels", type=int, default=16)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--inputs", type=int, default=1024)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import tfhe_research_b200 as T
    from tfhe_research_b200 import circuit, sharding

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = T.TfheParams.preset(a.preset)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)   # same seed on every rank: replicated keys
    ctx = T.Context(p, local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    bk = ctx.upload_key(bsk, ksk)
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, a.inputs)
    row = p.n + 1
    wires_h = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, int(b)), 1, i) for i, b in enumerate(bits)])
    wires = torch.from_numpy(wires_h.view(np.int32)).cuda()  # every rank encrypts the same inputs (seeded)

    def gate_fn(ops, ct0, ct1):
        return ctx.gate(bk, np.ascontiguousarray(ops), ct0, ct1)

    def dec(t):
        return [T.decode_rounded(p, T.decrypt_lwe(lwe_sk, r)) for r in t.cpu().numpy().view(np.uint32)]

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms.item())

    # (i) depth-1
    G = a.gates
    ops = rng.choice(np.array([T.NAND, T.AND, T.XOR], dtype=np.uint8), G)
    il, ir = rng.integers(0, a.inputs, G), rng.integers(0, a.inputs, G)

    def depth1():
        root0 = wires.index_select(0, torch.as_tensor(ir, device="cuda")) if rank == 0 else None   # right input -> ct0
        root1 = wires.index_select(0, torch.as_tensor(il, device="cuda")) if rank == 0 else None   # left  input -> ct1
        c0 = sharding.scatter_rows(root0, (row,), G, "cuda", torch.int32)
        c1 = sharding.scatter_rows(root1, (row,), G, "cuda", torch.int32)
        lo, hi = sharding.shard_range(G, rank, world)
        out = gate_fn(ops[lo:hi], c0.contiguous(), c1.contiguous()) if hi > lo else c0
        return sharding.gather_rows(out, G)

    depth1()  # warm-up
    out1, ms1 = timed(depth1)
    # (ii) layered
    levels = circuit.random_layered_circuit(a.inputs, [a.width] * a.levels, seed=3)
    circuit.evaluate_encrypted(levels[:1], wires, gate_fn)  # warm-up
    out2, ms2 = timed(lambda: circuit.evaluate_encrypted(levels, wires, gate_fn))
    if rank == 0:
        f = {T.NAND: lambda l, r: 1 - (l & r), T.AND: lambda l, r: l & r, T.XOR: lambda l, r: l ^ r}
        exp1 = [f[int(o)](int(bits[l]), int(bits[r])) for o, l, r in zip(ops, il, ir)]
        ok1 = dec(out1) == exp1
        ok2 = dec(out2) == circuit.evaluate_plain(levels, bits).tolist()
        print(json.dumps({"workload": f"{a.preset} boolean circuit (BASELINE config #4)", "n_gpus": world,
                          "depth1": {"gates": G, "ms": ms1, "gates_per_s": G / (ms1 * 1e-3), "correct": ok1,
                                     "collectives": "scatter + gather (NCCL)" if world > 1 else "none"},
                          "layered": {"levels": a.levels, "width": a.width, "ms": ms2, "gates_per_s": a.levels * a.width / (ms2 * 1e-3),
                                      "correct": ok2, "collectives": "one all-gather per level (NCCL)" if world > 1 else "none"}}))
        assert ok1 and ok2
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
