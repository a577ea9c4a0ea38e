#!/usr/bin/env python
"""Worker of tests/test_gpu_multigpu.py: N ranks (torchrun) bootstrap one batch through sharding.bootstrap_sharded and
evaluate one layered circuit through circuit.evaluate_encrypted; rank 0 checks both bit for bit against the same
work done on its GPU alone (N-GPU output == 1-GPU output).

  --backend nccl : one GPU per rank, CUDA tensors travel over NCCL (needs >= world GPUs)
  --backend gloo : every rank drives cuda:0 (ranks emulated on one GPU, B200_PROFILING.md), the shards travel as CPU
                   tensors over gloo; the compute is the real CUDA path either way.
Exit code 0 = equal.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--preset", default="P1")
    ap.add_argument("--lwe-dim", dest="n", type=int, default=12)
    ap.add_argument("--batch", type=int, default=1001)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import tfhe_research_b200 as T
    from tfhe_research_b200 import circuit, sharding

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    nccl = a.backend == "nccl"
    dev_index = local_rank if nccl else 0
    torch.cuda.set_device(dev_index)
    if nccl:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev_index))
    else:
        dist.init_process_group("gloo")
    comm_dev = torch.device("cuda", dev_index) if nccl else torch.device("cpu")

    p = T.TfheParams.preset(a.preset, lwe_dimension=a.n)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)     # same seed on every rank: replicated keys
    ctx = T.Context(p, dev_index)
    bk = ctx.upload_key(bsk, ksk)
    pm = 1 << p.log_p
    row = p.n + 1
    tv = T.construct_identity_test_vector(p)

    def to_comm(x_np):
        return torch.from_numpy(x_np.view(np.int32)).to(comm_dev)

    def bootstrap(t):          # tensor on comm_dev -> tensor on comm_dev, through the C ABI on this rank's GPU
        if nccl:
            return ctx.bootstrap(bk, t.contiguous(), torch.from_numpy(tv.view(np.int32)).cuda())
        return torch.from_numpy(ctx.bootstrap(bk, t.numpy().view(np.uint32), tv).view(np.int32))

    def gate_fn(ops, ct0, ct1):
        if nccl:
            return ctx.gate(bk, np.ascontiguousarray(ops), ct0, ct1)
        return torch.from_numpy(ctx.gate(bk, np.ascontiguousarray(ops), ct0.numpy().view(np.uint32), ct1.numpy().view(np.uint32)).view(np.int32))

    # ---- a ragged batch: scatter from rank 0 -> PBS -> gather on rank 0
    B = a.batch
    uniq = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(64)])
    cts = np.tile(uniq, ((B + 63) // 64, 1))[:B].copy()
    cts[:, 0] += np.arange(B, dtype=np.uint32) * np.uint32(0x9E3779B1)   # every ciphertext distinct (still valid inputs)
    root = to_comm(cts) if rank == 0 else None
    out = sharding.bootstrap_sharded(bootstrap, root, B, row, comm_dev, torch.int32)
    ok = True
    if rank == 0:
        single = ctx.bootstrap(bk, cts, tv)
        ok &= bool(np.array_equal(out.cpu().numpy().view(np.uint32), single))
        print(f"[dist_worker] sharded PBS batch {B} over {world} ranks == single GPU: {ok}", flush=True)

    # ---- a layered circuit: uneven level widths (incl. widths smaller than the world), one all-gather per level
    rng = np.random.default_rng(3)
    n_in = 32
    bits = rng.integers(0, 2, n_in)
    wires_np = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, int(b)), 2, i) for i, b in enumerate(bits)])
    levels = circuit.random_layered_circuit(n_in, [101, 64, 7, 1], seed=5)
    res = circuit.evaluate_encrypted(levels, to_comm(wires_np), gate_fn)
    if rank == 0:
        w = wires_np
        for lv in levels:                                    # the same circuit on this GPU alone
            w = ctx.gate(bk, lv.ops, w[lv.right], w[lv.left])
        same = bool(np.array_equal(res.cpu().numpy().view(np.uint32), w))
        dec = [T.decode_rounded(p, T.decrypt_lwe(lwe_sk, r)) for r in w]
        plain = circuit.evaluate_plain(levels, bits).tolist()
        print(f"[dist_worker] layered circuit over {world} ranks == single GPU: {same}; decrypts to the plain evaluation: {dec == plain}", flush=True)
        ok &= same and dec == plain
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=comm_dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
