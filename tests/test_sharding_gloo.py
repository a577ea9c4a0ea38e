"""Host-side multi-GPU logic on CPU: world_size 2 over gloo (no GPU).  The compute function is mocked;
what is tested is the sharding, padding, scatter/gather/all-gather plumbing and the layered-circuit driver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tfhe_research_b200 as T
from tfhe_research_b200 import circuit, sharding

ROW = 9  # mock ciphertext length (n + 1)


def test_shard_range_covers_everything():
    for total in (0, 1, 5, 7, 4096, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(total, world)


def mock_gate(ops, ct0, ct1):
    """'ciphertext' = vector whose last slot holds the bit; mask slots get a deterministic mix (order-sensitive)."""
    out = ct0 * 3 + ct1 * 5 + 1
    f = torch.tensor([[0, 0, 0, 1], [0, 1, 1, 1], [0, 1, 1, 0], [1, 1, 1, 0], [1, 0, 0, 0], [1, 0, 0, 1]], dtype=torch.int64)
    l, r = ct1[:, -1].long(), ct0[:, -1].long()
    out[:, -1] = f[torch.as_tensor(np.asarray(ops), dtype=torch.long), 2 * l + r].to(out.dtype)
    return out


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        full = torch.from_numpy(rng.integers(0, 1 << 31, (total, ROW)).astype(np.int32))
        # scatter -> compute -> gather
        out = sharding.bootstrap_sharded(lambda x: x * 3 + 1, full if rank == 0 else None, total, ROW, "cpu", torch.int32)
        if rank == 0:
            assert torch.equal(out, full * 3 + 1)
        else:
            assert out is None
        lo, hi = sharding.shard_range(total, rank, world)
        ag = sharding.all_gather_rows(full[lo:hi], total)
        assert torch.equal(ag, full)
        # layered circuit: 24 inputs, levels of 17, 5 and 1 gates (uneven shards, incl. an empty shard)
        bits = rng.integers(0, 2, 24)
        wires = torch.zeros((24, ROW), dtype=torch.int32)
        wires[:, -1] = torch.from_numpy(bits.astype(np.int32))
        levels = circuit.random_layered_circuit(24, [17, 5, 1], seed=3, ops=tuple(range(6)))
        res = circuit.evaluate_encrypted(levels, wires, mock_gate)
        assert res[:, -1].tolist() == circuit.evaluate_plain(levels, bits).tolist()
        q.put((rank, res.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("total", [7, 64])
def test_world2_gloo_matches_single_process(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(results[0], results[1])
    # single-process evaluation of the same circuit gives the same wires bit for bit
    rng = np.random.default_rng(0)
    rng.integers(0, 1 << 31, (total, ROW))
    bits = rng.integers(0, 2, 24)
    wires = torch.zeros((24, ROW), dtype=torch.int32)
    wires[:, -1] = torch.from_numpy(bits.astype(np.int32))
    levels = circuit.random_layered_circuit(24, [17, 5, 1], seed=3, ops=tuple(range(6)))
    single = circuit.evaluate_encrypted(levels, wires, mock_gate)
    assert np.array_equal(single.numpy(), results[0])
