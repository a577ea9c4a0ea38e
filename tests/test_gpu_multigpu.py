"""Multi-GPU paths on real hardware (-m gpu): N-GPU output == 1-GPU output, bit for bit.

* one process, all GPUs: the `tfhe_mgpu_*` C ABI (host batches: per-device copies; device batches: NCCL send/recv
  scatter + gather inside the library);
* one process per GPU: `sharding.bootstrap_sharded` / `circuit.evaluate_encrypted` under torchrun -- over NCCL when the
  box has >= 2 GPUs, and ALWAYS as two ranks emulated on cuda:0 over gloo (so the round-end 1-GPU run covers it too).
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import tfhe_research_b200 as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "== single GPU: True" in r.stdout
    return r.stdout


def test_two_ranks_on_one_gpu_gloo_equal_single_gpu():
    out = _torchrun(2, "--backend", "gloo", "--preset", "P1", "--lwe-dim", "12", "--batch", "1001")
    assert "layered circuit over 2 ranks == single GPU: True; decrypts to the plain evaluation: True" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_ranks_equal_single_gpu():
    n = min(torch.cuda.device_count(), 8)
    out = _torchrun(n, "--backend", "nccl", "--preset", "P0", "--lwe-dim", "10", "--batch", "2051")
    assert f"layered circuit over {n} ranks == single GPU: True; decrypts to the plain evaluation: True" in out


@pytest.mark.parametrize("preset,n", [("P1", 10), ("P0", 6)])
def test_mgpu_c_abi_equals_single_gpu(preset, n):
    G = min(torch.cuda.device_count(), 8)
    p = T.TfheParams.preset(preset, lwe_dimension=n)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    pm = 1 << p.log_p
    one = T.Context(p, 0)
    bk1 = one.upload_key(bsk, ksk)
    m = T.MultiGpuContext(p, G)
    bkm = m.upload_key(bsk, ksk)
    rng = np.random.default_rng(21)
    tvs = np.stack([T.construct_identity_test_vector(p), rng.integers(0, pm, p.N).astype(np.uint32)])
    for B in (1, G, 7 * G + 3, 1500):                 # fewer ciphertexts than GPUs, ragged shards, several waves
        cts = rng.integers(0, 1 << 32, (B, n + 1), dtype=np.uint64).astype(np.uint32)
        for i in range(min(B, 8)):
            cts[i] = T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i)
        idx = (np.arange(B) % 2).astype(np.uint32)
        ref = one.bootstrap(bk1, cts, tvs, idx)
        got = m.bootstrap(bkm, cts, tvs, idx)          # host batch: every GPU copies its own shard
        assert np.array_equal(got, ref), B
        d_in = torch.from_numpy(cts.view(np.int32)).cuda(0)
        d_out = m.bootstrap(bkm, d_in, tvs, idx)       # device batch on GPU 0: NCCL scatter / gather inside the library
        assert d_out.is_cuda and np.array_equal(d_out.cpu().numpy().view(np.uint32), ref), B
        if G > 1:                                      # the batch may live on any of the context's devices
            last = G - 1
            d_in2 = torch.from_numpy(cts.view(np.int32)).cuda(last)
            assert np.array_equal(m.bootstrap(bkm, d_in2, tvs, idx).cpu().numpy().view(np.uint32), ref), B
        if p.log_p == 2:
            ops = rng.integers(0, 6, B).astype(np.uint8)
            c1 = np.roll(cts, 1, axis=0).copy()
            gref = one.gate(bk1, ops, cts, c1)
            assert np.array_equal(m.gate(bkm, ops, cts, c1), gref), B
            dg = m.gate(bkm, ops, d_in, torch.from_numpy(c1.view(np.int32)).cuda(0))
            assert np.array_equal(dg.cpu().numpy().view(np.uint32), gref), B
    t = m.last_timing()
    assert t["total_ms"] > 0
    # errors: out-of-range test-vector index, foreign key, empty batch
    with pytest.raises(T.TfheError) as ei:
        m.bootstrap(bkm, cts[:4], tvs, np.array([0, 1, 2, 0], dtype=np.uint32))
    assert ei.value.code == T.TFHE_E_PARAM
    assert m.bootstrap(bkm, np.zeros((0, n + 1), dtype=np.uint32), tvs).shape == (0, n + 1)
    bkm.free(); m.close(); bk1.free(); one.close()


def test_lut_idx_out_of_range_and_key_lifetime():
    p = T.TfheParams.preset("P0", lwe_dimension=4)
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    for path in (T.PATH_FFT, T.PATH_NTT):
        ctx = T.Context(p, 0, path=path)
        bk = ctx.upload_key(bsk, ksk)
        tv = T.construct_identity_test_vector(p)
        cts = np.zeros((3, 5), dtype=np.uint32)
        with pytest.raises(T.TfheError) as ei:        # host index array
            ctx.bootstrap(bk, cts, tv, np.array([0, 1, 0], dtype=np.uint32))
        assert ei.value.code == T.TFHE_E_PARAM
        with pytest.raises(T.TfheError) as ei:        # device index array: checked by the kernel
            ctx.bootstrap(bk, torch.from_numpy(cts.view(np.int32)).cuda(), torch.from_numpy(tv.view(np.int32)).cuda(),
                          torch.tensor([0, 0, 7], dtype=torch.int32, device="cuda"))
        assert ei.value.code == T.TFHE_E_PARAM
        ok = ctx.bootstrap(bk, cts, tv, np.zeros(3, dtype=np.uint32))
        assert ok.shape == (3, 5)
        # the key outlives its context: destroying the context first must not crash the later free
        h = bk._h
        ctx._keys.discard(bk)
        T.lib().tfhe_ctx_destroy(ctx._h)
        ctx._h = None
        T.lib().tfhe_bk_free(h)
        bk._h = None
    with pytest.raises(ValueError):                   # wrong row width never reaches the device
        c2 = T.Context(p, 0)
        try:
            k2 = c2.upload_key(bsk, ksk)
            c2.bootstrap(k2, np.zeros((3, 9), dtype=np.uint32), tv)
        finally:
            c2.close()
